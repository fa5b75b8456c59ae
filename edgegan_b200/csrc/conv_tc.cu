// tcgen05 (5th-gen tensor core) implicit-GEMM convolution trio for sm_100a, kind::tf32.
//
//   forward / input-gradient  : D[128 pixels x BN channels] (TMEM, fp32) += A[128 px x 32 k] * B[BN x 32 k]^T
//       A = one TMA box {32 channels, bw, bh, bn} of the NHWC activation per (filter tap, 32-channel chunk):
//           the box lands in shared memory as 128 rows x 128 B with the 128-byte swizzle, which IS the
//           canonical K-major UMMA operand layout -- no im2col buffer ever exists.  Zero padding comes from
//           TMA out-of-bounds fill; stride 2 is handled by one tensor map per input-pixel parity.
//       B = filter slice [BN x 32] (K-major): the HWIO filter itself for the input gradient, a per-call
//           transposed copy (HWOI, <= 8 MB, L2 resident) for the forward.
//   filter-gradient           : D[128 x BN] += X^T[128 ci x 8 px] * DY[8 px x BN co]: both operands are the same
//       TMA pixel boxes, consumed as MN-major (channel-contiguous) UMMA operands; split over pixels with a
//       red.global.add epilogue.
//
// Forward / input-gradient kernel (conv_tc_kmajor): ONE persistent CTA per SM, 512 threads in warpgroup-aligned roles
// (setmaxnreg moves registers to the drain warps): warps 0-7 move the A operand into tensor memory (TS-mode MMA; in
// the 3xTF32 mode as hi + lo), warps 8-11 drain the chunk accumulators into fp32 registers and store finished tiles
// through a transposing staging tile (optionally with a fused activation / activation-derivative mask), warp 12 is
// the TMA producer, warp 13 the MMA issuer and TMEM owner.  Filter-gradient kernel (conv_tc_wgrad): one CTA per
// (row tile, column tile or tap pair, pixel split), 320 threads: warp 0 TMA, warp 1 MMA, warps 2-5 row operand ->
// TMEM + epilogue, warps 6-9 lo copy of the column operand.  Both use smem rings with full / ready / empty
// mbarriers; see DESIGN.md section 3.1 for the measurements behind each choice.
//
// Replaces tf.nn.conv2d / tf.nn.conv2d_transpose and their gradients
// (reference nn/modules/conv.py:26,29,49).
#include "common.cuh"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <map>
#include <mutex>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.b32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n\t"
        "@P bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// The tensor core reads the top 19 bits of an fp32 word (sign, 8 exponent, 10 mantissa) and IGNORES the rest
// (measured: tools/tc_probe.py, raw vs pre-truncated operands are bit-identical).  Truncation is biased, so:
//   mode 1 (EG_ALGO_TC):   round to nearest in place  -> unbiased TF32 (what cvt.rna.tf32.f32 would give)
//   mode 3 (EG_ALGO_TC3X): keep x as the "hi" operand and write lo = x - trunc(x) (exact in fp32) next to it;
//                          hi*hi + lo*hi + hi*lo then carries ~21 mantissa bits (fp32-class accuracy).
__device__ __forceinline__ uint32_t tf32_rna(uint32_t x) { return (x + 0x1000u) & 0xFFFFE000u; }
// lo is rounded to nearest too: its own truncation would otherwise leave a biased 2^-20 relative residual
__device__ __forceinline__ uint32_t tf32_lo(uint32_t x) { return tf32_rna(__float_as_uint(__uint_as_float(x) - __uint_as_float(x & 0xFFFFE000u))); }
// elementwise pass over `bytes` of shared memory by 128 threads (the swizzle is irrelevant: same offset in/out).
// Loads are issued in batches of 8 x 16 B per thread before any use, so the LDS latency is paid once per batch.
__device__ __forceinline__ void condition_tile(uint32_t src, uint32_t dst_lo, uint32_t bytes, int tid, int mode) {
    constexpr uint32_t kStep = 128u * 16u;                   // bytes covered by one pass of the 128 threads
    for (uint32_t base = (uint32_t)tid * 16u; base < bytes; base += 8u * kStep) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (base + j * kStep < bytes) v[j] = lds128(src + base + j * kStep);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (base + j * kStep < bytes) {
                uint4 o;
                if (mode == 3) {
                    o.x = tf32_lo(v[j].x); o.y = tf32_lo(v[j].y); o.z = tf32_lo(v[j].z); o.w = tf32_lo(v[j].w);
                    sts128(dst_lo + base + j * kStep, o);
                } else {
                    o.x = tf32_rna(v[j].x); o.y = tf32_rna(v[j].y); o.z = tf32_rna(v[j].z); o.w = tf32_rna(v[j].w);
                    sts128(src + base + j * kStep, o);
                }
            }
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// L2 prefetches (no shared-memory destination, no barrier) of the operand boxes of the work item two rounds ahead,
// issued by the producer thread for short-K items (P.dbg bit 4; see launch_kmajor for when it is on)
__device__ __forceinline__ void tma_prefetch_4d(const void* map, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::tf32, fp32 accumulate)
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// same with the A operand read from tensor memory (128 lanes x K 32-bit columns) instead of shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = accumulator row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// 32 registers per thread -> 32 lanes x 32 consecutive columns (thread = TMEM lane)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
// 8 registers per thread -> 32 lanes x 8 consecutive columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, const float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, sm_100 version field = 1)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;      // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B (MN-major tf32)
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, tf32 x tf32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4)                       // c_format = F32
         | (2u << 7)                       // a_format = TF32
         | (2u << 10)                      // b_format = TF32
         | ((uint32_t)a_mn_major << 15)    // a_major: 0 = K, 1 = MN
         | ((uint32_t)b_mn_major << 16)
         | ((uint32_t)(N >> 3) << 17)
         | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------------
constexpr int kMaxTaps = 25;
constexpr int kThreads = 192;

struct __align__(64) TcMaps {
    CUtensorMap a[4];     // activation maps (one per input-pixel parity for stride 2)
    CUtensorMap b[4];     // K-major kernels: b[0] = filter map; wgrad: second activation operand maps
};

struct TcTap { signed char amap, ax, ay, bsel; };   // bsel: filter tap index (K-major kernels) / b map (wgrad)

// division by a runtime constant: q = umulhi(n, m), exact while n * d < 2^32 (checked on the host; m == 0 -> plain /)
struct FastDiv { uint32_t d, m; };
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv f) { return f.m ? __umulhi(n, f.m) : n / f.d; }

struct TcPhase {          // one GEMM problem (a dgrad parity phase, or the whole forward)
    int ntaps, kchunks;   // k iterations = ntaps * kchunks (32 channels each)
    TcTap taps[kMaxTaps];
    int ext_w, ext_h, ext_n;          // extents of the pixel grid this phase covers
    int tiles_w, tiles_h, tiles_n;
    FastDiv ftw, fth;                 // tiles_w, tiles_h
    long long out_off, sn, sh, sw;    // output element offset of pixel (n,h,w) = out_off + n*sn + h*sh + w*sw
                                      // (sn, sh, sw are the same in every phase; only out_off differs)
};

// Thin-channel forward (Ci <= 8: the image-side layers): one filter tap is 12-32 bytes of K, below a TMA box, so the
// conditioning warps GATHER their accumulator row's im2col slice straight from the NHWC input instead (K = the whole
// receptive field KH*KW*Ci, padded to a multiple of 32) -- no patch matrix is ever written.  Rows are consecutive
// output pixels q = (n*OH + oh)*OW + ow; column j = (kh*KW + kw)*Ci + ci.
constexpr int kGatherMaxK = 160;
struct TcGather {
    const float* x;
    int on;
    int H, W, C, OH, OW, KH, KW, stride, pad_t, pad_l, K;
    long long P;                      // output pixels
    FastDiv fow, foh;
};

struct TcParams {
    TcPhase ph[4];
    TcGather g;
    int nphases;
    int bw, bh, bn;       // pixel box: bw*bh*bn == 128
    int BN;               // channel tile (UMMA N)
    int ldn;              // number of output channels (columns)
    int mode;             // 1 = TF32 (operands rounded to nearest), 3 = 3xTF32 split
    int b_lo_tap_off;     // 3x: tap offset of the filter's lo copy inside the filter map
    int dbg;              // timing experiments only: bit0 skip the B_lo load, bit1 skip the conditioning pass
    int ksplit;           // CTAs sharing one output tile, each taking a slice of the (tap, k-chunk) loop (red.add epilogue)
    int grid_x, grid_y;   // work items: grid_x pixel tiles x grid_y channel tiles x (nphases * ksplit)
    FastDiv fgx, fgy, fks;
    const float* bias;
    float* out;
    int epi;              // fused epilogue (EG_EPI_*): out = act(v) / out = v * act'(mask[same offset as out]); only the
    float epi_neg;        // sign-type activations are fused: value (or derivative) = x (1) on the positive side,
    int epi_ge;           // epi_neg * x (epi_neg) on the other; epi_ge: zero counts as positive (tf.maximum(x, 0.2x))
    const float* mask;
};

// ---------------------------------------------------------------------------------------------------
// forward / input-gradient kernel (both operands K-major)
// ---------------------------------------------------------------------------------------------------
// One persistent kernel, one CTA per SM, looping over (pixel tile, channel tile, phase x k-split) work items.
//
//  * A operand through tensor memory (TS-mode MMA): an SS-mode 128x128x8 tf32 MMA streams 8 KB of operands per 64
//    cycles = the whole 128 B/clk shared-memory port, so TMA fills and any smem-side conditioning starve it.  Warps
//    2-5 read the landed A tile once (ld.shared.v4 at the swizzled positions of their accumulator row) and write it
//    to TMEM (tcgen05.st): mode 3 writes hi = trunc(x) and lo = rna(x - hi), mode 1 writes rna(x).
//  * chunked accumulation: the tensor core accumulates with truncation (tools/tc_probe.py: the error of a length-K
//    chain grows ~K * 2^-24 with a sign bias), so every kChunkStages stages the MMA warp switches between two TMEM
//    accumulators and warps 2-5 drain the finished one into fp32 registers with round-to-nearest adds, one chunk
//    behind the MMAs.  The same lag carries across work items: the output tile of item i is stored while the tensor
//    core is already multiplying item i+1 (no per-tile prologue / epilogue bubble).
constexpr int kChunkStages = 4;

struct WorkItem {
    int w0, h0, n0, col0, split, it0, niter, pz;
};

__device__ __forceinline__ bool get_item(const TcParams& P, int id, WorkItem& t) {
    const int yz = (int)fd_div((uint32_t)id, P.fgx), x = id - yz * P.grid_x;
    const int z = (int)fd_div((uint32_t)yz, P.fgy), y = yz - z * P.grid_y;
    t.pz = (int)fd_div((uint32_t)z, P.fks); t.split = z - t.pz * P.ksplit;
    const TcPhase& ph = P.ph[t.pz];
    if (x >= ph.tiles_w * ph.tiles_h * ph.tiles_n) return false;
    const int total_iter = ph.ntaps * ph.kchunks;
    int per_split = total_iter;
    if (P.ksplit > 1) per_split = (total_iter + P.ksplit - 1) / P.ksplit;
    t.it0 = t.split * per_split;
    t.niter = min(total_iter, t.it0 + per_split) - t.it0;
    if (t.niter <= 0) return false;
    const int r = (int)fd_div((uint32_t)x, ph.ftw), tw = x - r * ph.tiles_w;
    const int tn = (int)fd_div((uint32_t)r, ph.fth), th = r - tn * ph.tiles_h;
    t.w0 = tw * P.bw; t.h0 = th * P.bh; t.n0 = tn * P.bn; t.col0 = y * P.BN;
    return true;
}

// col2im scatter (TcGather.on == 2): the accumulator row of output pixel (n, oh, ow) holds dy . w^T for every im2col column
// j = (kh, kw, ci); each valid one is ADDED to dx[n, oh*s - pad + kh, ow*s - pad + kw, ci] (red.global.add; with 8 channels
// per tap two 16-byte reductions per tap).  Replaces the [P, Npad] product matrix round trip + col2im pass of the thin
// layers' input gradient; dx is initialised (bias or zero) by the caller.
struct ScatterCtx { float* base; uint32_t hmask, wmask; };
__device__ __forceinline__ void drain_scatter_group(const TcParams& P, const ScatterCtx& sc, const int* gtab, const float (&a32)[32],
                                                    const int c) {
    if (sc.base == nullptr) return;
    if (P.g.C == 8) {
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {
            const int tv = gtab[c + tp * 8];
            if (((sc.hmask >> (tv >> 24)) & (sc.wmask >> ((tv >> 16) & 255)) & 1u) != 0) {
                float* o = sc.base + (tv & 0xffff);
                red_add_v4(o, make_float4(a32[tp * 8], a32[tp * 8 + 1], a32[tp * 8 + 2], a32[tp * 8 + 3]));
                red_add_v4(o + 4, make_float4(a32[tp * 8 + 4], a32[tp * 8 + 5], a32[tp * 8 + 6], a32[tp * 8 + 7]));
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int tv = gtab[c + j];
            if (((sc.hmask >> (tv >> 24)) & (sc.wmask >> ((tv >> 16) & 255)) & 1u) != 0) atomicAdd(sc.base + (tv & 0xffff), a32[j]);
        }
    }
}

// One 32-column group of a finished accumulator tile: registers -> per-warp staging tile (XOR-swizzled transpose) -> global.
// A thread owns one accumulator row, so direct stores would touch 32 different 128-byte lines per instruction; after the
// transpose each instruction writes 4 rows x 128 contiguous bytes.  `plain`: full tile, no mask / split (bias add and the
// sign-type activation folded in); otherwise the general path (ragged tiles, activation-derivative mask, k-split red.add).
struct DrainCtx {
    uint32_t stg, vmask, pos;
    int lane, sub, cj;
    float* obase;
    const float* brow;
    bool plain;
};
// EG_EPI_MASK: sign bits (4 per stored row) of the mask values under one 32-column group of this lane's 8 output rows.
// Called for every group of an item BEFORE the drain warp waits for the accumulator, so that the DRAM latency of the
// mask lies under the item's MMAs instead of between the chunk drains (where it stalled the tensor pipe: both
// accumulator buffers filled up while the drain warps waited for the previous item's mask).
__device__ __forceinline__ uint32_t drain_mask_bits(const TcParams& P, const float* mbase, const long long (&loff)[8],
                                                    uint32_t vmask, int c) {
    float4 mk[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        mk[i] = (vmask & (1u << i)) ? __ldg(reinterpret_cast<const float4*>(mbase + loff[i] + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t pos = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 m = mk[i];
        const bool px = P.epi_ge ? m.x >= 0.f : m.x > 0.f, py = P.epi_ge ? m.y >= 0.f : m.y > 0.f;
        const bool pz = P.epi_ge ? m.z >= 0.f : m.z > 0.f, pw = P.epi_ge ? m.w >= 0.f : m.w > 0.f;
        pos |= ((uint32_t)px | ((uint32_t)py << 1) | ((uint32_t)pz << 2) | ((uint32_t)pw << 3)) << (4 * i);
    }
    return pos;
}

__device__ __forceinline__ void drain_store_group(const TcParams& P, const DrainCtx& d, const long long (&loff)[8],
                                                  const float (&a32)[32], const int c) {
    const uint32_t stg = d.stg, vmask = d.vmask;
    const int lane = d.lane, sub = d.sub, cj = d.cj;
    float* const obase = d.obase;
    const float* const brow = d.brow;
    const bool plain = d.plain;
    {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sts128(stg + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4),
                   make_uint4(__float_as_uint(a32[4 * j]), __float_as_uint(a32[4 * j + 1]),
                              __float_as_uint(a32[4 * j + 2]), __float_as_uint(a32[4 * j + 3])));
        }
        __syncwarp();
        if (plain) {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (brow != nullptr) bb = __ldg(reinterpret_cast<const float4*>(brow + c));
            // (the two tie rules of the sign-type activations give the same VALUE at zero: +-0)
            const float thr = 0.f;
            if (P.epi == EG_EPI_ACT) {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + sub;
                    const uint4 u = lds128(stg + (uint32_t)rr * 128u + (uint32_t)((cj ^ (rr & 7)) << 4));
                    float4 v = make_float4(__uint_as_float(u.x) + bb.x, __uint_as_float(u.y) + bb.y,
                                           __uint_as_float(u.z) + bb.z, __uint_as_float(u.w) + bb.w);
                    v.x = v.x > thr ? v.x : P.epi_neg * v.x; v.y = v.y > thr ? v.y : P.epi_neg * v.y;
                    v.z = v.z > thr ? v.z : P.epi_neg * v.z; v.w = v.w > thr ? v.w : P.epi_neg * v.w;
                    *reinterpret_cast<float4*>(obase + loff[i] + c) = v;
                }
            } else if (P.epi == EG_EPI_MASK) {
                const float ng = P.epi_neg;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + sub;
                    const uint4 u = lds128(stg + (uint32_t)rr * 128u + (uint32_t)((cj ^ (rr & 7)) << 4));
                    const uint32_t b = d.pos >> (4 * i);
                    *reinterpret_cast<float4*>(obase + loff[i] + c) =
                        make_float4((__uint_as_float(u.x) + bb.x) * ((b & 1u) ? 1.f : ng), (__uint_as_float(u.y) + bb.y) * ((b & 2u) ? 1.f : ng),
                                    (__uint_as_float(u.z) + bb.z) * ((b & 4u) ? 1.f : ng), (__uint_as_float(u.w) + bb.w) * ((b & 8u) ? 1.f : ng));
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + sub;
                    const uint4 u = lds128(stg + (uint32_t)rr * 128u + (uint32_t)((cj ^ (rr & 7)) << 4));
                    *reinterpret_cast<float4*>(obase + loff[i] + c) =
                        make_float4(__uint_as_float(u.x) + bb.x, __uint_as_float(u.y) + bb.y,
                                    __uint_as_float(u.z) + bb.z, __uint_as_float(u.w) + bb.w);
                }
            }
        } else {
            float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
            if (brow != nullptr) bb = __ldg(reinterpret_cast<const float4*>(brow + c));
            const uint32_t pos = d.pos;          // sign bits of the activation-derivative mask (drain_mask_bits, loaded early)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = i * 4 + sub;
                const uint4 u = lds128(stg + (uint32_t)rr * 128u + (uint32_t)((cj ^ (rr & 7)) << 4));
                if (vmask & (1u << i)) {
                    float* dst = obase + loff[i] + c;
                    float4 v = make_float4(__uint_as_float(u.x) + bb.x, __uint_as_float(u.y) + bb.y,
                                           __uint_as_float(u.z) + bb.z, __uint_as_float(u.w) + bb.w);
                    if (P.epi == EG_EPI_ACT) {
                        const float ng = P.epi_neg;
                        if (P.epi_ge) {
                            v.x = v.x >= 0.f ? v.x : ng * v.x; v.y = v.y >= 0.f ? v.y : ng * v.y;
                            v.z = v.z >= 0.f ? v.z : ng * v.z; v.w = v.w >= 0.f ? v.w : ng * v.w;
                        } else {
                            v.x = v.x > 0.f ? v.x : ng * v.x; v.y = v.y > 0.f ? v.y : ng * v.y;
                            v.z = v.z > 0.f ? v.z : ng * v.z; v.w = v.w > 0.f ? v.w : ng * v.w;
                        }
                    } else if (P.epi == EG_EPI_MASK) {
                        const uint32_t b = pos >> (4 * i);
                        const float ng = P.epi_neg;
                        v.x *= (b & 1u) ? 1.f : ng; v.y *= (b & 2u) ? 1.f : ng;
                        v.z *= (b & 4u) ? 1.f : ng; v.w *= (b & 8u) ? 1.f : ng;
                    }
                    if (P.ksplit > 1) red_add_v4(dst, v);
                    else *reinterpret_cast<float4*>(dst) = v;
                }
            }
        }
        __syncwarp();
    }
}

// warpgroups 0 and 1 (warps 0-7): A -> TMEM, alternating stages (one group alone cannot condition a stage in the
// time the tensor core needs to consume it); warpgroup 2 (warps 8-11): drain + store; warp 12: TMA, warp 13: MMA.
// Roles are warpgroup-aligned so that setmaxnreg can move registers to the epilogue warps (128 accumulators each).
constexpr int kThreadsK = 512;

template <int kStages>
__global__ void __launch_bounds__(kThreadsK, 1)
conv_tc_kmajor(const __grid_constant__ TcMaps maps, const __grid_constant__ TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int BN = P.BN, mode = P.mode;
    const uint32_t a_bytes = 128 * 128, b_bytes = (uint32_t)BN * 128;
    // stage: [A landing tile][B][B_lo (3x)]
    const uint32_t b_off = a_bytes, b_lo_off = b_off + b_bytes;
    const uint32_t stage_bytes = a_bytes + (mode == 3 ? 2 : 1) * b_bytes;
    const uint32_t tx_bytes = P.g.on == 1 ? stage_bytes - a_bytes : stage_bytes;      // gather mode: only the filter arrives by TMA
    const uint32_t stg_base = smem_base + kStages * stage_bytes;   // 4 x 4 KB output staging tiles (one per epilogue warp)

    __shared__ __align__(8) uint64_t full_bar[kStages];      // TMA bytes landed
    __shared__ __align__(8) uint64_t ready_bar[kStages];     // A operand written to tensor memory (4 warp arrivals)
    __shared__ __align__(8) uint64_t empty_bar[kStages];     // MMAs that read the stage have completed
    __shared__ __align__(8) uint64_t acc_full_bar[2];        // chunk accumulator complete (tcgen05.commit)
    __shared__ __align__(8) uint64_t acc_empty_bar[2];       // chunk accumulator drained (4 warp arrivals)
    __shared__ uint32_t tmem_base_slot;
    __shared__ int gtab[kGatherMaxK];                        // gather mode: column j -> kh << 24 | kw << 16 | element offset

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (P.g.on) {
        for (int j = threadIdx.x; j < kGatherMaxK; j += blockDim.x) {
            int tv = 31 << 24;                               // row 31 never exists: padding columns read as zero
            if (j < P.g.K) {
                const int tap = j / P.g.C, ci = j - tap * P.g.C, kh = tap / P.g.KW, kw = tap - kh * P.g.KW;
                tv = (kh << 24) | (kw << 16) | ((kh * P.g.W + kw) * P.g.C + ci);
            }
            gtab[j] = tv;
        }
    }
    // TMEM columns: [0, 2*BN) two chunk accumulators, then kStages x (32 hi + 32 lo) A columns
    const uint32_t a_col0 = (uint32_t)(2 * BN);
    const int need_cols = 2 * BN + kStages * 64;
    const uint32_t tmem_cols = need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512);
    const int total_ids = P.grid_x * P.grid_y * P.nphases * P.ksplit;

    if (warp == 12 && lane == 0) {
        for (int i = 0; i < 4; ++i) prefetch_tmap(&maps.a[i]);
        prefetch_tmap(&maps.b[0]);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&ready_bar[s]), 4); mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&acc_full_bar[b]), 1); mbar_init(smem_u32(&acc_empty_bar[b]), 4); }
        fence_barrier_init();
    }
    if (warp == 13) { tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp >= 12) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
      if (warp == 12) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            WorkItem t;
            if (P.dbg & 8) __nanosleep((blockIdx.x & 15) * 500);
            for (int id = blockIdx.x; id < total_ids; id += gridDim.x) {
                if (!get_item(P, id, t)) continue;
                if (P.dbg & 16) {   // L2 prefetch of the A operand of the item two rounds ahead
                    WorkItem tn;
                    const int idn = id + 2 * (int)gridDim.x;
                    if (idn < total_ids && get_item(P, idn, tn) && tn.niter <= 6) {
                        if (P.g.on == 1) {
                            // gather mode: the input rows the 128 output pixels of that item read are one contiguous range
                            const long long p0 = tn.n0, p1 = min((long long)tn.n0 + 127, P.g.P - 1);
                            const long long r0 = p0 / P.g.OW, r1 = p1 / P.g.OW;           // global output rows n*OH + oh
                            const long long n0i = r0 / P.g.OH, n1i = r1 / P.g.OH;
                            long long h0 = (r0 - n0i * P.g.OH) * P.g.stride - P.g.pad_t, h1 = (r1 - n1i * P.g.OH) * P.g.stride - P.g.pad_t + P.g.KH;
                            if (h0 < 0) h0 = 0;
                            if (h1 > P.g.H) h1 = P.g.H;
                            const long long e0 = (n0i * P.g.H + h0) * P.g.W * P.g.C, e1 = (n1i * P.g.H + h1) * P.g.W * P.g.C;
                            const uintptr_t a0 = reinterpret_cast<uintptr_t>(P.g.x + e0) & ~(uintptr_t)15;
                            const uintptr_t a1 = (reinterpret_cast<uintptr_t>(P.g.x + e1)) & ~(uintptr_t)15;
                            if (a1 > a0) l2_prefetch_bulk(reinterpret_cast<const void*>(a0), (uint32_t)min((uintptr_t)(256 * 1024), a1 - a0));
                        } else {
                            const TcPhase& phn = P.ph[tn.pz];
                            int tapn = tn.it0 / phn.kchunks, kcn = tn.it0 - tapn * phn.kchunks;
                            for (int it = 0; it < tn.niter; ++it) {
                                const TcTap tq = phn.taps[tapn];
                                tma_prefetch_4d(&maps.a[tq.amap], kcn * 32, tn.w0 + tq.ax, tn.h0 + tq.ay, tn.n0);
                                if (++kcn == phn.kchunks) { kcn = 0; ++tapn; }
                            }
                        }
                    }
                }
                const TcPhase& ph = P.ph[t.pz];
                int tap = t.it0 / ph.kchunks, kc = t.it0 - tap * ph.kchunks;      // advanced incrementally: the single
                TcTap tp = ph.taps[tap];                                           // producer thread is latency-critical
                for (int it = 0; it < t.niter; ++it) {
                    mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    mbar_expect_tx(fb, tx_bytes);
                    if (P.g.on != 1) tma_load_4d(sa, &maps.a[tp.amap], fb, kc * 32, t.w0 + tp.ax, t.h0 + tp.ay, t.n0);
                    tma_load_3d(sa + b_off, &maps.b[0], fb, kc * 32, t.col0, tp.bsel);
                    if (mode == 3) tma_load_3d(sa + b_lo_off, &maps.b[0], fb, kc * 32, t.col0, tp.bsel + P.b_lo_tap_off);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                    if (++kc == ph.kchunks) { kc = 0; ++tap; if (it + 1 < t.niter) tp = ph.taps[tap]; }
                }
            }
        }
      } else if (warp == 13) {
        // ===== MMA issuer =====
        if (elect_one()) {
            const uint32_t idesc = make_idesc(128, BN, 0, 0);
            int stage = 0; uint32_t phase = 0;
            int cg = 0;                                          // chunks issued so far (over all work items)
            WorkItem t;
            for (int id = blockIdx.x; id < total_ids; id += gridDim.x) {
                if (!get_item(P, id, t)) continue;
                for (int it = 0; it < t.niter; ++it) {
                    const bool first = (it % kChunkStages == 0);
                    const int buf = cg & 1;
                    if (first && cg >= 2) { mbar_wait(smem_u32(&acc_empty_bar[buf]), (uint32_t)(((cg >> 1) - 1) & 1)); tc_fence_after(); }
                    const uint32_t dst = tmem_base + (uint32_t)(buf * BN);
                    mbar_wait(smem_u32(&ready_bar[stage]), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * stage_bytes;
                    const uint64_t bd = make_smem_desc(sa + b_off, 16, 1024);
                    const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)(stage * 64), a_lo = a_hi + 32;
                    if (mode == 3) {
                        const uint64_t bld = make_smem_desc(sa + b_lo_off, 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {            // 4 x (K = 8 tf32 = 32 B) inside the 128-byte swizzle span
                            umma_tf32_ts(dst, a_hi + k * 8, bd + (uint64_t)(k * 2), idesc, !(first && k == 0));
                            umma_tf32_ts(dst, a_lo + k * 8, bd + (uint64_t)(k * 2), idesc, 1);
                            umma_tf32_ts(dst, a_hi + k * 8, bld + (uint64_t)(k * 2), idesc, 1);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_tf32_ts(dst, a_hi + k * 8, bd + (uint64_t)(k * 2), idesc, !(first && k == 0));
                    }
                    umma_commit(smem_u32(&empty_bar[stage]));
                    if (it % kChunkStages == kChunkStages - 1 || it == t.niter - 1) {
                        umma_commit(smem_u32(&acc_full_bar[buf]));
                        ++cg;
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
      }
    } else if (warp < 8) {
        // ===== warps 0..7: A operand -> tensor memory (warps 0-3 take the even stages, warps 4-7 the odd ones) =====
        asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
        static_assert(kStages % 2 == 0, "stage parity selects the conditioning warpgroup");
        const int q = warp & 3;                                  // TMEM lane quarter this warp may access
        const int grp = warp >> 2;
        const int arow = q * 32 + lane;                          // A tile row of this thread = TMEM lane
        int stage = 0; uint32_t phase = 0;
        WorkItem t;
        for (int id = blockIdx.x; id < total_ids; id += gridDim.x) {
            if (!get_item(P, id, t)) continue;
            const int niter = t.niter;
            // gather mode: this thread's output pixel and the validity masks of its KH x KW window
            const float* gbase = nullptr;
            uint32_t hmask = 0, wmask = 0;
            if (P.g.on == 1) {
                const long long pix = (long long)t.n0 + arow;
                if (pix < P.g.P) {
                    const uint32_t r1 = fd_div((uint32_t)pix, P.g.fow), ow = (uint32_t)pix - r1 * (uint32_t)P.g.OW;
                    const uint32_t n = fd_div(r1, P.g.foh), oh = r1 - n * (uint32_t)P.g.OH;
                    const int h0 = (int)oh * P.g.stride - P.g.pad_t, w0 = (int)ow * P.g.stride - P.g.pad_l;
                    for (int k = 0; k < P.g.KH; ++k) hmask |= (uint32_t)(h0 + k >= 0 && h0 + k < P.g.H) << k;
                    for (int k = 0; k < P.g.KW; ++k) wmask |= (uint32_t)(w0 + k >= 0 && w0 + k < P.g.W) << k;
                    gbase = P.g.x + (((long long)n * P.g.H + h0) * P.g.W + w0) * P.g.C;
                }
            }
            for (int it = 0; it < niter; ++it) {
                if ((stage & 1) != grp) { if (++stage == kStages) { stage = 0; phase ^= 1; } continue; }
                const uint32_t sa = smem_base + stage * stage_bytes;
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a_col0 + (uint32_t)(stage * 64);
                if (P.g.on == 1) {
                    // Gather this thread's row of the stage (32 im2col columns) into ITS OWN row of the stage's shared-memory
                    // A tile (no other thread touches that row, so no barrier is involved), then take the common path.
                    // Deliberately compact, rolled loops: ncu's source view of the fully unrolled version (32 loads, ~830
                    // instructions per stage) showed 32 % of all stall samples as `no_inst` -- with four roles running
                    // different code the instruction cache, not the load latency, bounded the short-K thin layers.
                    const int j0 = (t.it0 + it) * 32;
                    const uint32_t rowaddr = sa + (uint32_t)arow * 128u, sw = (uint32_t)(arow & 7);
                    if (P.g.C == 8) {                        // a tap = 8 contiguous channels = two aligned 16-byte loads
#pragma unroll 1
                        for (int tp2 = 0; tp2 < 4; tp2 += 2) {
                            uint4 v[4];
#pragma unroll
                            for (int u = 0; u < 2; ++u) {
                                const int tv = gtab[j0 + (tp2 + u) * 8];
                                const bool ok = ((hmask >> (tv >> 24)) & (wmask >> ((tv >> 16) & 255)) & 1u) != 0;
                                v[2 * u] = make_uint4(0u, 0u, 0u, 0u); v[2 * u + 1] = v[2 * u];
                                if (ok) {
                                    const uint4* pp = reinterpret_cast<const uint4*>(gbase + (tv & 0xffff));
                                    v[2 * u] = __ldg(pp); v[2 * u + 1] = __ldg(pp + 1);
                                }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) sts128(rowaddr + ((((uint32_t)(2 * tp2 + u)) ^ sw) << 4), v[u]);
                        }
                    } else {
#pragma unroll 1
                        for (int e16 = 0; e16 < 32; e16 += 16) {
                            uint32_t v[16];
#pragma unroll
                            for (int u = 0; u < 16; ++u) {
                                const int tv = gtab[j0 + e16 + u];
                                const bool ok = ((hmask >> (tv >> 24)) & (wmask >> ((tv >> 16) & 255)) & 1u) != 0;
                                v[u] = ok ? __float_as_uint(__ldg(gbase + (tv & 0xffff))) : 0u;
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                                sts128(rowaddr + ((((uint32_t)((e16 >> 2) + u)) ^ sw) << 4), make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]));
                        }
                    }
                }
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                // 8 K-columns (two 16-byte chunks at the swizzled positions of row arow) per iteration of a ROLLED loop:
                // the whole conditioning pass is ~70 instructions of code instead of ~300 straight-line ones (see the note
                // on the instruction cache above; tcgen05.st moves 32 lanes x 8 columns per instruction here)
                if (!(P.dbg & 2)) {
                    const uint32_t rowaddr = sa + (uint32_t)arow * 128u, sw = (uint32_t)(arow & 7);
#pragma unroll 1
                    for (int qd = 0; qd < 4; ++qd) {
                        const uint4 v0 = lds128(rowaddr + ((((uint32_t)(2 * qd)) ^ sw) << 4));
                        const uint4 v1 = lds128(rowaddr + ((((uint32_t)(2 * qd + 1)) ^ sw) << 4));
                        uint32_t h8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                        if (mode == 3) {
                            uint32_t l8[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) { l8[j] = tf32_lo(h8[j]); h8[j] &= 0xFFFFE000u; }
                            tmem_st8(ta + (uint32_t)(8 * qd), h8);
                            tmem_st8(ta + 32u + (uint32_t)(8 * qd), l8);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) h8[j] = tf32_rna(h8[j]);
                            tmem_st8(ta + (uint32_t)(8 * qd), h8);
                        }
                    }
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ready_bar[stage]));
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== warps 8..11: drain the chunk accumulators into fp32 registers, store the finished tiles =====
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int q = warp & 3;                                  // this thread's accumulator row = TMEM lane q * 32 + lane
        float acc[128];
        // the 8 tile rows this lane stores (pass i: row q*32 + i*4 + lane/8): local pixel coordinates and output offset
        const int sub = lane >> 3, cj = lane & 7;                // row within a group of 4, 16-byte chunk within the row
        uint32_t loc[8];
        long long loff[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = q * 32 + i * 4 + sub;
            const int wl = r % P.bw, hl = (r / P.bw) % P.bh, nl = r / (P.bw * P.bh);
            loc[i] = (uint32_t)wl | ((uint32_t)hl << 8) | ((uint32_t)nl << 16);
            loff[i] = (long long)nl * P.ph[0].sn + (long long)hl * P.ph[0].sh + (long long)wl * P.ph[0].sw + cj * 4;
        }
        const uint32_t stg = stg_base + (uint32_t)q * 4096u;
        int cg = 0;                                              // chunks drained so far (over all work items)
        WorkItem t;
        for (int id = blockIdx.x; id < total_ids; id += gridDim.x) {
            if (!get_item(P, id, t)) continue;
            // Store through a per-warp staging tile (32 rows x 32 columns, 16-byte chunks XOR-swizzled by row): a thread
            // owns one accumulator row, so direct stores would touch 32 different 128-byte lines per instruction; after
            // the transpose each instruction writes 4 rows x 128 contiguous bytes.
            const TcPhase& ph = P.ph[t.pz];
            uint32_t vmask = 0xffu;
            if (t.w0 + P.bw > ph.ext_w || t.h0 + P.bh > ph.ext_h || t.n0 + P.bn > ph.ext_n) {   // ragged tile
                vmask = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int ow = t.w0 + (int)(loc[i] & 255u), oh = t.h0 + (int)((loc[i] >> 8) & 255u), on = t.n0 + (int)(loc[i] >> 16);
                    if (ow < ph.ext_w && oh < ph.ext_h && on < ph.ext_n) vmask |= 1u << i;
                }
            }
            if (P.dbg & 4) vmask = 0;
            float* obase = P.out + ph.out_off + (long long)t.n0 * ph.sn + (long long)t.h0 * ph.sh + (long long)t.w0 * ph.sw + t.col0;
            const float* brow = (P.bias != nullptr && t.split == 0) ? P.bias + t.col0 + cj * 4 : nullptr;
            // full tile, no mask / split: the transposed copy with the bias add and the sign-type activation folded in
            const bool plain = vmask == 0xffu && P.ksplit == 1;
            DrainCtx dc;
            dc.stg = stg; dc.vmask = vmask; dc.pos = 0; dc.lane = lane; dc.sub = sub; dc.cj = cj; dc.obase = obase; dc.brow = brow; dc.plain = plain;
            uint32_t mpos0 = 0, mpos1 = 0, mpos2 = 0, mpos3 = 0;
            if (P.epi == EG_EPI_MASK) {
                const float* mbase = P.mask + (obase - P.out);
                mpos0 = drain_mask_bits(P, mbase, loff, vmask, 0);
                if (BN > 32) mpos1 = drain_mask_bits(P, mbase, loff, vmask, 32);
                if (BN > 64) { mpos2 = drain_mask_bits(P, mbase, loff, vmask, 64); mpos3 = drain_mask_bits(P, mbase, loff, vmask, 96); }
            }
            ScatterCtx sc;
            sc.base = nullptr; sc.hmask = 0; sc.wmask = 0;
            if (P.g.on == 2) {                               // this thread's accumulator row = output pixel t.n0 + q*32 + lane
                const long long pix = (long long)t.n0 + q * 32 + lane;
                if (pix < P.g.P && !(P.dbg & 4)) {
                    const uint32_t r1 = fd_div((uint32_t)pix, P.g.fow), ow = (uint32_t)pix - r1 * (uint32_t)P.g.OW;
                    const uint32_t n = fd_div(r1, P.g.foh), oh = r1 - n * (uint32_t)P.g.OH;
                    const int h0 = (int)oh * P.g.stride - P.g.pad_t, w0 = (int)ow * P.g.stride - P.g.pad_l;
                    for (int k = 0; k < P.g.KH; ++k) sc.hmask |= (uint32_t)(h0 + k >= 0 && h0 + k < P.g.H) << k;
                    for (int k = 0; k < P.g.KW; ++k) sc.wmask |= (uint32_t)(w0 + k >= 0 && w0 + k < P.g.W) << k;
                    sc.base = P.out + (((long long)n * P.g.H + h0) * P.g.W + w0) * P.g.C;
                }
            }
            const int nchunks = (t.niter + kChunkStages - 1) / kChunkStages;
            if (nchunks == 1) {
                // Short-K item (one chunk): nothing accumulates across chunks, so the groups go TMEM -> staging -> global one
                // at a time in a ROLLED loop -- ~150 instructions of code per item instead of ~600 straight-line ones.
                // These items are instruction-issue-bound (icc hit rate 60 %, 1.4 IPC in the ncu capture of the thin layers).
                const int buf = cg & 1;
                mbar_wait(smem_u32(&acc_full_bar[buf]), (uint32_t)((cg >> 1) & 1));
                tc_fence_after();
                const int ngroups = BN >> 5;
#pragma unroll 1
                for (int g = 0; g < ngroups; ++g) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + g * 32), r);
                    tmem_ld_wait();
                    if (g == ngroups - 1) {                  // the accumulator is free as soon as its last columns are read
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&acc_empty_bar[buf]));
                    }
                    dc.pos = g == 0 ? mpos0 : (g == 1 ? mpos1 : (g == 2 ? mpos2 : mpos3));
                    if (P.g.on == 2) { if (g * 32 < P.g.K) drain_scatter_group(P, sc, gtab, reinterpret_cast<const float (&)[32]>(r), g * 32); }
                    else drain_store_group(P, dc, loff, reinterpret_cast<const float (&)[32]>(r), g * 32);
                }
                ++cg;
                continue;
            }
#pragma unroll 1
            for (int c = 0; c < nchunks; ++c, ++cg) {
                const int buf = cg & 1;
                mbar_wait(smem_u32(&acc_full_bar[buf]), (uint32_t)((cg >> 1) & 1));
                tc_fence_after();
                // the first chunk of an item ASSIGNS (no zeroing of 128 registers per item, no add): short-K items are
                // bound by the instruction count of these four warps (tools/gather_time.py)
                if (c == 0) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g * 32 < BN) {
                            uint32_t r[32];
                            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + g * 32), r);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) acc[g * 32 + j] = __uint_as_float(r[j]);
                        }
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        if (g * 32 < BN) {
                            uint32_t r[32];
                            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + g * 32), r);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) acc[g * 32 + j] += __uint_as_float(r[j]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&acc_empty_bar[buf]));
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (g * 32 < BN) {
                    dc.pos = g == 0 ? mpos0 : (g == 1 ? mpos1 : (g == 2 ? mpos2 : mpos3));
                    if (P.g.on == 2) { if (g * 32 < P.g.K) drain_scatter_group(P, sc, gtab, reinterpret_cast<const float (&)[32]>(acc[g * 32]), g * 32); }
                    else drain_store_group(P, dc, loff, reinterpret_cast<const float (&)[32]>(acc[g * 32]), g * 32);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 13) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------
// filter-gradient kernel (both operands MN-major), split over pixels
// ---------------------------------------------------------------------------------------------------
struct WgParams {
    int ntaps;
    TcTap taps[kMaxTaps];             // amap/ax/ay: tap offset applied to operand X; bsel unused
    int x_is_a;                       // 1: A (rows) = x channels, B (cols) = dy channels ; 0: swapped
    int bw, bh, bn, pix;              // pixel box, pix = bw*bh*bn (64; 32 in 3x mode)
    int tiles_w, tiles_h, tiles_n;    // pixel tiles
    int chunks_per_split;             // pixel tiles per CTA
    int BN;                           // column tile
    long long tap_stride, sm, sn;     // dw element = tap*tap_stride + m*sm + n*sn
    int rows_total, cols_total;       // valid rows / columns (channels)
    int tap_group;                    // taps packed side by side into the column tile (x on the column side, cols <= 64):
                                      // column c of the tile = tap (group*tap_group + c / cols_total), channel c % cols_total
    int layout_type, sbo;             // UMMA smem descriptor layout type / stride-byte-offset
    int mode;                         // 1 = TF32 rounded, 3 = 3xTF32
    float* out;
    TcGather g;                       // thin-channel filter gradient: the row operand (im2col columns of x) is gathered
    int gN;                           //   from the NHWC input by the row-operand warps (kTS kernel only); gN = batch
};

// kTS (3xTF32 mode): the row operand goes through tensor memory like in the K-major kernel.  Thread r of warps 2-5
// owns channel row r of the 128-row tile: for each of the 32 pixels of a stage it reads its channel from the
// MN-major slab (one 128-B line per pixel, 32-B-atom swizzle -> conflict-free across the warp) and stores hi / lo as
// 32 TMEM columns; the column operand stays in shared memory and gets its lo copy written next to it.
// In the kTS kernel the per-stage work of the helper warps is split: warps 2-5 move the row operand to tensor memory,
// warps 6-9 write the lo copy of the column operand (a single warp per scheduler issues ~one dependent instruction
// every 5 cycles).  A second row-operand warpgroup on alternate stages was measured and did not help.
constexpr int kThreadsW3 = 320;

template <int kStages, bool kTS>
__global__ void __launch_bounds__(kTS ? kThreadsW3 : kThreads, (kTS && kStages == 2) ? 2 : 1)
conv_tc_wgrad(const __grid_constant__ TcMaps maps, const __grid_constant__ WgParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int BN = P.BN, PIX = P.pix, mode = P.mode;
    const uint32_t sub_bytes = (uint32_t)PIX * 128;          // one 32-channel slab: PIX rows x 128 B
    const uint32_t a_bytes = 4 * sub_bytes, b_bytes = (uint32_t)(BN / 32) * sub_bytes;
    const uint32_t ab_bytes = a_bytes + b_bytes;
    // stage: SS: [A slabs][B slabs][A_lo][B_lo] (lo: 3x only) ; TS: [A slabs][B slabs][B_lo]
    const uint32_t stage_bytes = kTS ? ab_bytes + b_bytes : (mode == 3 ? 2 : 1) * ab_bytes;
    const uint32_t b_lo_off = kTS ? ab_bytes : ab_bytes + a_bytes;

    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ __align__(8) uint64_t ready_bar[kStages];
    __shared__ __align__(8) uint64_t empty_bar[kStages];
    __shared__ __align__(8) uint64_t tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;
    __shared__ int ptab[32];                                 // gather mode: pixel p of the box -> wl | hl << 8 | nl << 16

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        const int p = threadIdx.x;
        ptab[p] = (p % P.bw) | (((p / P.bw) % P.bh) << 8) | ((p / (P.bw * P.bh)) << 16);
    }
    const int G = P.tap_group, ngroups = (P.ntaps + G - 1) / G;
    const int tap = (blockIdx.z % ngroups) * G, split = blockIdx.z / ngroups;     // first tap of this CTA's group
    const TcTap tp = P.taps[tap];
    const int row0 = blockIdx.x * 128, col0 = G > 1 ? 0 : blockIdx.y * BN;
    const int slabs_per_tap = G > 1 ? P.cols_total / 32 : BN / 32;
    const int ntiles = P.tiles_w * P.tiles_h * P.tiles_n;
    const int t_beg = split * P.chunks_per_split;
    const int t_end = min(ntiles, t_beg + P.chunks_per_split);
    const int niter = t_end - t_beg;
    const uint32_t a_col0 = (uint32_t)BN;
    const int need_cols = kTS ? BN + kStages * 64 : BN;
    const uint32_t tmem_cols = need_cols <= 32 ? 32 : (need_cols <= 64 ? 64 : (need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512)));
    if (niter <= 0) return;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < 4; ++i) { prefetch_tmap(&maps.a[i]); }
        prefetch_tmap(&maps.b[0]);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&ready_bar[s]), kTS ? 8 : 4); mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        mbar_init(smem_u32(&tmem_full_bar), 1);
        fence_barrier_init();
    }
    if (warp == 1) { tmem_alloc(smem_u32(&tmem_base_slot), tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (elect_one()) {
            // column slabs (x on the column side): slab j belongs to tap (tap + j / slabs_per_tap); a group past the
            // last tap re-reads the last one (its columns are masked in the epilogue)
            int s_amap[4], s_ax[4], s_ay[4], s_ch[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const TcTap tq = P.taps[min(tap + j / slabs_per_tap, P.ntaps - 1)];
                s_amap[j] = tq.amap; s_ax[j] = tq.ax; s_ay[j] = tq.ay; s_ch[j] = col0 + (j % slabs_per_tap) * 32;
            }
            int stage = 0; uint32_t phase = 0;
            int tw = t_beg % P.tiles_w, th = (t_beg / P.tiles_w) % P.tiles_h, tn = t_beg / (P.tiles_w * P.tiles_h);
            for (int it = 0; it < niter; ++it) {
                const int w0 = tw * P.bw, h0 = th * P.bh, n0 = tn * P.bn;
                if (++tw == P.tiles_w) { tw = 0; if (++th == P.tiles_h) { th = 0; ++tn; } }
                mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
                const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + a_bytes;
                const uint32_t fb = smem_u32(&full_bar[stage]);
                mbar_expect_tx(fb, P.g.on ? b_bytes : ab_bytes);
                // rows operand: 4 slabs of 32 channels; cols operand: BN/32 slabs
                for (int i = 0; i < 4 && !P.g.on; ++i) {
                    if (P.x_is_a) tma_load_4d(sa + i * sub_bytes, &maps.a[tp.amap], fb, row0 + i * 32, w0 + tp.ax, h0 + tp.ay, n0);
                    else          tma_load_4d(sa + i * sub_bytes, &maps.b[0], fb, row0 + i * 32, w0, h0, n0);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (j < BN / 32) {
                        if (P.x_is_a) tma_load_4d(sb + j * sub_bytes, &maps.b[0], fb, col0 + j * 32, w0, h0, n0);
                        else          tma_load_4d(sb + j * sub_bytes, &maps.a[s_amap[j]], fb, s_ch[j], w0 + s_ax[j], h0 + s_ay[j], n0);
                    }
                }
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc(128, BN, kTS ? 0 : 1, 1);
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < niter; ++it) {
                mbar_wait(smem_u32(&ready_bar[stage]), phase);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * stage_bytes, sb = sa + a_bytes;
                // MN-major tf32 operands must use the "128B swizzle with 32B atoms" layout (Swizzle<2,5,2>): an atom is
                // 32 channels (128 B) x 4 pixels; LBO = next 32-channel slab, SBO = next 4 pixels (512 B)
                const uint64_t bd = make_smem_desc(sb, sub_bytes, P.sbo, P.layout_type);
                if (kTS) {
                    const uint64_t bld = make_smem_desc(sa + b_lo_off, sub_bytes, P.sbo, P.layout_type);
                    const uint32_t a_hi = tmem_base + a_col0 + (uint32_t)(stage * 64), a_lo = a_hi + 32;
                    for (int k = 0; k < PIX / 8; ++k) {
                        umma_tf32_ts(tmem_base, a_hi + k * 8, bd + (uint64_t)(k * 64), idesc, (it | k) != 0);
                        umma_tf32_ts(tmem_base, a_lo + k * 8, bd + (uint64_t)(k * 64), idesc, 1);
                        umma_tf32_ts(tmem_base, a_hi + k * 8, bld + (uint64_t)(k * 64), idesc, 1);
                    }
                } else {
                    const uint64_t ad = make_smem_desc(sa, sub_bytes, P.sbo, P.layout_type);
                    for (int k = 0; k < PIX / 8; ++k)
                        umma_tf32(tmem_base, ad + (uint64_t)(k * 64), bd + (uint64_t)(k * 64), idesc, (it | k) != 0);
                    if (mode == 3) {
                        const uint64_t ald = make_smem_desc(sa + ab_bytes, sub_bytes, P.sbo, P.layout_type);
                        const uint64_t bld = make_smem_desc(sb + ab_bytes, sub_bytes, P.sbo, P.layout_type);
                        for (int k = 0; k < PIX / 8; ++k) umma_tf32(tmem_base, ald + (uint64_t)(k * 64), bd + (uint64_t)(k * 64), idesc, 1);
                        for (int k = 0; k < PIX / 8; ++k) umma_tf32(tmem_base, ad + (uint64_t)(k * 64), bld + (uint64_t)(k * 64), idesc, 1);
                    }
                }
                umma_commit(smem_u32(&empty_bar[stage]));
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            umma_commit(smem_u32(&tmem_full_bar));
        }
    } else if (kTS && warp >= 6) {
        // column operand: lo copy next to it in shared memory
        const int ctid = threadIdx.x - 192;
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < niter; ++it) {
            mbar_wait(smem_u32(&full_bar[stage]), phase);
            const uint32_t sa = smem_base + stage * stage_bytes;
            condition_tile(sa + a_bytes, sa + b_lo_off, b_bytes, ctid, 3);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ready_bar[stage]));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    } else {
        const int ctid = threadIdx.x - 64;
        const int q = warp & 3;
        if (kTS && P.g.on) {
            // thin-channel filter gradient: row r of the tile is im2col column j = (kh*KW + kw)*C + ci; its 32 values of a
            // stage are that tap of the 32 output pixels of the stage's pixel box, read straight from the NHWC input
            const int j = row0 + q * 32 + lane;
            const bool jv = j < P.g.K;
            int kh = 0, kw = 0, ci = 0;
            if (jv) { const int tp_ = j / P.g.C; ci = j - tp_ * P.g.C; kh = tp_ / P.g.KW; kw = tp_ - kh * P.g.KW; }
            const int dh = kh - P.g.pad_t, dw_ = kw - P.g.pad_l;
            int stage = 0; uint32_t phase = 0;
            int tw = t_beg % P.tiles_w, th = (t_beg / P.tiles_w) % P.tiles_h, tn = t_beg / (P.tiles_w * P.tiles_h);
            for (int it = 0; it < niter; ++it) {
                const int w0 = tw * P.bw, h0 = th * P.bh, n0 = tn * P.bn;
                if (++tw == P.tiles_w) { tw = 0; if (++th == P.tiles_h) { th = 0; ++tn; } }
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int p = 0; p < 32; ++p) {
                    const int pt = ptab[p];                  // wl | hl << 8 | nl << 16
                    const int n = n0 + (pt >> 16), h = (h0 + ((pt >> 8) & 255)) * P.g.stride + dh, w = (w0 + (pt & 255)) * P.g.stride + dw_;
                    const bool ok = jv && n < P.gN && h >= 0 && h < P.g.H && w >= 0 && w < P.g.W;
                    hi[p] = ok ? __float_as_uint(__ldg(P.g.x + (((long long)n * P.g.H + h) * P.g.W + w) * P.g.C + ci)) : 0u;
                }
                mbar_wait(smem_u32(&full_bar[stage]), phase);  // the TMEM slot is free once the stage has been refilled
#pragma unroll
                for (int p = 0; p < 32; ++p) { lo[p] = tf32_lo(hi[p]); hi[p] &= 0xFFFFE000u; }
                const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a_col0 + (uint32_t)(stage * 64);
                tmem_st32(ta, hi);
                tmem_st32(ta + 32, lo);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ready_bar[stage]));
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        } else {
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < niter; ++it) {
                mbar_wait(smem_u32(&full_bar[stage]), phase);
                const uint32_t sa = smem_base + stage * stage_bytes;
                if (kTS) {
                    // row operand: channel (q*32 + lane) of slab q, pixels 0..31 -> 32 hi + 32 lo TMEM columns
                    uint32_t hi[32], lo[32];
                    const uint32_t slab = sa + (uint32_t)q * sub_bytes + (uint32_t)((lane & 7) << 2);
#pragma unroll
                    for (int p = 0; p < 32; ++p)
                        hi[p] = lds32(slab + (uint32_t)p * 128u + (uint32_t)((((lane >> 3) ^ (p & 3)) << 5)));
#pragma unroll
                    for (int p = 0; p < 32; ++p) { lo[p] = tf32_lo(hi[p]); hi[p] &= 0xFFFFE000u; }
                    const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + a_col0 + (uint32_t)(stage * 64);
                    tmem_st32(ta, hi);
                    tmem_st32(ta + 32, lo);
                    tmem_st_wait();
                    tc_fence_before();
                } else {
                    condition_tile(sa, sa + ab_bytes, ab_bytes, ctid, mode);       // both operands are activations
                    fence_proxy_async();
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&ready_bar[stage]));
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        const int row = row0 + q * 32 + lane;
        const bool valid = row < P.rows_total;
        float* obase = P.out + (long long)row * P.sm;
        mbar_wait(smem_u32(&tmem_full_bar), 0);
        tc_fence_after();
        for (int c = 0; c < BN; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
            tmem_ld_wait();
            if (valid) {
                // 32 consecutive columns never straddle two taps (cols_total is a multiple of 32 when taps are packed)
                const int tj = G > 1 ? tap + c / P.cols_total : tap;
                const int cbase = G > 1 ? c % P.cols_total : col0 + c;
                float* ob = obase + (long long)tj * P.tap_stride;
                if (tj < P.ntaps) {
                    if (P.sn == 1 && cbase + 32 <= P.cols_total && ((P.sm | P.tap_stride) & 3) == 0) {
                        // columns are contiguous in dw: 8 x 16-byte reductions per thread instead of 32 scalar ones (the
                        // 32 rows of a warp lie sm floats apart, so every lane is its own L2 transaction either way)
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            red_add_v4(ob + cbase + 4 * j, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                                       __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int col = cbase + j;
                            if (col < P.cols_total) atomicAdd(ob + (long long)col * P.sn, __uint_as_float(r[j]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// Filter operand preparation (global -> global, the filter is small and L2 resident):
//   dst_hi[tap][r][c] = cond(src)   with optional [Ci][Co] -> [Co][Ci] transpose per tap
//   mode 1: cond = round-to-nearest TF32 ; mode 3: hi = src unchanged, lo = src - trunc(src) written to dst_lo
__global__ void prep_filter_k(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int taps,
                              int Ci, int Co, int transpose, int mode) {
    __shared__ float tile[32][33];
    const int tap = blockIdx.z;
    const size_t tb = (size_t)tap * Ci * Co;
    const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int ci = ci0 + i, co = co0 + threadIdx.x;
        tile[i][threadIdx.x] = (ci < Ci && co < Co) ? w[tb + (size_t)ci * Co + co] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        float v; size_t o; bool ok;
        if (transpose) { const int co = co0 + i, ci = ci0 + threadIdx.x; v = tile[threadIdx.x][i]; o = tb + (size_t)co * Ci + ci; ok = ci < Ci && co < Co; }
        else           { const int ci = ci0 + i, co = co0 + threadIdx.x; v = tile[i][threadIdx.x]; o = tb + (size_t)ci * Co + co; ok = ci < Ci && co < Co; }
        if (!ok) continue;
        const uint32_t u = __float_as_uint(v);
        if (mode == 3) { hi[o] = v; lo[o] = __uint_as_float(tf32_lo(u)); }
        else hi[o] = __uint_as_float(tf32_rna(u));
    }
}

// gather-mode filter: w [K][Co] (HWIO flattened) -> K-major [Co][Kpad] hi copy followed by the lo copy, zero padded
__global__ void prep_filter_gather_k(const float* __restrict__ w, float* __restrict__ hi, float* __restrict__ lo, int K,
                                     int Kpad, int Co, int mode) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Co * Kpad) return;
    const int co = i / Kpad, j = i - co * Kpad;
    const float v = j < K ? __ldg(w + (size_t)j * Co + co) : 0.f;
    const uint32_t u = __float_as_uint(v);
    if (mode == 3) { hi[i] = v; lo[i] = __uint_as_float(tf32_lo(u)); }
    else hi[i] = __uint_as_float(tf32_rna(u));
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int get_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) return eg_fail_arg("cuTensorMapEncodeTiled entry point", __FILE__, __LINE__);
    g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    return 0;
}

// 4-D NHWC activation map: dims {C, Wd, Hd, N} with element strides (1, sw, sh, sn) [floats], box {32, bw, bh, bn}
int make_act_map(CUtensorMap* m, const float* base, int C, int Wd, int Hd, int N, long long sw, long long sh,
                 long long sn, int bw, int bh, int bn, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)sw * 4, (cuuint64_t)sh * 4, (cuuint64_t)sn * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)base, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return eg_fail_arg("cuTensorMapEncodeTiled(activation)", __FILE__, __LINE__);
    return 0;
}
// 3-D filter map: dims {K, Nn, taps} (K contiguous), box {32, BN, 1}
int make_filter_map(CUtensorMap* m, const float* base, int K, int Nn, int taps, int BN) {
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)Nn, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)K * Nn * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return eg_fail_arg("cuTensorMapEncodeTiled(filter)", __FILE__, __LINE__);
    return 0;
}

int gcd(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

// pixel box {bw, bh, bn} with bw*bh*bn == pix, bw | W, bh | H
void pick_box(int W, int H, int pix, int& bw, int& bh, int& bn) {
    bw = gcd(W, pix); bh = gcd(H, pix / bw); bn = pix / (bw * bh);
}

// per-stream scratch (the only memory the library owns; grown at first use, i.e. during warm-up, never inside a
// captured region).  slot 0: prepared filter copies; slots 1, 2: patch matrix / small operands of conv_thin.cu
struct Scratch { float* p = nullptr; size_t bytes = 0; };
std::mutex g_mu;
std::map<std::pair<cudaStream_t, int>, Scratch> g_scratch;

int get_scratch(cudaStream_t st, size_t bytes, float** out, int slot = 0) {
    std::lock_guard<std::mutex> lk(g_mu);
    Scratch& s = g_scratch[std::make_pair(st, slot)];
    if (s.bytes < bytes) {
        if (s.p) { cudaStreamSynchronize(st); cudaFree(s.p); s.p = nullptr; s.bytes = 0; }
        const size_t floor_bytes = slot == 1 ? (64u << 20) : (slot == 0 ? (32u << 20) : (1u << 20));
        size_t want = bytes < floor_bytes ? floor_bytes : bytes;
        cudaError_t e = cudaMalloc(&s.p, want);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
        s.bytes = want;
    }
    *out = s.p;
    return 0;
}

int g_sms = 148;
constexpr int kStagesK3 = 4;     // TS-mode (3xTF32) kernel: 4 x (16 KB A landing + 2 x BN*128 B filter hi/lo)
constexpr int kStagesW = 3;
constexpr int kStagesW3 = 4;     // TS-mode wgrad: 4 x (16 KB rows + 2 x BN/32*4 KB columns hi/lo)

bool g_attr_set = false;
int set_attrs() {
    if (g_attr_set) return 0;
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kmajor<kStagesK3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) g_sms = n;
    }
    e = cudaFuncSetAttribute(conv_tc_wgrad<kStagesW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    e = cudaFuncSetAttribute(conv_tc_wgrad<kStagesW3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    e = cudaFuncSetAttribute(conv_tc_wgrad<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    g_attr_set = true;
    return 0;
}

int pick_bn(int n) { return n % 128 == 0 ? 128 : 64; }

// wgrad operand layout knobs: {TMA swizzle enum, UMMA layout type, SBO bytes} (tools/tc_probe.py can sweep them)
int g_dbg[8] = {(int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 1, 512, 0, 64, 2 | 4 | 32, 0, 0};

// Prepared-filter sets (eg_filter_set_*): the tensor-core kernels read a re-laid-out / hi+lo-split copy of the filter.
// Preparing it before every launch cost 221 small kernels per 14-class training step although the weights only change
// at the optimizer updates.  A set holds, for ALL conv filters of one network, persistent buffers with both prepared
// forms (forward: [tap][Co][Ci] hi, lo; input gradient: [tap][Ci][Co] hi, lo) and refreshes them with ONE launch
// (eg_filter_set_prepare) that the owner of the weights enqueues right after it writes them (RMSProp apply, upload,
// spectral normalisation).  Conv calls find the prepared copy by filter pointer and launch nothing.  Freshness is a
// stream-order property (writer kernel -> prepare kernel -> conv kernels, eager or replayed from a CUDA graph), not a
// host-side flag, so a replayed graph cannot leave a stale copy behind.  Filters outside any set are prepared per call.
struct FilterSetEntry {          // device-side descriptor of one filter
    const float* w; float* fwd; float* dgr;
    int taps, A, B;              // filter [taps][A][B]  (A = conv input channels, B = output channels)
    int tile0, tiles_b, tiles_ab;
    // kind 1 = thin filter (A <= 8): fwd = the gathered forward's K-major [B][Kpad] hi, lo; dgr = the patch-matrix input
    // gradient's operand [Npad][B] hi, lo (= the filter itself, zero-padded to Npad rows), see conv_thin.cu
    int kind, Kpad, Npad;
};
struct ManagedFilter { float* fwd; float* dgr; size_t n; int mode; int set; int kind; };
// buffers of a set that already ARE a forward-prepared operand (the thin input gradient's [Npad][B] copy): prep_filter()
// serves them as they are when such a buffer is passed as the filter of the dense product
std::map<const float*, size_t> g_preformed;
struct FilterSet { std::vector<const float*> keys, preformed; float* buf = nullptr; FilterSetEntry* table = nullptr; int nfilters = 0, tiles = 0, mode = 0; bool live = false; };
std::map<const float*, ManagedFilter> g_managed;
std::vector<FilterSet> g_sets;
unsigned long long g_managed_hits = 0;

__global__ void prep_filter_set_k(const FilterSetEntry* __restrict__ table, int nfilters, int mode) {
    __shared__ float tile[32][33];
    __shared__ FilterSetEntry e;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        int lo = 0, hi = nfilters - 1;                       // last filter with tile0 <= blockIdx.x
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (table[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1; }
        e = table[lo];
    }
    __syncthreads();
    const int local = (int)blockIdx.x - e.tile0;
    if (e.kind == 1) {
        const int K = e.taps * e.A, Co = e.B;
        const int i = local * 256 + threadIdx.y * 32 + threadIdx.x;
        if (i < Co * e.Kpad) {
            const int co = i / e.Kpad, j = i - co * e.Kpad;
            const float v = j < K ? e.w[(size_t)j * Co + co] : 0.f;
            const uint32_t u = __float_as_uint(v);
            if (mode == 3) { e.fwd[i] = v; e.fwd[(size_t)Co * e.Kpad + i] = __uint_as_float(tf32_lo(u)); }
            else e.fwd[i] = __uint_as_float(tf32_rna(u));
        }
        if (i < e.Npad * Co) {
            const float v = i < K * Co ? e.w[i] : 0.f;
            const uint32_t u = __float_as_uint(v);
            if (mode == 3) { e.dgr[i] = v; e.dgr[(size_t)e.Npad * Co + i] = __uint_as_float(tf32_lo(u)); }
            else e.dgr[i] = __uint_as_float(tf32_rna(u));
        }
        return;
    }
    const int tap = local / e.tiles_ab, rem = local - tap * e.tiles_ab;
    const int a0 = (rem / e.tiles_b) * 32, b0 = (rem % e.tiles_b) * 32;
    const size_t n = (size_t)e.taps * e.A * e.B, tb = (size_t)tap * e.A * e.B;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int a = a0 + i, b = b0 + threadIdx.x;
        float v = 0.f;
        if (a < e.A && b < e.B) {
            v = e.w[tb + (size_t)a * e.B + b];
            const uint32_t u = __float_as_uint(v);
            const size_t o = tb + (size_t)a * e.B + b;
            if (mode == 3) { e.dgr[o] = v; e.dgr[n + o] = __uint_as_float(tf32_lo(u)); }
            else e.dgr[o] = __uint_as_float(tf32_rna(u));
        }
        tile[i][threadIdx.x] = v;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int b = b0 + i, a = a0 + threadIdx.x;
        if (a < e.A && b < e.B) {
            const float v = tile[threadIdx.x][i];
            const uint32_t u = __float_as_uint(v);
            const size_t o = tb + (size_t)b * e.A + a;
            if (mode == 3) { e.fwd[o] = v; e.fwd[n + o] = __uint_as_float(tf32_lo(u)); }
            else e.fwd[o] = __uint_as_float(tf32_rna(u));
        }
    }
}

// prepared filter: returns the base of [hi copy (taps*Ci*Co)][lo copy (3x only)]
int prep_filter(const eg_conv_shape* s, const float* w, int transpose, int mode, cudaStream_t st, float** out) {
    const int taps = s->KH * s->KW;
    const size_t n = (size_t)taps * s->Ci * s->Co;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_managed.find(w);
        if (it != g_managed.end() && it->second.kind == 0 && it->second.n == n && it->second.mode == mode) {
            ++g_managed_hits;
            *out = transpose ? it->second.fwd : it->second.dgr;
            return 0;
        }
        auto pf = g_preformed.find(w);
        if (pf != g_preformed.end() && transpose && pf->second == n) {
            ++g_managed_hits;
            *out = const_cast<float*>(w);
            return 0;
        }
    }
    float* buf = nullptr;
    if (int r = get_scratch(st, sizeof(float) * n * 2, &buf)) return r;
    dim3 grid(eg_ceil_div(s->Co, 32), eg_ceil_div(s->Ci, 32), taps), block(32, 8);
    prep_filter_k<<<grid, block, 0, st>>>(w, buf, buf + n, taps, s->Ci, s->Co, transpose, mode);
    EG_CHECK_LAUNCH();
    *out = buf;
    return 0;
}

}  // namespace

extern "C" int eg_filter_set_create(const eg_filter_desc* descs, int n, int algo, long long* handle) {
    if (!descs || n <= 0 || !handle) return eg_fail_arg("eg_filter_set_create: arguments", __FILE__, __LINE__);
    if (algo != EG_ALGO_TC && algo != EG_ALGO_TC3X) return eg_fail_arg("eg_filter_set_create: algo must be EG_ALGO_TC or EG_ALGO_TC3X", __FILE__, __LINE__);
    const int mode = algo == EG_ALGO_TC3X ? 3 : 1;
    std::vector<FilterSetEntry> host(n);
    size_t total = 0;
    int tiles = 0;
    for (int i = 0; i < n; ++i) {
        const eg_filter_desc& d = descs[i];
        if (!d.w || d.taps <= 0 || d.Ci <= 0 || d.Co <= 0) return eg_fail_arg("eg_filter_set_create: descriptor", __FILE__, __LINE__);
        FilterSetEntry& e = host[i];
        e.w = d.w; e.taps = d.taps; e.A = d.Ci; e.B = d.Co;
        e.kind = 0; e.Kpad = 0; e.Npad = 0;
        const int K = d.taps * d.Ci;
        if (d.Ci <= 8 && K <= 128 && d.Co % 32 == 0) {           // thin filter: the operands of the gathered forward and of
            e.kind = 1;                                          // the patch-matrix input gradient (conv_thin.cu)
            e.Kpad = (K + 31) / 32 * 32; e.Npad = K <= 64 ? 64 : 128;
            e.tiles_b = 0; e.tiles_ab = 0;
            e.tile0 = tiles; tiles += eg_ceil_div((long long)std::max(e.Kpad, e.Npad) * d.Co, 256);
            total += 2 * (size_t)d.Co * e.Kpad + 2 * (size_t)e.Npad * d.Co + 128;
            continue;
        }
        e.tiles_b = eg_ceil_div(d.Co, 32); e.tiles_ab = e.tiles_b * eg_ceil_div(d.Ci, 32);
        e.tile0 = tiles; tiles += e.taps * e.tiles_ab;
        total += 4 * (size_t)d.taps * d.Ci * d.Co + 64;
    }
    FilterSet fs;
    cudaError_t err = cudaMalloc(&fs.buf, sizeof(float) * total);
    if (err != cudaSuccess) return eg_fail(err, __FILE__, __LINE__);
    err = cudaMalloc(&fs.table, sizeof(FilterSetEntry) * n);
    if (err != cudaSuccess) { cudaFree(fs.buf); return eg_fail(err, __FILE__, __LINE__); }
    float* p = fs.buf;
    for (int i = 0; i < n; ++i) {
        const size_t ni = (size_t)host[i].taps * host[i].A * host[i].B;
        if (host[i].kind == 1) {
            const size_t nf = 2 * (size_t)host[i].B * host[i].Kpad, nd = 2 * (size_t)host[i].Npad * host[i].B;
            host[i].fwd = p; p += (nf + 63) / 64 * 64;
            host[i].dgr = p; p += (nd + 63) / 64 * 64;
            continue;
        }
        host[i].fwd = p; host[i].dgr = p + 2 * ni;
        p += (4 * ni + 63) / 64 * 64;                        // keep every copy 256-byte aligned (TMA global address)
    }
    err = cudaMemcpy(fs.table, host.data(), sizeof(FilterSetEntry) * n, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) { cudaFree(fs.buf); cudaFree(fs.table); return eg_fail(err, __FILE__, __LINE__); }
    fs.nfilters = n; fs.tiles = tiles; fs.mode = mode; fs.live = true;
    std::lock_guard<std::mutex> lk(g_mu);
    const int id = (int)g_sets.size();
    for (int i = 0; i < n; ++i) {
        const size_t ni = (size_t)host[i].taps * host[i].A * host[i].B;
        auto old = g_managed.find(host[i].w);                // a pointer belongs to at most one set: the newest
        if (old != g_managed.end()) g_managed.erase(old);
        g_managed[host[i].w] = ManagedFilter{host[i].fwd, host[i].dgr, ni, mode, id, host[i].kind};
        fs.keys.push_back(host[i].w);
        if (host[i].kind == 1) {
            g_preformed[host[i].dgr] = (size_t)host[i].Npad * host[i].B;
            fs.preformed.push_back(host[i].dgr);
        }
    }
    g_sets.push_back(fs);
    *handle = id;
    return 0;
}

extern "C" int eg_filter_set_prepare(long long handle, cudaStream_t st) {
    FilterSet fs;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (handle < 0 || handle >= (long long)g_sets.size() || !g_sets[handle].live) return eg_fail_arg("eg_filter_set_prepare: handle", __FILE__, __LINE__);
        fs = g_sets[handle];
    }
    prep_filter_set_k<<<fs.tiles, dim3(32, 8), 0, st>>>(fs.table, fs.nfilters, fs.mode);
    EG_CHECK_LAUNCH();
    return 0;
}

extern "C" int eg_filter_set_destroy(long long handle) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (handle < 0 || handle >= (long long)g_sets.size() || !g_sets[handle].live) return 0;
    FilterSet& fs = g_sets[handle];
    for (const float* k : fs.keys) {
        auto it = g_managed.find(k);
        if (it != g_managed.end() && it->second.set == (int)handle) g_managed.erase(it);
    }
    for (const float* k : fs.preformed) g_preformed.erase(k);
    cudaDeviceSynchronize();                                 // no kernel may still read the copies
    cudaFree(fs.buf); cudaFree(fs.table);
    fs.buf = nullptr; fs.table = nullptr; fs.live = false; fs.keys.clear(); fs.preformed.clear();
    return 0;
}

// thin filter of a set: the gathered forward's operand (which = 0) or the patch-matrix input gradient's (which = 1);
// NULL when `w` is not in a set of this mode
float* eg_tc_managed_thin(const float* w, size_t n, int mode, int which) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_managed.find(w);
    if (it == g_managed.end() || it->second.kind != 1 || it->second.n != n || it->second.mode != mode) return nullptr;
    if (which == 0) ++g_managed_hits;        // (the input gradient's copy is counted when prep_filter() serves it)
    return which == 0 ? it->second.fwd : it->second.dgr;
}

// conv calls served from a prepared-filter set so far (no preparation launch)
extern "C" long long eg_filter_set_hits(void) { return (long long)g_managed_hits; }

int g_eg_thin_wgrad_off = 0;
extern int g_eg_small_off;        // conv_small.cu

extern "C" int eg_debug_set(int key, int value) {
    if (key < 0 || key >= 8) return -2;
    g_dbg[key] = value;
    if (key == 7) { g_eg_thin_wgrad_off = value & 1; g_eg_small_off = (value >> 1) & 1; }
    return 0;
}

int eg_tc_scratch(cudaStream_t st, int slot, size_t bytes, float** out) { return get_scratch(st, bytes, out, slot); }

// conv_thin.cu
int eg_thin_supported_fwd(const eg_conv_shape* s);
int eg_thin_supported_bwd_data(const eg_conv_shape* s);
int eg_thin_supported_bwd_weight(const eg_conv_shape* s);
int eg_thin_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int three_x, cudaStream_t st);
int eg_thin_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx, int three_x, cudaStream_t st);
int eg_thin_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int three_x, cudaStream_t st);

// ---- capability queries ------------------------------------------------------------------------------
// g_dbg[5]: which passes take the thin-channel route of conv_thin.cu (bit 0 fwd, bit 1 input grad, bit 2 filter grad).
// Measured on B200: the input gradient is 1.9x faster than the FFMA kernel (tools/thin_time.py); the filter gradient is
// 1.15x (critic first layer) to 2x (8 -> 128 classifier layers) faster since the round-2 filter-gradient kernel
// (tools/thin_routes_time.py: 442 -> 385, 452 -> 271, 216 -> 106 us); the forward is faster gathered (TcGather).  Default:
// bits 1 and 2 (and bit 5: the input gradient's scatter epilogue, below).
static bool thin_fwd(const eg_conv_shape* s) { return (g_dbg[5] & 1) && eg_thin_supported_fwd(s); }
static bool thin_bwd_data(const eg_conv_shape* s) { return (g_dbg[5] & 2) && eg_thin_supported_bwd_data(s); }
static bool thin_bwd_weight(const eg_conv_shape* s) { return (g_dbg[5] & 4) && eg_thin_supported_bwd_weight(s); }

// g_dbg[5] bit 3 turns the gather route off (tests compare the routes)
static bool gather_fwd(const eg_conv_shape* s) {
    if (g_dbg[5] & 8) return false;
    if (s->Ci < 1 || s->Ci > 8 || s->Co % 64) return false;
    if (s->KH > 8 || s->KW > 8 || s->KH * s->KW * s->Ci > kGatherMaxK) return false;
    if (((s->KH - 1) * s->W + s->KW) * s->Ci >= 65536) return false;              // table offsets are 16 bits
    if ((long long)s->N * s->OH * s->OW >= (1ll << 31)) return false;
    return s->stride >= 1;
}

int eg_tc_supported_fwd(const eg_conv_shape* s) {
    if (gather_fwd(s)) return 1;
    if (thin_fwd(s)) return 1;
    if (s->Ci % 32 || s->Co % 64) return 0;
    if (s->stride != 1 && s->stride != 2) return 0;
    if (s->KH * s->KW > kMaxTaps) return 0;
    return 1;
}
int eg_tc_supported_bwd_data(const eg_conv_shape* s) {
    if (thin_bwd_data(s)) return 1;
    if (s->Co % 32 || s->Ci % 64) return 0;
    if (s->stride != 1 && s->stride != 2) return 0;
    if (s->KH * s->KW > kMaxTaps) return 0;
    if (s->H % s->stride || s->W % s->stride) return 0;
    return 1;
}
// thin-channel filter gradient with the im2col rows gathered by the row-operand warps (3xTF32 kernel only; the plain
// TF32 mode of these layers runs on the FFMA kernel).  OFF by default (g_dbg[5] bit 4 turns it on): measured on B200
// (bench.py by_shape, critic first layer at 3 x 128 samples) 735 us against 445 us for the FFMA kernel -- only the two
// warps whose TMEM lanes hold valid im2col columns (48 of 128 rows) gather, 32 dependent-latency scalar loads per stage,
// one stage at a time; the forward's gather (8 warps, rows = pixels) does not have that problem.  Kept for the tests.
static bool gather_wgrad(const eg_conv_shape* s) {
    if (!(g_dbg[5] & 16)) return false;
    if (s->Ci < 1 || s->Ci > 8 || s->Co % 32) return false;
    if (s->KH > 16 || s->KW > 16 || s->KH * s->KW * s->Ci > 256) return false;
    if (s->OW > 255 * 32 || s->OH > 255 * 32) return false;
    return s->stride >= 1;
}

int eg_tc_supported_bwd_weight(const eg_conv_shape* s) {
    if (gather_wgrad(s)) return 1;
    if (thin_bwd_weight(s)) return 1;
    if (s->Ci % 32 || s->Co % 32) return 0;        // (a row side below 128 channels is zero-filled by TMA)
    if (s->stride != 1 && s->stride != 2) return 0;
    if (s->KH * s->KW > kMaxTaps) return 0;
    return 1;
}

// maps of x seen through the conv's stride: one per parity (ph, pw)
static int make_x_maps(CUtensorMap* maps, const eg_conv_shape* s, const float* x, int bw, int bh, int bn,
                       CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    const int S = s->stride;
    for (int ph = 0; ph < S; ++ph)
        for (int pw = 0; pw < S; ++pw) {
            const int Hp = (s->H - ph + S - 1) / S, Wp = (s->W - pw + S - 1) / S;
            const float* base = x + ((long long)ph * s->W + pw) * s->Ci;
            if (int r = make_act_map(&maps[ph * S + pw], base, s->Ci, Wp, Hp, s->N, (long long)S * s->Ci,
                                     (long long)S * s->W * s->Ci, (long long)s->H * s->W * s->Ci, bw, bh, bn, swz))
                return r;
        }
    return 0;
}
// tap (r, q) of a stride-S conv -> (parity map, offset) of the input pixel it reads for output pixel (oh, ow)
static TcTap x_tap(const eg_conv_shape* s, int r, int q, int bsel) {
    const int S = s->stride;
    const int ty = r - s->pad_t, tx = q - s->pad_l;
    const int ph = ((ty % S) + S) % S, pw = ((tx % S) + S) % S;
    TcTap t;
    t.amap = (signed char)(ph * S + pw);
    t.ay = (signed char)((ty - ph) / S);
    t.ax = (signed char)((tx - pw) / S);
    t.bsel = (signed char)bsel;
    return t;
}

// split the (tap, k-chunk) loop over several CTAs when the output tiles alone cannot fill the machine
static int pick_ksplit(int ctas, int min_iters) {
    if (ctas >= 100) return 1;
    int k = 148 / ctas;
    const int cap = min_iters / 8;            // keep >= 8 stages per CTA
    if (k > cap) k = cap;
    return k < 1 ? 1 : k;
}

static size_t kmajor_smem(int BN, int mode) {
    return (size_t)kStagesK3 * (128 * 128 + (mode == 3 ? 2 : 1) * BN * 128) + 4 * 4096 + 1024;
}
// -> 0: epilogue fused (or none requested); 1: requested but not of the sign type -> the caller applies it afterwards
static int set_epilogue(TcParams& P, const EgEpi* epi) {
    P.epi = EG_EPI_NONE; P.epi_neg = 0.f; P.epi_ge = 0; P.mask = nullptr;
    if (!epi || epi->mode == EG_EPI_NONE) return 0;
    switch (epi->act) {
        case EG_ACT_RELU: P.epi_neg = 0.f; P.epi_ge = 0; break;
        case EG_ACT_LRELU_BLOCK: P.epi_neg = 0.2f; P.epi_ge = 1; break;
        case EG_ACT_LRELU: P.epi_neg = 0.2f; P.epi_ge = 0; break;
        default: return 1;
    }
    P.epi = epi->mode; P.mask = epi->mask;
    return 0;
}

static int launch_kmajor(const TcMaps& maps, TcParams& P, int gx, int gy, int mode, cudaStream_t st) {
    P.grid_x = gx; P.grid_y = gy;
    const int total = gx * gy * P.nphases * P.ksplit;
    auto fd = [&](int d, long long nmax) {
        FastDiv f; f.d = (uint32_t)d;
        f.m = (d > 1 && nmax * d < (1ll << 32)) ? (uint32_t)(((1ull << 32) + (uint32_t)d - 1) / (uint32_t)d) : 0u;
        return f;
    };
    P.fgx = fd(gx, total); P.fgy = fd(gy, total); P.fks = fd(P.ksplit, total);
    for (int i = 0; i < P.nphases; ++i) { P.ph[i].ftw = fd(P.ph[i].tiles_w, gx); P.ph[i].fth = fd(P.ph[i].tiles_h, gx); }
    const int ctas = (total < g_sms || (P.dbg & 32)) ? total : g_sms;              // persistent: at most one CTA per SM
    conv_tc_kmajor<kStagesK3><<<ctas, kThreadsK, kmajor_smem(P.BN, mode), st>>>(maps, P);
    return 0;
}

// `epi` (may be NULL): fused epilogue; returns 1 instead of 0 when it was NOT applied (thin route): the caller runs it
// thin-channel forward, A gathered by the conditioning warps (TcGather): y[P, Co] = im2col(x)[P, Kpad] . w[Kpad, Co]
static int tc_conv2d_fwd_gather(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y,
                                int three_x, cudaStream_t st, const EgEpi* epi) {
    if (int r = get_encode()) return r;
    if (int r = set_attrs()) return r;
    const int mode = three_x ? 3 : 1;
    const int K = s->KH * s->KW * s->Ci, Kpad = (K + 31) / 32 * 32;
    const long long Ppix = (long long)s->N * s->OH * s->OW;
    float* wt = eg_tc_managed_thin(w, (size_t)K * s->Co, mode, 0);
    if (wt == nullptr) {
        if (int r = get_scratch(st, sizeof(float) * (size_t)2 * s->Co * Kpad, &wt, 3)) return r;
        prep_filter_gather_k<<<eg_ceil_div((long long)s->Co * Kpad, 256), 256, 0, st>>>(w, wt, wt + (size_t)s->Co * Kpad, K, Kpad, s->Co, mode);
        EG_CHECK_LAUNCH();
    }
    TcMaps maps;
    TcParams P{};
    P.bw = 1; P.bh = 1; P.bn = 128;
    P.BN = pick_bn(s->Co);
    P.mode = mode; P.b_lo_tap_off = 1; P.dbg = g_dbg[3];
    if (int r = make_filter_map(&maps.b[0], wt, Kpad, s->Co, 2, P.BN)) return r;
    for (int i = 1; i < 4; ++i) maps.b[i] = maps.b[0];
    for (int i = 0; i < 4; ++i) maps.a[i] = maps.b[0];       // never used for loads in gather mode (prefetch only)
    P.nphases = 1; P.ldn = s->Co; P.bias = bias; P.out = y;
    TcPhase& ph = P.ph[0];
    ph.ntaps = 1; ph.kchunks = Kpad / 32;
    TcTap t0; t0.amap = 0; t0.ax = 0; t0.ay = 0; t0.bsel = 0;
    ph.taps[0] = t0;
    ph.ext_w = 1; ph.ext_h = 1; ph.ext_n = (int)Ppix;
    ph.tiles_w = 1; ph.tiles_h = 1; ph.tiles_n = eg_ceil_div(Ppix, 128);
    ph.out_off = 0; ph.sw = s->Co; ph.sh = s->Co; ph.sn = s->Co;
    P.ksplit = 1;
    TcGather& g = P.g;
    g.x = x; g.on = 1; g.H = s->H; g.W = s->W; g.C = s->Ci; g.OH = s->OH; g.OW = s->OW; g.KH = s->KH; g.KW = s->KW;
    g.stride = s->stride; g.pad_t = s->pad_t; g.pad_l = s->pad_l; g.K = K; g.P = Ppix;
    auto fd = [&](int d, long long nmax) {
        FastDiv f; f.d = (uint32_t)d;
        f.m = (d > 1 && nmax * d < (1ll << 32)) ? (uint32_t)(((1ull << 32) + (uint32_t)d - 1) / (uint32_t)d) : 0u;
        return f;
    };
    g.fow = fd(s->OW, Ppix); g.foh = fd(s->OH, Ppix);
    const int unfused = set_epilogue(P, epi);
    launch_kmajor(maps, P, ph.tiles_n, s->Co / P.BN, mode, st);
    EG_CHECK_LAUNCH();
    return unfused;
}

// `scatter` (conv_thin.cu): the product is not stored as y[P, Co] but col2im-scattered into y = dx of the conv `scatter`
// describes (s is then the dense [P pixels] x [Ci] . [Ci x Npad] product, Npad = one column tile)
// conv_small.cu: streaming kernels of the 1x1 convolution between 8 and 128 channels (-100 = other shape / switched off)
int eg_thin1x1(const eg_conv_shape* s, int which, const float* a, const float* w, const float* bias, float* out, int sms,
               cudaStream_t st);

int eg_tc_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int three_x,
                     cudaStream_t st, const EgEpi* epi, const eg_conv_shape* scatter) {
    if (scatter == nullptr && !(g_dbg[5] & 64)) {
        const int r = eg_thin1x1(s, 0, x, w, bias, y, g_sms, st);
        if (r != -100) return r ? r : ((epi && epi->mode != EG_EPI_NONE) ? 1 : 0);
    }
    if (scatter == nullptr && gather_fwd(s)) return tc_conv2d_fwd_gather(s, x, w, bias, y, three_x, st, epi);
    if (thin_fwd(s)) {
        if (int r = eg_thin_conv2d_fwd(s, x, w, bias, y, three_x, st)) return r;
        return (epi && epi->mode != EG_EPI_NONE) ? 1 : 0;
    }
    if (int r = get_encode()) return r;
    if (int r = set_attrs()) return r;
    const int taps = s->KH * s->KW, mode = three_x ? 3 : 1;
    float* wt = nullptr;
    if (int r = prep_filter(s, w, 1, mode, st, &wt)) return r;       // HWIO -> HWOI (K-major B operand)
    TcMaps maps;
    TcParams P{};
    pick_box(s->OW, s->OH, 128, P.bw, P.bh, P.bn);
    P.BN = pick_bn(s->Co);
    P.mode = mode; P.b_lo_tap_off = taps; P.dbg = g_dbg[3];
    if (int r = make_x_maps(maps.a, s, x, P.bw, P.bh, P.bn)) return r;
    for (int i = s->stride * s->stride; i < 4; ++i) maps.a[i] = maps.a[0];
    if (int r = make_filter_map(&maps.b[0], wt, s->Ci, s->Co, 2 * taps, P.BN)) return r;
    for (int i = 1; i < 4; ++i) maps.b[i] = maps.b[0];
    P.nphases = 1; P.ldn = s->Co; P.bias = bias; P.out = y;
    TcPhase& ph = P.ph[0];
    ph.ntaps = taps; ph.kchunks = s->Ci / 32;
    for (int r = 0; r < s->KH; ++r)
        for (int q = 0; q < s->KW; ++q) ph.taps[r * s->KW + q] = x_tap(s, r, q, r * s->KW + q);
    ph.ext_w = s->OW; ph.ext_h = s->OH; ph.ext_n = s->N;
    ph.tiles_w = s->OW / P.bw; ph.tiles_h = s->OH / P.bh; ph.tiles_n = eg_ceil_div(s->N, P.bn);
    ph.out_off = 0; ph.sw = s->Co; ph.sh = (long long)s->OW * s->Co; ph.sn = (long long)s->OH * s->OW * s->Co;
    P.ksplit = pick_ksplit(ph.tiles_w * ph.tiles_h * ph.tiles_n * (s->Co / P.BN), ph.ntaps * ph.kchunks);
    const int unfused = set_epilogue(P, epi);
    if (P.epi == EG_EPI_ACT) P.ksplit = 1;                  // a non-linear epilogue needs the whole sum in one CTA
    if (scatter != nullptr) {
        if (s->Co != P.BN || s->H != 1 || s->W != 1 || P.epi != EG_EPI_NONE || bias != nullptr)
            return eg_fail_arg("scatter epilogue: one column tile, pixel-list product, no bias / activation", __FILE__, __LINE__);
        TcGather& g = P.g;
        g.x = nullptr; g.on = 2; g.H = scatter->H; g.W = scatter->W; g.C = scatter->Ci; g.OH = scatter->OH; g.OW = scatter->OW;
        g.KH = scatter->KH; g.KW = scatter->KW; g.stride = scatter->stride; g.pad_t = scatter->pad_t; g.pad_l = scatter->pad_l;
        g.K = scatter->KH * scatter->KW * scatter->Ci; g.P = (long long)scatter->N * scatter->OH * scatter->OW;
        auto fd = [&](int d, long long nmax) {
            FastDiv f; f.d = (uint32_t)d;
            f.m = (d > 1 && nmax * d < (1ll << 32)) ? (uint32_t)(((1ull << 32) + (uint32_t)d - 1) / (uint32_t)d) : 0u;
            return f;
        };
        g.fow = fd(scatter->OW, g.P); g.foh = fd(scatter->OH, g.P);
    } else if (P.ksplit > 1) {
        cudaError_t e = cudaMemsetAsync(y, 0, sizeof(float) * (size_t)s->N * s->OH * s->OW * s->Co, st);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    launch_kmajor(maps, P, ph.tiles_w * ph.tiles_h * ph.tiles_n, s->Co / P.BN, mode, st);
    EG_CHECK_LAUNCH();
    return unfused;
}

// whether the thin input gradient may use the scatter epilogue (conv_thin.cu)
int eg_tc_scatter_supported(const eg_conv_shape* c) {
    const int K = c->KH * c->KW * c->Ci;
    // ON by default (g_dbg[5] bit 5).  With warm inputs it measured equal to the product matrix + col2im pass (94.5 vs 94 us,
    // critic first layer at batch 64); with cold inputs, as inside the step, it is 6-40 % faster at batch 128
    // (tools/thin_routes_time.py: 201 -> 123, 126 -> 99, 465 -> 436, 174 -> 159, 100 -> 73, 65 -> 50 us): the product matrix
    // never goes to HBM.  The sums arrive by red.add, so the last bits depend on the order (like the filter gradients').
    if (!(g_dbg[5] & 32)) return 0;
    if (K > kGatherMaxK || K > 128 || c->KH > 8 || c->KW > 8) return 0;
    if (((c->KH - 1) * c->W + c->KW) * c->Ci >= 65536) return 0;
    if ((long long)c->N * c->OH * c->OW >= (1ll << 31)) return 0;
    return 1;
}

int eg_tc_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx,
                          int three_x, cudaStream_t st, const EgEpi* epi) {
    if (!(g_dbg[5] & 64)) {
        const int r = eg_thin1x1(s, 1, dy, w, bias, dx, g_sms, st);
        if (r != -100) return r ? r : ((epi && epi->mode != EG_EPI_NONE) ? 1 : 0);
    }
    if (thin_bwd_data(s)) {
        if (int r = eg_thin_conv2d_bwd_data(s, dy, w, bias, dx, three_x, st)) return r;
        return (epi && epi->mode != EG_EPI_NONE) ? 1 : 0;
    }
    if (int r = get_encode()) return r;
    if (int r = set_attrs()) return r;
    const int S = s->stride, taps = s->KH * s->KW, mode = three_x ? 3 : 1;
    float* wp = nullptr;
    if (int r = prep_filter(s, w, 0, mode, st, &wp)) return r;       // HWIO is already K-major for this GEMM
    TcMaps maps;
    TcParams P{};
    const int Hp = s->H / S, Wp = s->W / S;
    pick_box(Wp, Hp, 128, P.bw, P.bh, P.bn);
    P.BN = pick_bn(s->Ci);
    P.mode = mode; P.b_lo_tap_off = taps; P.dbg = g_dbg[3];
    // A = dy, dense stride-1 window
    if (int r = make_act_map(&maps.a[0], dy, s->Co, s->OW, s->OH, s->N, s->Co, (long long)s->OW * s->Co,
                             (long long)s->OH * s->OW * s->Co, P.bw, P.bh, P.bn)) return r;
    for (int i = 1; i < 4; ++i) maps.a[i] = maps.a[0];
    // B = filter viewed as [tap][Ci rows][Co contiguous]
    if (int r = make_filter_map(&maps.b[0], wp, s->Co, s->Ci, 2 * taps, P.BN)) return r;
    for (int i = 1; i < 4; ++i) maps.b[i] = maps.b[0];
    P.nphases = S * S; P.ldn = s->Ci; P.bias = bias; P.out = dx;
    int max_tiles = 0;
    for (int py = 0; py < S; ++py)
        for (int px = 0; px < S; ++px) {
            TcPhase& ph = P.ph[py * S + px];
            const int r0 = (py + s->pad_t) % S, q0 = (px + s->pad_l) % S;
            const int nr = r0 < s->KH ? (s->KH - r0 + S - 1) / S : 0, nq = q0 < s->KW ? (s->KW - q0 + S - 1) / S : 0;
            const int d0y = (py + s->pad_t - r0) / S, d0x = (px + s->pad_l - q0) / S;
            ph.ntaps = nr * nq; ph.kchunks = s->Co / 32;
            for (int j = 0; j < nr; ++j)
                for (int jq = 0; jq < nq; ++jq) {
                    TcTap t;
                    t.amap = 0; t.ay = (signed char)(d0y - j); t.ax = (signed char)(d0x - jq);
                    t.bsel = (signed char)((r0 + S * j) * s->KW + (q0 + S * jq));
                    ph.taps[j * nq + jq] = t;
                }
            ph.ext_w = Wp; ph.ext_h = Hp; ph.ext_n = s->N;
            ph.tiles_w = Wp / P.bw; ph.tiles_h = Hp / P.bh; ph.tiles_n = eg_ceil_div(s->N, P.bn);
            ph.out_off = ((long long)py * s->W + px) * s->Ci;
            ph.sw = (long long)S * s->Ci; ph.sh = (long long)S * s->W * s->Ci; ph.sn = (long long)s->H * s->W * s->Ci;
            const int nt = ph.tiles_w * ph.tiles_h * ph.tiles_n;
            if (nt > max_tiles) max_tiles = nt;
            if (ph.ntaps == 0) return eg_fail_arg("dgrad phase without taps", __FILE__, __LINE__);
        }
    int min_iters = 1 << 30;
    for (int i = 0; i < S * S; ++i) min_iters = P.ph[i].ntaps * P.ph[i].kchunks < min_iters ? P.ph[i].ntaps * P.ph[i].kchunks : min_iters;
    P.ksplit = pick_ksplit(max_tiles * (s->Ci / P.BN) * S * S, min_iters);
    const int unfused = set_epilogue(P, epi);
    if (P.epi == EG_EPI_ACT) P.ksplit = 1;
    if (P.ksplit > 1) {
        cudaError_t e = cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)s->N * s->H * s->W * s->Ci, st);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    launch_kmajor(maps, P, max_tiles, s->Ci / P.BN, mode, st);
    EG_CHECK_LAUNCH();
    return unfused;
}

int eg_simt_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int sms,
                              cudaStream_t st);

// dw[K, Co] (+)= im2col(x)^T[K, P] . dy[P, Co]: rows = im2col columns (gathered), columns = dy channels (TMA pixel boxes)
static int tc_conv2d_bwd_weight_gather(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate,
                                       cudaStream_t st) {
    if (int r = get_encode()) return r;
    if (int r = set_attrs()) return r;
    TcMaps maps;
    WgParams P{};
    P.mode = 3; P.pix = 32;
    pick_box(s->OW, s->OH, P.pix, P.bw, P.bh, P.bn);
    const CUtensorMapSwizzle swz = (CUtensorMapSwizzle)g_dbg[0];
    P.layout_type = g_dbg[1]; P.sbo = g_dbg[2];
    if (int r = make_act_map(&maps.b[0], dy, s->Co, s->OW, s->OH, s->N, s->Co, (long long)s->OW * s->Co,
                             (long long)s->OH * s->OW * s->Co, P.bw, P.bh, P.bn, swz)) return r;
    for (int i = 1; i < 4; ++i) maps.b[i] = maps.b[0];
    for (int i = 0; i < 4; ++i) maps.a[i] = maps.b[0];       // never used for loads
    const int K = s->KH * s->KW * s->Ci;
    P.ntaps = 1;
    TcTap t0; t0.amap = 0; t0.ax = 0; t0.ay = 0; t0.bsel = 0;
    P.taps[0] = t0;
    P.x_is_a = 1;
    P.BN = s->Co % 128 == 0 ? 128 : (s->Co % 64 == 0 ? 64 : 32);
    P.rows_total = K; P.cols_total = s->Co;
    P.tap_group = 1; P.tap_stride = 0; P.sm = s->Co; P.sn = 1;
    P.tiles_w = s->OW / P.bw; P.tiles_h = s->OH / P.bh; P.tiles_n = eg_ceil_div(s->N, P.bn);
    const int ntiles = P.tiles_w * P.tiles_h * P.tiles_n;
    const int row_tiles = eg_ceil_div(K, 128), col_tiles = s->Co / P.BN;
    int splits = eg_ceil_div(2 * 148, row_tiles * col_tiles);
    if (splits > ntiles) splits = ntiles;
    if (splits < 1) splits = 1;
    P.chunks_per_split = eg_ceil_div(ntiles, splits);
    if (P.chunks_per_split > g_dbg[4]) P.chunks_per_split = g_dbg[4];
    splits = eg_ceil_div(ntiles, P.chunks_per_split);
    P.out = dw;
    TcGather& g = P.g;
    g.x = x; g.on = 1; g.H = s->H; g.W = s->W; g.C = s->Ci; g.OH = s->OH; g.OW = s->OW; g.KH = s->KH; g.KW = s->KW;
    g.stride = s->stride; g.pad_t = s->pad_t; g.pad_l = s->pad_l; g.K = K; g.P = (long long)s->N * s->OH * s->OW;
    P.gN = s->N;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)K * s->Co, st);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    dim3 grid(row_tiles, col_tiles, splits);
    const size_t smem = (size_t)kStagesW3 * ((4 + 2 * (P.BN / 32)) * P.pix * 128) + 1024;
    conv_tc_wgrad<kStagesW3, true><<<grid, kThreadsW3, smem, st>>>(maps, P);
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_tc_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate,
                            int three_x, cudaStream_t st) {
    if (gather_wgrad(s)) {
        if (three_x) return tc_conv2d_bwd_weight_gather(s, x, dy, dw, accumulate, st);
        if (!thin_bwd_weight(s)) return eg_simt_conv2d_bwd_weight(s, x, dy, dw, accumulate, g_sms, st);
    }
    if (thin_bwd_weight(s)) return eg_thin_conv2d_bwd_weight(s, x, dy, dw, accumulate, three_x, st);
    if (int r = get_encode()) return r;
    if (int r = set_attrs()) return r;
    TcMaps maps;
    WgParams P{};
    P.mode = three_x ? 3 : 1;
    P.pix = three_x ? 32 : 64;
    pick_box(s->OW, s->OH, P.pix, P.bw, P.bh, P.bn);
    const CUtensorMapSwizzle swz = (CUtensorMapSwizzle)g_dbg[0];
    P.layout_type = g_dbg[1]; P.sbo = g_dbg[2];
    if (int r = make_x_maps(maps.a, s, x, P.bw, P.bh, P.bn, swz)) return r;
    for (int i = s->stride * s->stride; i < 4; ++i) maps.a[i] = maps.a[0];
    if (int r = make_act_map(&maps.b[0], dy, s->Co, s->OW, s->OH, s->N, s->Co, (long long)s->OW * s->Co,
                             (long long)s->OH * s->OW * s->Co, P.bw, P.bh, P.bn, swz)) return r;
    for (int i = 1; i < 4; ++i) maps.b[i] = maps.b[0];
    P.ntaps = s->KH * s->KW;
    for (int r = 0; r < s->KH; ++r)
        for (int q = 0; q < s->KW; ++q) P.taps[r * s->KW + q] = x_tap(s, r, q, 0);
    // rows = whichever channel count fills 128 accumulator rows; prefer the wider one as rows
    if (s->Ci % 128 && s->Co % 128) P.x_is_a = s->Ci >= s->Co ? 1 : 0;   // neither fills the rows: the rest is zero-filled
    else P.x_is_a = (s->Ci % 128 == 0 && (s->Co % 128 != 0 || s->Ci >= s->Co)) ? 1 : 0;
    const int rows = P.x_is_a ? s->Ci : s->Co, cols = P.x_is_a ? s->Co : s->Ci;
    P.BN = cols % 128 == 0 ? 128 : (cols % 64 == 0 ? 64 : 32);
    P.rows_total = rows; P.cols_total = cols;
    // a 64-channel x side: two taps side by side fill the 128-column tile (N = 64 MMAs leave the tensor pipe half
    // empty and every extra CTA re-conditions the same row operand)
    P.tap_group = 1;
    if (!P.x_is_a && cols == 64 && (g_dbg[6] & 1) == 0) { P.tap_group = 2; P.BN = 128; }
    P.tap_stride = (long long)s->Ci * s->Co;
    P.sm = P.x_is_a ? s->Co : 1; P.sn = P.x_is_a ? 1 : s->Co;
    P.tiles_w = s->OW / P.bw; P.tiles_h = s->OH / P.bh; P.tiles_n = eg_ceil_div(s->N, P.bn);
    const int ntiles = P.tiles_w * P.tiles_h * P.tiles_n;
    const int row_tiles = eg_ceil_div(rows, 128);
    const int col_tiles = P.tap_group > 1 ? 1 : cols / P.BN;
    const int tap_groups = eg_ceil_div(P.ntaps, P.tap_group);
    const int base_ctas = row_tiles * col_tiles * tap_groups;
    int splits = eg_ceil_div(2 * 148, base_ctas);
    if (splits > ntiles) splits = ntiles;
    if (splits < 1) splits = 1;
    P.chunks_per_split = eg_ceil_div(ntiles, splits);
    if (P.mode == 3 && P.chunks_per_split > g_dbg[4]) P.chunks_per_split = g_dbg[4];   // bound the truncating accumulation chain
    splits = eg_ceil_div(ntiles, P.chunks_per_split);
    P.out = dw;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)P.ntaps * s->Ci * s->Co, st);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    dim3 grid(row_tiles, col_tiles, tap_groups * splits);
    if (P.mode == 3) {
        // Two CTAs per SM with a 2-stage ring each (one CTA's prologue / atomic epilogue overlaps the other's main loop)
        // instead of one CTA with 4 stages: +2 ... +10 % on the single-tap-per-CTA layouts, -2 % with tap pairs
        // (tools/wgrad_time.py).  g_dbg[6] bit 1 forces the 4-stage kernel.
        if (!(g_dbg[6] & 2) && P.tap_group == 1) {
            const size_t smem = (size_t)2 * ((4 + 2 * (P.BN / 32)) * P.pix * 128) + 1024;
            conv_tc_wgrad<2, true><<<grid, kThreadsW3, smem, st>>>(maps, P);
        } else {
            const size_t smem = (size_t)kStagesW3 * ((4 + 2 * (P.BN / 32)) * P.pix * 128) + 1024;
            conv_tc_wgrad<kStagesW3, true><<<grid, kThreadsW3, smem, st>>>(maps, P);
        }
    } else {
        const size_t smem = (size_t)kStagesW * ((4 + P.BN / 32) * P.pix * 128) + 1024;
        conv_tc_wgrad<kStagesW, false><<<grid, kThreads, smem, st>>>(maps, P);
    }
    EG_CHECK_LAUNCH();
    return 0;
}
