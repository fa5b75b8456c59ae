// fp32 FFMA implicit-GEMM convolution trio (forward / input-gradient / filter-gradient).
//
// Role in the design (DESIGN.md "kernels"): exact-fp32 path for the thin-channel, HBM-bound layers
// (Cin = 3 or Cout = 3: d_conv_0, e_resnet_64_0, g_dconv_4, g_lin_0) and the on-GPU checker for the
// tcgen05 kernels in conv_tc.cu.  One kernel template serves the three GEMMs through a "row table" /
// "k table" gather:  A(m,k) = src[row[m].off + kk[k].off] if the (y,x) sums are inside the image,
// B(k,n) = wsrc[kk[k].aux + n * bstride].
//
// Replaces tf.nn.conv2d / tf.nn.conv2d_transpose and their gradients
// (reference nn/modules/conv.py:26,29,49).
#include "common.cuh"

namespace {

struct ConvP { int N, H, W, Ci, OH, OW, Co, KH, KW, S, PT, PL; int epi, epi_ge; float epi_neg; const float* mask; };

enum { M_FWD = 0, M_DGRAD = 1, M_WGRAD = 2 };

struct Phase {            // per-launch GEMM extents (+ dgrad phase / wgrad split parameters)
    int M, N, K;
    int ph, pw, Hp, Wp;   // dgrad: parity of the input pixel this launch covers, extents of that sub-grid
    int r0, q0, nr, nq;   // dgrad: first valid tap and tap counts of this parity
    int d0y, d0x;         // dgrad: output-pixel offset of tap (r0, q0)
    int k_per_split;      // wgrad: pixels per split-K slice
};

struct Info { int off, y, x, aux; };

constexpr int kInvalid = -(1 << 28);

template <int MODE>
__device__ __forceinline__ Info decode_row(const ConvP& p, const Phase& f, int m) {
    Info r;
    if (m >= f.M) { r.off = 0; r.y = kInvalid; r.x = kInvalid; r.aux = 0; return r; }
    if (MODE == M_FWD) {
        int ow = m % p.OW, t = m / p.OW, oh = t % p.OH, n = t / p.OH;
        r.y = oh * p.S - p.PT; r.x = ow * p.S - p.PL;
        r.off = ((n * p.H + r.y) * p.W + r.x) * p.Ci;
        r.aux = m * p.Co;
    } else if (MODE == M_DGRAD) {
        int iw2 = m % f.Wp, t = m / f.Wp, ih2 = t % f.Hp, n = t / f.Hp;
        r.y = ih2; r.x = iw2;
        r.off = ((n * p.OH + ih2) * p.OW + iw2) * p.Co;
        r.aux = ((n * p.H + ih2 * p.S + f.ph) * p.W + iw2 * p.S + f.pw) * p.Ci;
    } else {
        int ci = m % p.Ci, t = m / p.Ci, q = t % p.KW, rr = t / p.KW;
        r.y = rr; r.x = q;
        r.off = (rr * p.W + q) * p.Ci + ci;
        r.aux = m * p.Co;
    }
    return r;
}

template <int MODE>
__device__ __forceinline__ Info decode_k(const ConvP& p, const Phase& f, int k, int kend) {
    Info r;
    if (k >= kend) { r.off = 0; r.y = kInvalid; r.x = kInvalid; r.aux = -1; return r; }
    if (MODE == M_FWD) {
        int ci = k % p.Ci, t = k / p.Ci, q = t % p.KW, rr = t / p.KW;
        r.y = rr; r.x = q;
        r.off = (rr * p.W + q) * p.Ci + ci;
        r.aux = k * p.Co;
    } else if (MODE == M_DGRAD) {
        int co = k % p.Co, t = k / p.Co, jq = t % f.nq, j = t / f.nq;
        r.y = f.d0y - j; r.x = f.d0x - jq;
        r.off = (r.y * p.OW + r.x) * p.Co + co;
        int rr = f.r0 + p.S * j, q = f.q0 + p.S * jq;
        r.aux = ((rr * p.KW + q) * p.Ci) * p.Co + co;
    } else {
        int ow = k % p.OW, t = k / p.OW, oh = t % p.OH, n = t / p.OH;
        r.y = oh * p.S - p.PT; r.x = ow * p.S - p.PL;
        r.off = ((n * p.H + r.y) * p.W + r.x) * p.Ci;
        r.aux = k * p.Co;
    }
    return r;
}

template <int MODE, int BM, int BN, int TM, int TN, int VA, int VB>
__global__ void __launch_bounds__(256)
igemm_simt(const float* __restrict__ asrc, const float* __restrict__ bsrc, const float* __restrict__ bias,
           float* __restrict__ dst, ConvP p, Phase f0, int atomic_out) {
    constexpr int BK = 16;
    constexpr int NT = 256;
    static_assert((BM / TM) * (BN / TN) == NT, "tile/thread mismatch");
    constexpr bool A_KC = (MODE != M_WGRAD);   // A contiguous along k (else along m)
    constexpr bool B_KC = (MODE == M_DGRAD);   // B contiguous along k (else along n)
    constexpr int A_ITEMS = (BM * BK / VA + NT - 1) / NT;
    constexpr int B_ITEMS = (BK * BN / VB + NT - 1) / NT;

    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    __shared__ Info rowinfo[BM];
    __shared__ Info kinfo[2][BK];

    Phase f = f0;
    int kbeg = 0, kend = f.K;
    if (MODE == M_DGRAD) {
        // blockIdx.z enumerates the stride*stride input-pixel parities
        const int S = p.S;
        f.ph = blockIdx.z / S; f.pw = blockIdx.z % S;
        f.Hp = (p.H - f.ph + S - 1) / S; f.Wp = (p.W - f.pw + S - 1) / S;
        f.r0 = (f.ph + p.PT) % S; f.q0 = (f.pw + p.PL) % S;
        f.nr = f.r0 < p.KH ? (p.KH - f.r0 + S - 1) / S : 0;
        f.nq = f.q0 < p.KW ? (p.KW - f.q0 + S - 1) / S : 0;
        f.d0y = (f.ph + p.PT - f.r0) / S; f.d0x = (f.pw + p.PL - f.q0) / S;
        f.M = p.N * f.Hp * f.Wp;
        f.K = f.nr * f.nq * p.Co;
        kend = f.K;
    } else if (MODE == M_WGRAD) {
        kbeg = blockIdx.z * f.k_per_split;
        kend = min(f.K, kbeg + f.k_per_split);
    }
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    if (m0 >= f.M) return;
    const int tid = threadIdx.x;
    const int ylim = (MODE == M_DGRAD) ? p.OH : p.H;
    const int xlim = (MODE == M_DGRAD) ? p.OW : p.W;
    const int bstride = (MODE == M_DGRAD) ? p.Co : 1;

    for (int i = tid; i < BM; i += NT) rowinfo[i] = decode_row<MODE>(p, f, m0 + i);
    if (tid < BK) kinfo[0][tid] = decode_k<MODE>(p, f, kbeg + tid, kend);
    __syncthreads();

    float ra[A_ITEMS][VA], rb[B_ITEMS][VB];

    auto load_tile = [&](int buf) {
#pragma unroll
        for (int it = 0; it < A_ITEMS; ++it) {
            const int v = tid + it * NT;
            int mm, kk;
            if (A_KC) { kk = (v % (BK / VA)) * VA; mm = v / (BK / VA); }
            else      { mm = (v % (BM / VA)) * VA; kk = v / (BM / VA); }
#pragma unroll
            for (int e = 0; e < VA; ++e) ra[it][e] = 0.f;
            if (v < BM * BK / VA) {
                if (VA == 4) {
                    const Info r = rowinfo[mm], k = kinfo[buf][kk];
                    if ((unsigned)(r.y + k.y) < (unsigned)ylim && (unsigned)(r.x + k.x) < (unsigned)xlim) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(asrc + (r.off + k.off)));
                        ra[it][0] = t.x; ra[it][1 % VA] = t.y; ra[it][2 % VA] = t.z; ra[it][3 % VA] = t.w;
                    }
                } else {
                    const Info r = rowinfo[mm], k = kinfo[buf][kk];
                    if ((unsigned)(r.y + k.y) < (unsigned)ylim && (unsigned)(r.x + k.x) < (unsigned)xlim)
                        ra[it][0] = __ldg(asrc + (r.off + k.off));
                }
            }
        }
#pragma unroll
        for (int it = 0; it < B_ITEMS; ++it) {
            const int v = tid + it * NT;
            int nn, kk;
            if (B_KC) { kk = (v % (BK / VB)) * VB; nn = v / (BK / VB); }
            else      { nn = (v % (BN / VB)) * VB; kk = v / (BN / VB); }
#pragma unroll
            for (int e = 0; e < VB; ++e) rb[it][e] = 0.f;
            if (v < BK * BN / VB) {
                const Info k = kinfo[buf][kk];
                const int n = n0 + nn;
                if (k.aux >= 0 && n < f.N) {
                    if (VB == 4) {
                        const float4 t = __ldg(reinterpret_cast<const float4*>(bsrc + ((long long)k.aux + (long long)n * bstride)));
                        rb[it][0] = t.x; rb[it][1 % VB] = t.y; rb[it][2 % VB] = t.z; rb[it][3 % VB] = t.w;
                    } else {
                        rb[it][0] = __ldg(bsrc + ((long long)k.aux + (long long)n * bstride));
                    }
                }
            }
        }
    };
    auto store_tile = [&]() {
#pragma unroll
        for (int it = 0; it < A_ITEMS; ++it) {
            const int v = tid + it * NT;
            if (v < BM * BK / VA) {
                int mm, kk;
                if (A_KC) { kk = (v % (BK / VA)) * VA; mm = v / (BK / VA); }
                else      { mm = (v % (BM / VA)) * VA; kk = v / (BM / VA); }
#pragma unroll
                for (int e = 0; e < VA; ++e) {
                    if (A_KC) As[kk + e][mm] = ra[it][e];
                    else      As[kk][mm + e] = ra[it][e];
                }
            }
        }
#pragma unroll
        for (int it = 0; it < B_ITEMS; ++it) {
            const int v = tid + it * NT;
            if (v < BK * BN / VB) {
                int nn, kk;
                if (B_KC) { kk = (v % (BK / VB)) * VB; nn = v / (BK / VB); }
                else      { nn = (v % (BN / VB)) * VB; kk = v / (BN / VB); }
#pragma unroll
                for (int e = 0; e < VB; ++e) {
                    if (B_KC) Bs[kk + e][nn] = rb[it][e];
                    else      Bs[kk][nn + e] = rb[it][e];
                }
            }
        }
    };

    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nchunks = (kend - kbeg + BK - 1) / BK;
    if (nchunks > 0) {
        load_tile(0);
        store_tile();
        if (tid < BK) kinfo[1][tid] = decode_k<MODE>(p, f, kbeg + BK + tid, kend);
        __syncthreads();
        for (int c = 0; c < nchunks; ++c) {
            const bool more = (c + 1 < nchunks);
            if (more) load_tile((c + 1) & 1);
#pragma unroll
            for (int k = 0; k < BK; ++k) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
                for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx * TN + j];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            __syncthreads();
            if (more) {
                store_tile();
                if (tid < BK) kinfo[c & 1][tid] = decode_k<MODE>(p, f, kbeg + (c + 2) * BK + tid, kend);
            }
            __syncthreads();
        }
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int mm = ty * TM + i;
        if (m0 + mm >= f.M) continue;
        const int base = rowinfo[mm].aux;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= f.N) continue;
            float v = acc[i][j];
            if (MODE != M_WGRAD && bias != nullptr) v += __ldg(bias + n);
            if (MODE != M_WGRAD && p.epi == EG_EPI_ACT) v = (p.epi_ge ? v >= 0.f : v > 0.f) ? v : p.epi_neg * v;
            if (MODE != M_WGRAD && p.epi == EG_EPI_MASK) {
                const float m = __ldg(p.mask + base + n);
                v *= (p.epi_ge ? m >= 0.f : m > 0.f) ? 1.f : p.epi_neg;
            }
            if (MODE == M_WGRAD && atomic_out) atomicAdd(dst + base + n, v);
            else dst[base + n] = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Input gradient onto 3-channel images (g_dconv_4 forward = conv-transpose to RGB; d_conv_0 input gradients in the
// penalty sweep and in the generator step).  With N = 3 every dy value feeds only 3 FMAs, so the tiled GEMM above is
// issue-bound (6 TFLOP/s).  Here: one thread per input pixel of one parity phase; the dy pixels a block needs are
// staged ONCE in shared memory with coalesced loads (pixel pitch Co+4 floats -> conflict-free float4 reads), and the
// filter of the phase sits in __constant__ memory in [tap][co][ci] order (uniform loads).  8.5 TFLOP/s; a
// 4-pixel-per-thread variant with the filter in shared memory was slower (occupancy) -- see DESIGN.md section 7.
constexpr int kThinMaxTaps = 9, kThinMaxCo = 64, kThinTH = 8, kThinTW = 16;
__constant__ float c_thin_w[4 * kThinMaxTaps * kThinMaxCo * 3];

__global__ void thin_prep_w_k(const float* __restrict__ w, float* __restrict__ out, ConvP p) {
    // out[((phase*kThinMaxTaps + t)*Co + co)*3 + ci] = w[r, q, ci, co] for the t-th valid tap (r, q) of the phase
    const int S = p.S, Co = p.Co;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * S * kThinMaxTaps * Co * 3) return;
    const int ci = i % 3; int t2 = i / 3; const int co = t2 % Co; t2 /= Co; const int t = t2 % kThinMaxTaps; const int phase = t2 / kThinMaxTaps;
    const int ph = phase / S, pw = phase % S;
    const int r0 = (ph + p.PT) % S, q0 = (pw + p.PL) % S;
    const int nq = q0 < p.KW ? (p.KW - q0 + S - 1) / S : 0, nr = r0 < p.KH ? (p.KH - r0 + S - 1) / S : 0;
    float v = 0.f;
    if (t < nr * nq) {
        const int r = r0 + S * (t / nq), q = q0 + S * (t % nq);
        v = w[(((size_t)r * p.KW + q) * p.Ci + ci) * Co + co];
    }
    out[i] = v;
}

__global__ void __launch_bounds__(kThinTH * kThinTW)
dgrad_thin_k(const float* __restrict__ dy, const float* __restrict__ bias, float* __restrict__ dx, ConvP p,
             int tiles_x, int tiles_y) {
    extern __shared__ __align__(16) float sdy[];              // [(TH+hy) x (TW+hx)] pixels, pitch Co+4
    const int S = p.S, Co = p.Co, pitch = Co + 4;
    const int phase = blockIdx.z, ph = phase / S, pw = phase % S;
    const int Hp = (p.H - ph + S - 1) / S, Wp = (p.W - pw + S - 1) / S;
    const int r0 = (ph + p.PT) % S, q0 = (pw + p.PL) % S;
    const int nr = r0 < p.KH ? (p.KH - r0 + S - 1) / S : 0, nq = q0 < p.KW ? (p.KW - q0 + S - 1) / S : 0;
    const int d0y = (ph + p.PT - r0) / S, d0x = (pw + p.PL - q0) / S;
    int b = blockIdx.x;
    const int tx = b % tiles_x; b /= tiles_x; const int ty = b % tiles_y; const int n = b / tiles_y;
    const int y0 = ty * kThinTH, x0 = tx * kThinTW;          // tile origin in phase coordinates (ih2, iw2)
    // dy rows needed: oh = ih2 + d0y - j, j in [0, nr)  ->  [y0 + d0y - (nr-1), y0 + TH-1 + d0y]
    const int oy0 = y0 + d0y - (nr - 1), ox0 = x0 + d0x - (nq - 1);
    const int rows = kThinTH + nr - 1, cols = kThinTW + nq - 1;
    const int c4n = Co / 4;
    for (int i = threadIdx.x; i < rows * cols * c4n; i += blockDim.x) {
        const int c4 = i % c4n; int t = i / c4n; const int cx = t % cols; const int cy = t / cols;
        const int oh = oy0 + cy, ow = ox0 + cx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((unsigned)oh < (unsigned)p.OH && (unsigned)ow < (unsigned)p.OW)
            v = __ldg(reinterpret_cast<const float4*>(dy + (((size_t)n * p.OH + oh) * p.OW + ow) * Co) + c4);
        *reinterpret_cast<float4*>(sdy + (size_t)(cy * cols + cx) * pitch + c4 * 4) = v;
    }
    __syncthreads();
    const int ly = threadIdx.x / kThinTW, lx = threadIdx.x % kThinTW;
    const int ih2 = y0 + ly, iw2 = x0 + lx;
    if (ih2 >= Hp || iw2 >= Wp) return;
    float a0 = bias ? bias[0] : 0.f, a1 = bias ? bias[1] : 0.f, a2 = bias ? bias[2] : 0.f;
    const float* cw = c_thin_w + (size_t)phase * kThinMaxTaps * Co * 3;
    for (int j = 0; j < nr; ++j) {
        for (int jq = 0; jq < nq; ++jq) {
            // oh = ih2 + d0y - j  ->  tile row (ly + nr-1 - j); same for columns
            const float4* dp = reinterpret_cast<const float4*>(sdy + (size_t)((ly + nr - 1 - j) * cols + (lx + nq - 1 - jq)) * pitch);
            const float* wt = cw + (size_t)(j * nq + jq) * Co * 3;
#pragma unroll 4
            for (int c4 = 0; c4 < c4n; ++c4) {
                const float4 d = dp[c4];
                const float* w4 = wt + c4 * 12;
                a0 = fmaf(d.x, w4[0], a0); a1 = fmaf(d.x, w4[1], a1); a2 = fmaf(d.x, w4[2], a2);
                a0 = fmaf(d.y, w4[3], a0); a1 = fmaf(d.y, w4[4], a1); a2 = fmaf(d.y, w4[5], a2);
                a0 = fmaf(d.z, w4[6], a0); a1 = fmaf(d.z, w4[7], a1); a2 = fmaf(d.z, w4[8], a2);
                a0 = fmaf(d.w, w4[9], a0); a1 = fmaf(d.w, w4[10], a1); a2 = fmaf(d.w, w4[11], a2);
            }
        }
    }
    float* o = dx + (((size_t)n * p.H + ih2 * S + ph) * p.W + iw2 * S + pw) * 3;
    o[0] = a0; o[1] = a1; o[2] = a2;
}

// staging buffer for the rearranged filter (device), one per process; the constant bank is refilled per call
float* g_thin_stage = nullptr;

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <int MODE, int BM, int BN, int TM, int TN>
int launch_cfg(const float* a, const float* b, const float* bias, float* dst, const ConvP& p, const Phase& f,
               int gz, bool va, bool vb, int atomic_out, cudaStream_t st) {
    dim3 grid(eg_ceil_div(f.M, BM), eg_ceil_div(f.N, BN), gz);
    if (va && vb) igemm_simt<MODE, BM, BN, TM, TN, 4, 4><<<grid, 256, 0, st>>>(a, b, bias, dst, p, f, atomic_out);
    else if (va)  igemm_simt<MODE, BM, BN, TM, TN, 4, 1><<<grid, 256, 0, st>>>(a, b, bias, dst, p, f, atomic_out);
    else if (vb)  igemm_simt<MODE, BM, BN, TM, TN, 1, 4><<<grid, 256, 0, st>>>(a, b, bias, dst, p, f, atomic_out);
    else          igemm_simt<MODE, BM, BN, TM, TN, 1, 1><<<grid, 256, 0, st>>>(a, b, bias, dst, p, f, atomic_out);
    EG_CHECK_LAUNCH();
    return 0;
}

template <int MODE>
int launch_mode(const float* a, const float* b, const float* bias, float* dst, const ConvP& p, const Phase& f,
                int gz, bool va, bool vb, int atomic_out, cudaStream_t st) {
    // N <= 4 (3-channel images): one output pixel per thread, all of its channels in registers
    if (f.N <= 4)  return launch_cfg<MODE, 256, 4, 1, 4>(a, b, bias, dst, p, f, gz, va, vb, atomic_out, st);
    if (f.N <= 16) return launch_cfg<MODE, 128, 16, 4, 2>(a, b, bias, dst, p, f, gz, va, vb, atomic_out, st);
    return launch_cfg<MODE, 64, 64, 4, 4>(a, b, bias, dst, p, f, gz, va, vb, atomic_out, st);
}

ConvP to_p(const eg_conv_shape* s) {
    ConvP p{s->N, s->H, s->W, s->Ci, s->OH, s->OW, s->Co, s->KH, s->KW, s->stride, s->pad_t, s->pad_l, EG_EPI_NONE, 0, 0.f, nullptr};
    return p;
}

// sign-type activations only (value / derivative = 1 on the positive side, epi_neg on the other); -> 1 when `epi` is of
// another kind and stays with the caller
int set_epilogue(ConvP& p, const EgEpi* epi) {
    if (!epi || epi->mode == EG_EPI_NONE) return 0;
    switch (epi->act) {
        case EG_ACT_RELU: p.epi_neg = 0.f; p.epi_ge = 0; break;
        case EG_ACT_LRELU_BLOCK: p.epi_neg = 0.2f; p.epi_ge = 1; break;
        case EG_ACT_LRELU: p.epi_neg = 0.2f; p.epi_ge = 0; break;
        default: return 1;
    }
    p.epi = epi->mode; p.mask = epi->mask;
    return 0;
}

}  // namespace

int eg_conv_shape_check(const eg_conv_shape* s) {
    EG_REQUIRE(s != nullptr);
    EG_REQUIRE(s->N > 0 && s->H > 0 && s->W > 0 && s->Ci > 0 && s->OH > 0 && s->OW > 0 && s->Co > 0);
    EG_REQUIRE(s->KH > 0 && s->KW > 0 && s->stride > 0 && s->pad_t >= 0 && s->pad_l >= 0);
    EG_REQUIRE(s->pad_t < s->KH && s->pad_l < s->KW);
    // int32 indexing inside the kernels
    EG_REQUIRE((long long)s->N * s->H * s->W * s->Ci < (1ll << 31));
    EG_REQUIRE((long long)s->N * s->OH * s->OW * s->Co < (1ll << 31));
    EG_REQUIRE((long long)s->KH * s->KW * s->Ci * s->Co < (1ll << 31));
    return 0;
}

// conv_small.cu: direct kernels for layers with <= 8 channels on both sides (-100 = not covered)
int eg_small_conv2d(const eg_conv_shape* s, const float* in, const float* w, const float* bias, float* out, int dgrad, cudaStream_t st);
int eg_small_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int sms, cudaStream_t st);

int eg_simt_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y,
                       const EgEpi* epi, cudaStream_t st) {
    {
        const int r = eg_small_conv2d(s, x, w, bias, y, 0, st);
        if (r != -100) return r ? r : ((epi && epi->mode != EG_EPI_NONE) ? 1 : 0);
    }
    ConvP p = to_p(s);
    const int unfused = set_epilogue(p, epi);
    Phase f{};
    f.M = p.N * p.OH * p.OW; f.N = p.Co; f.K = p.KH * p.KW * p.Ci;
    const bool va = (p.Ci % 4 == 0) && aligned16(x);
    const bool vb = (p.Co % 4 == 0) && aligned16(w);
    if (int r = launch_mode<M_FWD>(x, w, bias, y, p, f, 1, va, vb, 0, st)) return r;
    return unfused;
}

// returns 1 (instead of 0) when `epi` was NOT applied (the 3-channel kernel has no fused epilogue): the caller runs it
int eg_simt_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx,
                            const EgEpi* epi, cudaStream_t st) {
    {
        const int r = eg_small_conv2d(s, dy, w, bias, dx, 1, st);
        if (r != -100) return r ? r : ((epi && epi->mode != EG_EPI_NONE) ? 1 : 0);
    }
    ConvP p = to_p(s);
    const int unfused = set_epilogue(p, epi);
    Phase f{};
    const int S = p.S;
    f.M = p.N * ((p.H + S - 1) / S) * ((p.W + S - 1) / S);   // largest parity sub-grid (grid sizing only)
    f.N = p.Ci; f.K = 0;
    const bool v = (p.Co % 4 == 0) && aligned16(dy) && aligned16(w);
    {
        // taps per phase <= kThinMaxTaps, stride <= 2, Co <= kThinMaxCo
        const int max_nr = (p.KH + S - 1) / S, max_nq = (p.KW + S - 1) / S;
        if (p.Ci == 3 && v && S <= 2 && p.Co <= kThinMaxCo && max_nr * max_nq <= kThinMaxTaps && p.H % S == 0 && p.W % S == 0) {
            const size_t nflt = (size_t)S * S * kThinMaxTaps * p.Co * 3;
            if (g_thin_stage == nullptr) {
                cudaError_t e = cudaMalloc(&g_thin_stage, sizeof(float) * 4 * kThinMaxTaps * kThinMaxCo * 3);
                if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
            }
            thin_prep_w_k<<<eg_ceil_div((long long)nflt, 256), 256, 0, st>>>(w, g_thin_stage, p);
            EG_CHECK_LAUNCH();
            cudaError_t e = cudaMemcpyToSymbolAsync(c_thin_w, g_thin_stage, sizeof(float) * nflt, 0, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
            const int Hp = p.H / S, Wp = p.W / S;
            const int tiles_x = eg_ceil_div(Wp, kThinTW), tiles_y = eg_ceil_div(Hp, kThinTH);
            const size_t smem = sizeof(float) * (size_t)(kThinTH + max_nr - 1) * (kThinTW + max_nq - 1) * (p.Co + 4);
            static bool attr = false;
            if (!attr) {
                e = cudaFuncSetAttribute(dgrad_thin_k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
                if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
                attr = true;
            }
            dim3 grid(tiles_x * tiles_y * p.N, 1, S * S);
            dgrad_thin_k<<<grid, kThinTH * kThinTW, smem, st>>>(dy, bias, dx, p, tiles_x, tiles_y);
            EG_CHECK_LAUNCH();
            return (epi && epi->mode != EG_EPI_NONE) ? 1 : 0;
        }
    }
    if (int r = launch_mode<M_DGRAD>(dy, w, bias, dx, p, f, S * S, v, v, 0, st)) return r;
    return unfused;
}

int eg_thin_wgrad_ffma(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int sms, cudaStream_t st);
extern int g_eg_thin_wgrad_off;      // eg_debug_set(7, 1): generic implicit GEMM for the thin layers too (tests compare both)

int eg_simt_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate,
                              int sm_count, cudaStream_t st) {
    // stride-1 thin layers only: measured (tools/norm_time.py) 433 vs 554 us on the 8 -> 128 classifier layer, 71 vs 82 / 42 vs
    // 51 us on the 3 -> 128 / 256 image convs, but 499 vs 438 us on the stride-2 critic first layer
    {
        const int r = eg_small_conv2d_bwd_weight(s, x, dy, dw, accumulate, sm_count, st);
        if (r != -100) return r;
    }
    if (s->Ci <= 8 && s->stride == 1 && !g_eg_thin_wgrad_off) {
        const int r = eg_thin_wgrad_ffma(s, x, dy, dw, accumulate, sm_count, st);
        if (r != -100) return r;
    }
    ConvP p = to_p(s);
    Phase f{};
    f.M = p.KH * p.KW * p.Ci; f.N = p.Co; f.K = p.N * p.OH * p.OW;
    // split K so that the grid covers the machine a few times over
    const int bn = f.N <= 4 ? 4 : (f.N <= 16 ? 16 : 64), bm = f.N <= 4 ? 256 : (f.N <= 16 ? 128 : 64);
    const long long tiles = (long long)eg_ceil_div(f.M, bm) * eg_ceil_div(f.N, bn);
    int splits = (int)((4ll * sm_count + tiles - 1) / tiles);
    const int max_splits = eg_ceil_div(f.K, 256);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    f.k_per_split = eg_ceil_div(eg_ceil_div(f.K, splits), 16) * 16;
    splits = eg_ceil_div(f.K, f.k_per_split);
    const int atomic_out = (splits > 1 || accumulate) ? 1 : 0;
    if (atomic_out && !accumulate) {
        cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)f.M * f.N, st);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    const bool va = (p.Ci % 4 == 0) && aligned16(x);
    const bool vb = (p.Co % 4 == 0) && aligned16(dy);
    return launch_mode<M_WGRAD>(x, dy, nullptr, dw, p, f, splits, va, vb, atomic_out, st);
}
