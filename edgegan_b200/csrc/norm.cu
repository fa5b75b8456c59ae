// Instance norm (+ fused activation) forward / backward / double-backward and the split batch norm of
// the generator's h0.  All kernels are HBM/L2-bound reductions over the H*W positions of one
// (sample, 32-channel group) slab; the slab is streamed twice (stats, apply) and the second pass
// hits the 126 MB L2.
//
// Reference: nn/modules/normalization.py:10-29.
#include "common.cuh"
#include <cooperative_groups.h>

#include <initializer_list>

namespace {

constexpr int CG = 32;   // channels per block (one 128-byte line of an NHWC row)
constexpr int RY = 8;    // row-threads per channel

// sum K quantities over the RY row-threads that share a channel; result broadcast to all of them
template <int K>
__device__ __forceinline__ void reduce_cols(float (&v)[K], float (*sm)[RY][CG]) {
    const int cx = threadIdx.x, ry = threadIdx.y;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) sm[k][ry][cx] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < RY; ++r) s += sm[k][r][cx];
        v[k] = s;
    }
}

__global__ void __launch_bounds__(CG * RY)
instnorm_fwd_k(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ stats, int P, int C,
               float eps, int act) {
    __shared__ float sm[1][RY][CG];
    const int n = blockIdx.y, c = blockIdx.x * CG + threadIdx.x;
    const bool ok = c < C;
    const float* xp = x + (size_t)n * P * C + c;
    float v[1] = {0.f};
    if (ok) for (int p = threadIdx.y; p < P; p += RY) v[0] += xp[(size_t)p * C];
    reduce_cols<1>(v, sm);
    const float mean = v[0] / P;
    v[0] = 0.f;
    if (ok) for (int p = threadIdx.y; p < P; p += RY) { float d = xp[(size_t)p * C] - mean; v[0] = fmaf(d, d, v[0]); }
    reduce_cols<1>(v, sm);
    const float sd = sqrtf(v[0] / P);
    const float r = 1.f / (sd + eps);
    if (!ok) return;
    if (threadIdx.y == 0) { stats[((size_t)n * C + c) * 2] = mean; stats[((size_t)n * C + c) * 2 + 1] = sd; }
    float* yp = y + (size_t)n * P * C + c;
    for (int p = threadIdx.y; p < P; p += RY) yp[(size_t)p * C] = act_fwd(act, (xp[(size_t)p * C] - mean) * r);
}

__global__ void __launch_bounds__(CG * RY)
instnorm_bwd_k(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
               const float* __restrict__ addend, float* __restrict__ gx, int P, int C, float eps, int act) {
    __shared__ float sm[2][RY][CG];
    const int n = blockIdx.y, c = blockIdx.x * CG + threadIdx.x;
    const bool ok = c < C;
    const size_t base = (size_t)n * P * C + c;
    float mean = 0.f, sd = 1.f;
    if (ok) { mean = stats[((size_t)n * C + c) * 2]; sd = stats[((size_t)n * C + c) * 2 + 1]; }
    const float r = 1.f / (sd + eps);
    float v[2] = {0.f, 0.f};
    if (ok) for (int p = threadIdx.y; p < P; p += RY) {
        const float cc = x[base + (size_t)p * C] - mean;
        const float gn = gy[base + (size_t)p * C] * act_grad(act, cc * r);
        v[0] += gn; v[1] = fmaf(gn, cc, v[1]);
    }
    reduce_cols<2>(v, sm);
    if (!ok) return;
    const float mg = v[0] / P, q = v[1] / P;
    const float kq = r * r / sd * q;
    for (int p = threadIdx.y; p < P; p += RY) {
        const size_t i = base + (size_t)p * C;
        const float cc = x[i] - mean;
        const float gn = gy[i] * act_grad(act, cc * r);
        float o = r * (gn - mg) - kq * cc;
        if (addend != nullptr) o += addend[i];
        gx[i] = o;
    }
}

__global__ void __launch_bounds__(CG * RY)
instnorm_bwd2_k(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
                const float* __restrict__ t, float* __restrict__ out_gy, float* __restrict__ out_x, int P, int C,
                float eps, int act) {
    __shared__ float sm[5][RY][CG];
    const int n = blockIdx.y, c = blockIdx.x * CG + threadIdx.x;
    const bool ok = c < C;
    const size_t base = (size_t)n * P * C + c;
    float mean = 0.f, sd = 1.f;
    if (ok) { mean = stats[((size_t)n * C + c) * 2]; sd = stats[((size_t)n * C + c) * 2 + 1]; }
    const float r = 1.f / (sd + eps);
    // pass 1: means of gn and t ; pass 2: centred second moments (a one-pass E[t*gn] - E[t]E[gn] cancels badly
    // in fp32 and the penalty gradient amplifies that error ~1000x, see tools/parity_report.py)
    float v0[2] = {0.f, 0.f};
    if (ok) for (int p = threadIdx.y; p < P; p += RY) {
        const size_t i = base + (size_t)p * C;
        const float cc = x[i] - mean;
        v0[0] += gy[i] * act_grad(act, cc * r); v0[1] += t[i];
    }
    reduce_cols<2>(v0, reinterpret_cast<float (*)[RY][CG]>(sm));
    const float mg = v0[0] / P, mt = v0[1] / P;
    float v[3] = {0.f, 0.f, 0.f};             // sum gn*c, sum t*c, sum (t-mt)*(gn-mg)
    if (ok) for (int p = threadIdx.y; p < P; p += RY) {
        const size_t i = base + (size_t)p * C;
        const float cc = x[i] - mean;
        const float gn = gy[i] * act_grad(act, cc * r);
        const float tt = t[i];
        v[0] = fmaf(gn, cc, v[0]); v[1] = fmaf(tt, cc, v[1]); v[2] = fmaf(tt - mt, gn - mg, v[2]);
    }
    reduce_cols<3>(v, reinterpret_cast<float (*)[RY][CG]>(sm));
    if (!ok) return;
    const float q = v[0] / P, u = v[1] / P, w = v[2] / P;
    const float kap = r * r / sd;
    const float coef_c = -kap * w + (2.f * r * r * r / (sd * sd) + r * r / (sd * sd * sd)) * q * u;
    for (int p = threadIdx.y; p < P; p += RY) {
        const size_t i = base + (size_t)p * C;
        const float cc = x[i] - mean;
        const float ag = act_grad(act, cc * r);
        const float gn = gy[i] * ag;
        const float tt = t[i];
        out_gy[i] = ag * (r * (tt - mt) - kap * u * cc);
        out_x[i] = cc * coef_c - kap * u * (gn - mg) - kap * q * (tt - mt);
    }
}

// ---------------------------------------------------------------------------------------------------
// float4 variants (C % 4 == 0, 16-byte aligned): a thread owns 4 consecutive channels, 8 threads cover the 32-channel
// slab, 32 rows of the slab are in flight per block iteration -> 4x fewer load instructions, 4x the bytes in flight.
constexpr int VQ = 8, VR = 32;     // channel quads per block, row-threads

template <int K>
__device__ __forceinline__ void reduce_cols4(float4 (&v)[K], float4 (*sm)[VR][VQ]) {
    const int cx = threadIdx.x, ry = threadIdx.y;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) sm[k][ry][cx] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int r = 0; r < VR; ++r) { const float4 t = sm[k][r][cx]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
        v[k] = s;
    }
}
#define F4OP(dst, expr) { dst.x = expr(x); dst.y = expr(y); dst.z = expr(z); dst.w = expr(w); }

__global__ void __launch_bounds__(VQ * VR)
instnorm_fwd_v4(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ stats, int P, int C, float eps, int act) {
    __shared__ float4 sm[1][VR][VQ];
    const int n = blockIdx.y, c = (blockIdx.x * VQ + threadIdx.x) * 4;
    const bool ok = c < C;
    const float4* xp = reinterpret_cast<const float4*>(x + (size_t)n * P * C + c);
    const size_t pitch = C / 4;
    float4 v[1] = {make_float4(0.f, 0.f, 0.f, 0.f)};
    if (ok) for (int p = threadIdx.y; p < P; p += VR) { const float4 t = __ldg(xp + p * pitch); v[0].x += t.x; v[0].y += t.y; v[0].z += t.z; v[0].w += t.w; }
    reduce_cols4<1>(v, sm);
    const float4 mean = make_float4(v[0].x / P, v[0].y / P, v[0].z / P, v[0].w / P);
    v[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) for (int p = threadIdx.y; p < P; p += VR) {
        const float4 t = __ldg(xp + p * pitch);
        float d;
        d = t.x - mean.x; v[0].x = fmaf(d, d, v[0].x); d = t.y - mean.y; v[0].y = fmaf(d, d, v[0].y);
        d = t.z - mean.z; v[0].z = fmaf(d, d, v[0].z); d = t.w - mean.w; v[0].w = fmaf(d, d, v[0].w);
    }
    reduce_cols4<1>(v, sm);
    if (!ok) return;
    const float4 sd = make_float4(sqrtf(v[0].x / P), sqrtf(v[0].y / P), sqrtf(v[0].z / P), sqrtf(v[0].w / P));
    const float4 r = make_float4(1.f / (sd.x + eps), 1.f / (sd.y + eps), 1.f / (sd.z + eps), 1.f / (sd.w + eps));
    if (threadIdx.y == 0) {
        float* st = stats + ((size_t)n * C + c) * 2;
        st[0] = mean.x; st[1] = sd.x; st[2] = mean.y; st[3] = sd.y; st[4] = mean.z; st[5] = sd.z; st[6] = mean.w; st[7] = sd.w;
    }
    float4* yp = reinterpret_cast<float4*>(y + (size_t)n * P * C + c);
    for (int p = threadIdx.y; p < P; p += VR) {
        const float4 t = __ldg(xp + p * pitch);
        float4 o;
        o.x = act_fwd(act, (t.x - mean.x) * r.x); o.y = act_fwd(act, (t.y - mean.y) * r.y);
        o.z = act_fwd(act, (t.z - mean.z) * r.z); o.w = act_fwd(act, (t.w - mean.w) * r.w);
        yp[p * pitch] = o;
    }
}

struct Stat4 { float4 mean, sd, r; };
__device__ __forceinline__ Stat4 load_stats4(const float* stats, size_t idx, float eps) {
    const float4 a = *reinterpret_cast<const float4*>(stats + idx * 2), b = *reinterpret_cast<const float4*>(stats + idx * 2 + 4);
    Stat4 s;
    s.mean = make_float4(a.x, a.z, b.x, b.z); s.sd = make_float4(a.y, a.w, b.y, b.w);
    s.r = make_float4(1.f / (s.sd.x + eps), 1.f / (s.sd.y + eps), 1.f / (s.sd.z + eps), 1.f / (s.sd.w + eps));
    return s;
}

__global__ void __launch_bounds__(VQ * VR)
instnorm_bwd_v4(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
                const float* __restrict__ addend, float* __restrict__ gx, int P, int C, float eps, int act) {
    __shared__ float4 sm[2][VR][VQ];
    const int n = blockIdx.y, c = (blockIdx.x * VQ + threadIdx.x) * 4;
    const bool ok = c < C;
    const size_t base = (size_t)n * P * C + c, pitch = C / 4;
    const float4* xp = reinterpret_cast<const float4*>(x + base);
    const float4* gp = reinterpret_cast<const float4*>(gy + base);
    Stat4 s{};
    if (ok) s = load_stats4(stats, (size_t)n * C + c, eps);
    float4 v[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    if (ok) for (int p = threadIdx.y; p < P; p += VR) {
        const float4 t = __ldg(xp + p * pitch), g = __ldg(gp + p * pitch);
        float cc, gn;
        cc = t.x - s.mean.x; gn = g.x * act_grad(act, cc * s.r.x); v[0].x += gn; v[1].x = fmaf(gn, cc, v[1].x);
        cc = t.y - s.mean.y; gn = g.y * act_grad(act, cc * s.r.y); v[0].y += gn; v[1].y = fmaf(gn, cc, v[1].y);
        cc = t.z - s.mean.z; gn = g.z * act_grad(act, cc * s.r.z); v[0].z += gn; v[1].z = fmaf(gn, cc, v[1].z);
        cc = t.w - s.mean.w; gn = g.w * act_grad(act, cc * s.r.w); v[0].w += gn; v[1].w = fmaf(gn, cc, v[1].w);
    }
    reduce_cols4<2>(v, sm);
    if (!ok) return;
    const float4 mg = make_float4(v[0].x / P, v[0].y / P, v[0].z / P, v[0].w / P);
    const float4 kq = make_float4(s.r.x * s.r.x / s.sd.x * (v[1].x / P), s.r.y * s.r.y / s.sd.y * (v[1].y / P),
                                  s.r.z * s.r.z / s.sd.z * (v[1].z / P), s.r.w * s.r.w / s.sd.w * (v[1].w / P));
    const float4* ap = addend ? reinterpret_cast<const float4*>(addend + base) : nullptr;
    float4* op = reinterpret_cast<float4*>(gx + base);
    for (int p = threadIdx.y; p < P; p += VR) {
        const float4 t = __ldg(xp + p * pitch), g = __ldg(gp + p * pitch);
        float4 o; float cc, gn;
        cc = t.x - s.mean.x; gn = g.x * act_grad(act, cc * s.r.x); o.x = s.r.x * (gn - mg.x) - kq.x * cc;
        cc = t.y - s.mean.y; gn = g.y * act_grad(act, cc * s.r.y); o.y = s.r.y * (gn - mg.y) - kq.y * cc;
        cc = t.z - s.mean.z; gn = g.z * act_grad(act, cc * s.r.z); o.z = s.r.z * (gn - mg.z) - kq.z * cc;
        cc = t.w - s.mean.w; gn = g.w * act_grad(act, cc * s.r.w); o.w = s.r.w * (gn - mg.w) - kq.w * cc;
        if (ap) { const float4 a = __ldg(ap + p * pitch); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
        op[p * pitch] = o;
    }
}

// ---------------------------------------------------------------------------------------------------
// Shared-memory-resident variants: the (sample, channel-group) slab is read from HBM exactly ONCE, parked in shared
// memory (as the centred value / masked cotangent the later passes need), and every further pass runs out of shared
// memory.  The multi-pass kernels above re-read the slab from L2 two to three times per tensor, which made the norms
// run at ~22 % of the HBM roofline (r01 launch list); here the traffic is the algorithmic minimum: fwd 4 + 4 B/element,
// bwd 8 + 4 (+4 addend), bwd2 12 + 8.  Block = COLS float4 columns x (512 / COLS) rows; the channel-group width
// (32 / 16 / 8 channels = 128 / 64 / 32-byte rows) is chosen so that the slab(s) fit in 200 KB.
constexpr int kSmThreads = 512;
constexpr size_t kSmCap = 200 * 1024;

__device__ __forceinline__ float4 f4z() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 shfl_xor4(float4 v, int o) {
    return make_float4(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o),
                       __shfl_xor_sync(0xffffffffu, v.z, o), __shfl_xor_sync(0xffffffffu, v.w, o));
}
// sum of v over all rows of the block, per float4 column; result in every thread.  red: [16 warps][8] float4
__device__ __forceinline__ float4 colsum4(float4 v, float4 (*red)[8], int cols, int tx) {
    for (int o = cols; o < 32; o <<= 1) { const float4 t = shfl_xor4(v, o); v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w; }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    __syncthreads();
    if (lane < cols) red[warp][lane] = v;
    __syncthreads();
    float4 s = f4z();
    const int nwarps = blockDim.x >> 5;
    for (int w = 0; w < nwarps; ++w) { const float4 t = red[w][tx]; s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w; }
    return s;
}

// The same over a thread-block CLUSTER of CL blocks that split the pixels of one (sample, channel group) slab: every
// block publishes its K partial sums in its own shared memory, one cluster barrier, every block adds the CL partials in
// rank order (so all blocks hold bit-identical sums).  Used for the slabs one block cannot hold below 72 KB (64x64-pixel
// maps and larger, see sm_pick).  `part` must not be reused by a later exchange (no second barrier).
template <int K>
__device__ __forceinline__ void clsum4(float4 (&v)[K], float4 (*red)[8], float4 (*part)[8], int cols, int tx, int ty, int CL) {
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = colsum4(v[k], red, cols, tx);
    if (CL == 1) return;
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    if (ty == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) part[k][tx] = v[k];
    }
    cl.sync();
    if (ty == 0) {                                           // one row of threads reads the peers, the block shares the totals
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = f4z();
        for (int r = 0; r < CL; ++r) {
            const float4* rp = reinterpret_cast<const float4*>(cl.map_shared_rank(&part[0][0], r));
#pragma unroll
            for (int k = 0; k < K; ++k) { const float4 t = rp[k * 8 + tx]; v[k].x += t.x; v[k].y += t.y; v[k].z += t.z; v[k].w += t.w; }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) red[k][tx] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = red[k][tx];
}
__device__ __forceinline__ void cl_exit(int CL) {            // no block may leave while a peer can still read its partials
    if (CL > 1) cooperative_groups::this_cluster().sync();
}

__global__ void __launch_bounds__(kSmThreads)
instnorm_fwd_sm(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ stats, int Pall, int C, float eps,
                int act, int cols, int CL) {
    extern __shared__ float4 slab[];                         // [P][cols], P = this block's share of the Pall pixels
    __shared__ float4 red[kSmThreads / 32][8];
    __shared__ float4 part[2][8];
    const int rows = blockDim.x / cols, tx = threadIdx.x % cols, ty = threadIdx.x / cols;
    const int n = blockIdx.y, c = ((blockIdx.x / CL) * cols + tx) * 4;
    const int P = Pall / CL, p_off = (blockIdx.x % CL) * P;
    const size_t pitch = C / 4;
    const float4* xp = reinterpret_cast<const float4*>(x + ((size_t)n * Pall + p_off) * C + c);
    float4 a[1] = {f4z()};
    // four independent 16-byte loads in flight per thread before the first use (a one-load-per-iteration loop left the
    // kernel latency-bound at ~25 % of the HBM roofline)
    int p = ty;
    for (; p + 3 * rows < P; p += 4 * rows) {
        const float4 t0 = __ldg(xp + (size_t)p * pitch), t1 = __ldg(xp + (size_t)(p + rows) * pitch);
        const float4 t2 = __ldg(xp + (size_t)(p + 2 * rows) * pitch), t3 = __ldg(xp + (size_t)(p + 3 * rows) * pitch);
        slab[p * cols + tx] = t0; slab[(p + rows) * cols + tx] = t1; slab[(p + 2 * rows) * cols + tx] = t2; slab[(p + 3 * rows) * cols + tx] = t3;
        a[0].x += (t0.x + t1.x) + (t2.x + t3.x); a[0].y += (t0.y + t1.y) + (t2.y + t3.y);
        a[0].z += (t0.z + t1.z) + (t2.z + t3.z); a[0].w += (t0.w + t1.w) + (t2.w + t3.w);
    }
    for (; p < P; p += rows) {
        const float4 t = __ldg(xp + (size_t)p * pitch);
        slab[p * cols + tx] = t;
        a[0].x += t.x; a[0].y += t.y; a[0].z += t.z; a[0].w += t.w;
    }
    clsum4<1>(a, red, part, cols, tx, ty, CL);
    const float4 mean = make_float4(a[0].x / Pall, a[0].y / Pall, a[0].z / Pall, a[0].w / Pall);
    a[0] = f4z();
    for (int p = ty; p < P; p += rows) {
        const float4 t = slab[p * cols + tx];
        float d;
        d = t.x - mean.x; a[0].x = fmaf(d, d, a[0].x); d = t.y - mean.y; a[0].y = fmaf(d, d, a[0].y);
        d = t.z - mean.z; a[0].z = fmaf(d, d, a[0].z); d = t.w - mean.w; a[0].w = fmaf(d, d, a[0].w);
    }
    clsum4<1>(a, red, part + 1, cols, tx, ty, CL);
    const float4 sd = make_float4(sqrtf(a[0].x / Pall), sqrtf(a[0].y / Pall), sqrtf(a[0].z / Pall), sqrtf(a[0].w / Pall));
    const float4 r = make_float4(1.f / (sd.x + eps), 1.f / (sd.y + eps), 1.f / (sd.z + eps), 1.f / (sd.w + eps));
    if (ty == 0 && p_off == 0) {
        float* st = stats + ((size_t)n * C + c) * 2;
        *reinterpret_cast<float4*>(st) = make_float4(mean.x, sd.x, mean.y, sd.y);
        *reinterpret_cast<float4*>(st + 4) = make_float4(mean.z, sd.z, mean.w, sd.w);
    }
    float4* yp = reinterpret_cast<float4*>(y + ((size_t)n * Pall + p_off) * C + c);
#pragma unroll 4
    for (int p = ty; p < P; p += rows) {
        const float4 t = slab[p * cols + tx];
        yp[p * pitch] = make_float4(act_fwd(act, (t.x - mean.x) * r.x), act_fwd(act, (t.y - mean.y) * r.y),
                                    act_fwd(act, (t.z - mean.z) * r.z), act_fwd(act, (t.w - mean.w) * r.w));
    }
    cl_exit(CL);
}

__global__ void __maxnreg__(80)        // 80 registers (ptxas: no spills): three 256-thread blocks per SM instead of two
instnorm_bwd_sm(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
                const float* __restrict__ addend, float* __restrict__ gx, int Pall, int C, float eps, int act, int cols, int CL) {
    extern __shared__ float4 slab[];                         // [2][P][cols]: centred x, masked cotangent
    __shared__ float4 red[kSmThreads / 32][8];
    __shared__ float4 part[2][8];
    const int rows = blockDim.x / cols, tx = threadIdx.x % cols, ty = threadIdx.x / cols;
    const int n = blockIdx.y, c = ((blockIdx.x / CL) * cols + tx) * 4;
    const int P = Pall / CL, p_off = (blockIdx.x % CL) * P;
    const size_t base = ((size_t)n * Pall + p_off) * C + c, pitch = C / 4;
    const float4* xp = reinterpret_cast<const float4*>(x + base);
    const float4* gp = reinterpret_cast<const float4*>(gy + base);
    float4* scc = slab;
    float4* sgn = slab + (size_t)P * cols;
    const Stat4 s = load_stats4(stats, (size_t)n * C + c, eps);
    float4 vv[2] = {f4z(), f4z()};
    float4 &v0 = vv[0], &v1 = vv[1];
    auto row = [&](int p, const float4 t, const float4 g) {
        float4 cc, gn;
        cc.x = t.x - s.mean.x; gn.x = g.x * act_grad(act, cc.x * s.r.x); v0.x += gn.x; v1.x = fmaf(gn.x, cc.x, v1.x);
        cc.y = t.y - s.mean.y; gn.y = g.y * act_grad(act, cc.y * s.r.y); v0.y += gn.y; v1.y = fmaf(gn.y, cc.y, v1.y);
        cc.z = t.z - s.mean.z; gn.z = g.z * act_grad(act, cc.z * s.r.z); v0.z += gn.z; v1.z = fmaf(gn.z, cc.z, v1.z);
        cc.w = t.w - s.mean.w; gn.w = g.w * act_grad(act, cc.w * s.r.w); v0.w += gn.w; v1.w = fmaf(gn.w, cc.w, v1.w);
        scc[p * cols + tx] = cc; sgn[p * cols + tx] = gn;
    };
    int p = ty;
    for (; p + 3 * rows < P; p += 4 * rows) {                // eight independent 16-byte loads in flight per thread
        const float4 t0 = __ldg(xp + (size_t)p * pitch), g0 = __ldg(gp + (size_t)p * pitch);
        const float4 t1 = __ldg(xp + (size_t)(p + rows) * pitch), g1 = __ldg(gp + (size_t)(p + rows) * pitch);
        const float4 t2 = __ldg(xp + (size_t)(p + 2 * rows) * pitch), g2 = __ldg(gp + (size_t)(p + 2 * rows) * pitch);
        const float4 t3 = __ldg(xp + (size_t)(p + 3 * rows) * pitch), g3 = __ldg(gp + (size_t)(p + 3 * rows) * pitch);
        row(p, t0, g0); row(p + rows, t1, g1); row(p + 2 * rows, t2, g2); row(p + 3 * rows, t3, g3);
    }
    for (; p < P; p += rows) row(p, __ldg(xp + (size_t)p * pitch), __ldg(gp + (size_t)p * pitch));
    clsum4<2>(vv, red, part, cols, tx, ty, CL);
    const float4 mg = make_float4(v0.x / Pall, v0.y / Pall, v0.z / Pall, v0.w / Pall);
    const float4 kq = make_float4(s.r.x * s.r.x / s.sd.x * (v1.x / Pall), s.r.y * s.r.y / s.sd.y * (v1.y / Pall),
                                  s.r.z * s.r.z / s.sd.z * (v1.z / Pall), s.r.w * s.r.w / s.sd.w * (v1.w / Pall));
    const float4* ap = addend ? reinterpret_cast<const float4*>(addend + base) : nullptr;
    float4* op = reinterpret_cast<float4*>(gx + base);
#pragma unroll 4
    for (int p = ty; p < P; p += rows) {
        const float4 cc = scc[p * cols + tx], gn = sgn[p * cols + tx];
        float4 o = make_float4(s.r.x * (gn.x - mg.x) - kq.x * cc.x, s.r.y * (gn.y - mg.y) - kq.y * cc.y,
                               s.r.z * (gn.z - mg.z) - kq.z * cc.z, s.r.w * (gn.w - mg.w) - kq.w * cc.w);
        if (ap) { const float4 a = __ldg(ap + p * pitch); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
        op[p * pitch] = o;
    }
    cl_exit(CL);
}

// second-order term of the gradient penalty (same formulas as instnorm_bwd2_k: centred second moments)
__global__ void __launch_bounds__(kSmThreads)     // (its three slabs leave room for two blocks per SM at most: no register cap)
instnorm_bwd2_sm(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
                 const float* __restrict__ t, float* __restrict__ out_gy, float* __restrict__ out_x, int Pall, int C,
                 float eps, int act, int cols, int CL) {
    extern __shared__ float4 slab[];                         // [3][P][cols]: centred x, masked cotangent, tangent
    __shared__ float4 red[kSmThreads / 32][8];
    __shared__ float4 part[5][8];
    const int rows = blockDim.x / cols, tx = threadIdx.x % cols, ty = threadIdx.x / cols;
    const int n = blockIdx.y, c = ((blockIdx.x / CL) * cols + tx) * 4;
    const int P = Pall / CL, p_off = (blockIdx.x % CL) * P;
    const size_t base = ((size_t)n * Pall + p_off) * C + c, pitch = C / 4;
    const float4* xp = reinterpret_cast<const float4*>(x + base);
    const float4* gp = reinterpret_cast<const float4*>(gy + base);
    const float4* tp = reinterpret_cast<const float4*>(t + base);
    float4* scc = slab;
    float4* sgn = slab + (size_t)P * cols;
    float4* stt = slab + (size_t)2 * P * cols;
    const Stat4 s = load_stats4(stats, (size_t)n * C + c, eps);
    float4 aa[2] = {f4z(), f4z()};
    float4 &a0 = aa[0], &a1 = aa[1];
    auto row = [&](int p, const float4 xv, const float4 g, const float4 tt) {
        float4 cc, gn;
        cc.x = xv.x - s.mean.x; gn.x = g.x * act_grad(act, cc.x * s.r.x);
        cc.y = xv.y - s.mean.y; gn.y = g.y * act_grad(act, cc.y * s.r.y);
        cc.z = xv.z - s.mean.z; gn.z = g.z * act_grad(act, cc.z * s.r.z);
        cc.w = xv.w - s.mean.w; gn.w = g.w * act_grad(act, cc.w * s.r.w);
        a0.x += gn.x; a0.y += gn.y; a0.z += gn.z; a0.w += gn.w;
        a1.x += tt.x; a1.y += tt.y; a1.z += tt.z; a1.w += tt.w;
        scc[p * cols + tx] = cc; sgn[p * cols + tx] = gn; stt[p * cols + tx] = tt;
    };
    int p = ty;
    for (; p + 2 * rows < P; p += 3 * rows) {                // nine independent 16-byte loads in flight per thread
        const float4 x0 = __ldg(xp + (size_t)p * pitch), g0 = __ldg(gp + (size_t)p * pitch), t0 = __ldg(tp + (size_t)p * pitch);
        const float4 x1 = __ldg(xp + (size_t)(p + rows) * pitch), g1 = __ldg(gp + (size_t)(p + rows) * pitch), t1 = __ldg(tp + (size_t)(p + rows) * pitch);
        const float4 x2 = __ldg(xp + (size_t)(p + 2 * rows) * pitch), g2 = __ldg(gp + (size_t)(p + 2 * rows) * pitch), t2 = __ldg(tp + (size_t)(p + 2 * rows) * pitch);
        row(p, x0, g0, t0); row(p + rows, x1, g1, t1); row(p + 2 * rows, x2, g2, t2);
    }
    for (; p < P; p += rows) row(p, __ldg(xp + (size_t)p * pitch), __ldg(gp + (size_t)p * pitch), __ldg(tp + (size_t)p * pitch));
    clsum4<2>(aa, red, part, cols, tx, ty, CL);
    const float4 mg = make_float4(a0.x / Pall, a0.y / Pall, a0.z / Pall, a0.w / Pall), mt = make_float4(a1.x / Pall, a1.y / Pall, a1.z / Pall, a1.w / Pall);
    float4 bb[3] = {f4z(), f4z(), f4z()};                     // sum gn*c, sum t*c, sum (t-mt)*(gn-mg)
    float4 &b0 = bb[0], &b1 = bb[1], &b2 = bb[2];
    for (int p = ty; p < P; p += rows) {
        const float4 cc = scc[p * cols + tx], gn = sgn[p * cols + tx], tt = stt[p * cols + tx];
        b0.x = fmaf(gn.x, cc.x, b0.x); b1.x = fmaf(tt.x, cc.x, b1.x); b2.x = fmaf(tt.x - mt.x, gn.x - mg.x, b2.x);
        b0.y = fmaf(gn.y, cc.y, b0.y); b1.y = fmaf(tt.y, cc.y, b1.y); b2.y = fmaf(tt.y - mt.y, gn.y - mg.y, b2.y);
        b0.z = fmaf(gn.z, cc.z, b0.z); b1.z = fmaf(tt.z, cc.z, b1.z); b2.z = fmaf(tt.z - mt.z, gn.z - mg.z, b2.z);
        b0.w = fmaf(gn.w, cc.w, b0.w); b1.w = fmaf(tt.w, cc.w, b1.w); b2.w = fmaf(tt.w - mt.w, gn.w - mg.w, b2.w);
    }
    clsum4<3>(bb, red, part + 2, cols, tx, ty, CL);
    float4 kap, ku, kqv, coef;
#define EG_B2C(f) { const float q = b0.f / Pall, u = b1.f / Pall, w = b2.f / Pall, r = s.r.f, sd = s.sd.f; kap.f = r * r / sd; ku.f = kap.f * u; \
                    kqv.f = kap.f * q; coef.f = -kap.f * w + (2.f * r * r * r / (sd * sd) + r * r / (sd * sd * sd)) * q * u; }
    EG_B2C(x) EG_B2C(y) EG_B2C(z) EG_B2C(w)
#undef EG_B2C
    float4* og = reinterpret_cast<float4*>(out_gy + base);
    float4* ox = reinterpret_cast<float4*>(out_x + base);
    for (int p = ty; p < P; p += rows) {
        const float4 cc = scc[p * cols + tx], gn = sgn[p * cols + tx], tt = stt[p * cols + tx];
        float4 o1, o2;
#define EG_B2O(f) { const float ag = act_grad(act, cc.f * s.r.f); o1.f = ag * (s.r.f * (tt.f - mt.f) - ku.f * cc.f); \
                    o2.f = cc.f * coef.f - ku.f * (gn.f - mg.f) - kqv.f * (tt.f - mt.f); }
        EG_B2O(x) EG_B2O(y) EG_B2O(z) EG_B2O(w)
#undef EG_B2O
        og[p * pitch] = o1; ox[p * pitch] = o2;
    }
    cl_exit(CL);
}

// ---------------------------------------------------------------------------------------------------
// Streaming two-kernel backward for large slabs: (A) per-(sample, channel) sums of gn and gn*c by a grid over
// (channel group, pixel chunk, sample) with a red.global.add of 8 floats per block column, (B) a plain element-wise apply.
// 5 tensor passes instead of 3.  Built to test whether two streaming kernels beat the one-block-per-slab kernel above, which
// serialises load / reduce / store phases inside few resident blocks (2.2-2.4 TB/s on 32x32-pixel slabs): they do not --
// 281 vs 252 us on [384, 32x32, 128], 102 vs 42 us on [384, 8x8, 512] (tools/norm_time.py) -- so this path is OFF by default.
constexpr int kRedRows = 256;      // pixel rows per block of kernel A

__global__ void __launch_bounds__(256)
instnorm_bwd_reduce_k(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
                      float* __restrict__ sums, int P, int C, float eps, int act) {
    __shared__ float4 red[8][8];
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;          // 8 float4 columns x 32 rows
    const int n = blockIdx.z, c = (blockIdx.x * 8 + tx) * 4;
    const int p0 = blockIdx.y * kRedRows, p1 = min(P, p0 + kRedRows);
    const size_t base = (size_t)n * P * C + c, pitch = C / 4;
    const float4* xp = reinterpret_cast<const float4*>(x + base);
    const float4* gp = reinterpret_cast<const float4*>(gy + base);
    const Stat4 s = load_stats4(stats, (size_t)n * C + c, eps);
    float4 v0 = f4z(), v1 = f4z();
#pragma unroll 4
    for (int p = p0 + ty; p < p1; p += 32) {
        const float4 t = __ldg(xp + (size_t)p * pitch), g = __ldg(gp + (size_t)p * pitch);
        float cc, gn;
        cc = t.x - s.mean.x; gn = g.x * act_grad(act, cc * s.r.x); v0.x += gn; v1.x = fmaf(gn, cc, v1.x);
        cc = t.y - s.mean.y; gn = g.y * act_grad(act, cc * s.r.y); v0.y += gn; v1.y = fmaf(gn, cc, v1.y);
        cc = t.z - s.mean.z; gn = g.z * act_grad(act, cc * s.r.z); v0.z += gn; v1.z = fmaf(gn, cc, v1.z);
        cc = t.w - s.mean.w; gn = g.w * act_grad(act, cc * s.r.w); v0.w += gn; v1.w = fmaf(gn, cc, v1.w);
    }
    v0 = colsum4(v0, red, 8, tx);
    v1 = colsum4(v1, red, 8, tx);
    if (ty == 0) {
        float* o = sums + ((size_t)n * C + c) * 2;               // [n][c][2] = (sum gn, sum gn*c)
        atomicAdd(o + 0, v0.x); atomicAdd(o + 1, v1.x); atomicAdd(o + 2, v0.y); atomicAdd(o + 3, v1.y);
        atomicAdd(o + 4, v0.z); atomicAdd(o + 5, v1.z); atomicAdd(o + 6, v0.w); atomicAdd(o + 7, v1.w);
    }
}

__global__ void __launch_bounds__(256)
instnorm_bwd_apply_k(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
                     const float* __restrict__ addend, const float* __restrict__ sums, float* __restrict__ gx, long long n4,
                     int P, int C, float eps, int act) {
    const int c4 = C / 4;
    const float invP = 1.f / P;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(i % c4);
        const long long n = i / ((long long)c4 * P);
        const size_t sidx = (size_t)n * C + cq * 4;
        const Stat4 s = load_stats4(stats, sidx, eps);
        const float4 a = __ldg(reinterpret_cast<const float4*>(sums + sidx * 2)), b = __ldg(reinterpret_cast<const float4*>(sums + sidx * 2 + 4));
        const float4 t = __ldg(reinterpret_cast<const float4*>(x) + i), g = __ldg(reinterpret_cast<const float4*>(gy) + i);
        float4 o; float cc, gn;
        cc = t.x - s.mean.x; gn = g.x * act_grad(act, cc * s.r.x); o.x = s.r.x * (gn - a.x * invP) - s.r.x * s.r.x / s.sd.x * (a.y * invP) * cc;
        cc = t.y - s.mean.y; gn = g.y * act_grad(act, cc * s.r.y); o.y = s.r.y * (gn - a.z * invP) - s.r.y * s.r.y / s.sd.y * (a.w * invP) * cc;
        cc = t.z - s.mean.z; gn = g.z * act_grad(act, cc * s.r.z); o.z = s.r.z * (gn - b.x * invP) - s.r.z * s.r.z / s.sd.z * (b.y * invP) * cc;
        cc = t.w - s.mean.w; gn = g.w * act_grad(act, cc * s.r.w); o.w = s.r.w * (gn - b.z * invP) - s.r.w * s.r.w / s.sd.w * (b.w * invP) * cc;
        if (addend != nullptr) { const float4 ad = __ldg(reinterpret_cast<const float4*>(addend) + i); o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w; }
        reinterpret_cast<float4*>(gx)[i] = o;
    }
}

// channel-group width (in float4 columns) such that `tensors` slabs of P rows fit in shared memory; 0: use the
// multi-pass kernels
// Preference: the widest group whose slabs stay below 64 KB (>= 3 blocks per SM, so that one block's load phase overlaps
// another's reduction / store phases -- with one 128 KB block per SM the r02 capture showed 25 % of the HBM roofline),
// down to 32-byte rows;
// above that whatever still fits in 200 KB.
int g_in_cluster_off = 0;         // eg_norm_debug(-2): one block per slab only (the tests compare both)
struct SmPick { int cols, cl; };
// (row width in float4 columns, cluster size); cols = 0: use the multi-pass kernels.  Measured (tools/norm_time.py, cold
// inputs): for slabs that fit one block below 72 KB at some width (down to 32-byte rows) one block per slab is as fast or
// faster than a cluster (32x32 maps: backward 254 us against 361 us split over 4 blocks -- the cluster barriers and the
// distributed-shared-memory exchange cost more than the wider rows give), so clusters serve the slabs that do not fit:
// 64x64 maps and larger (forward 165 -> 115 us, backward 234 -> 165 us on [128, 64x64, 64]; the second-order kernel
// 1315 -> 356 us, it had fallen back to the multi-pass kernel).
SmPick sm_pick(int P, int C, int tensors) {
    const size_t soft = 72 * 1024;
    for (int cols = 8; cols >= 2; cols >>= 1)                                  // one block per slab, >= 3 blocks per SM
        if (C % (cols * 4) == 0 && (size_t)P * cols * 16 * tensors <= soft) return SmPick{cols, 1};
    if (!g_in_cluster_off && P >= 4096)
        for (int cols = 8; cols >= 2; cols >>= 1) {                            // slab split over 2 / 4 / 8 blocks
            if (C % (cols * 4)) continue;
            for (int cl = 2; cl <= 8; cl <<= 1) {
                if (P % cl || P / cl < 512) break;
                if ((size_t)(P / cl) * cols * 16 * tensors <= soft) return SmPick{cols, cl};
            }
        }
    for (int cols = 8; cols >= 2; cols >>= 1)                                  // one big block per SM
        if (C % (cols * 4) == 0 && (size_t)P * cols * 16 * tensors <= kSmCap) return SmPick{cols, 1};
    if (!g_in_cluster_off && P >= 4096)
        for (int cols = 8; cols >= 2; cols >>= 1)
            if (C % (cols * 4) == 0 && P % 8 == 0 && (size_t)(P / 8) * cols * 16 * tensors <= kSmCap) return SmPick{cols, 8};
    return SmPick{0, 1};
}
template <typename... KA, typename... A>
int launch_sm(void (*kern)(KA...), int groups, int N, int threads, size_t smem, int CL, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(groups * CL), (unsigned)N); cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = CL > 1 ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    return 0;
}
// threads per block: about four rows per thread (small slabs are bound by the block-wide reductions, not by bytes), at
// most 256 threads whatever the row width (rows = 256 / cols): with 80 registers three such blocks share an SM, while a
// 512-thread block is alone on it.  Measured, same position in the sweep (tools/norm_time.py with ROWS=1; the sweep's
// first configuration is favoured by ~30 %, so only like positions are compared): [384, 16x16, 256] forward 75 -> 53 us,
// backward 130 -> 88 us with 256 instead of 512 threads; [384, 32x32, 128] (32-byte rows) backward 275 -> 257 us with 256
// instead of 128 threads.
int g_sm_threads_target = 256;
int sm_threads(int P, int cols) {
    int rows = 4;
    while (rows * cols < g_sm_threads_target && rows * 4 < P) rows <<= 1;
    int t = rows * cols;
    if (t < 64) t = 64;
    if (t > kSmThreads) t = kSmThreads;
    return t;
}
int g_in_stream = -1;             // eg_norm_debug: -1 never stream (default: measured slower at every size), 0 threshold
                                  // P >= 512, > 0 threshold in pixels
bool g_sm_attr = false;
int sm_attrs() {
    if (g_sm_attr) return 0;
    cudaError_t e = cudaFuncSetAttribute(instnorm_fwd_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmCap);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(instnorm_bwd_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmCap);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(instnorm_bwd2_sm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmCap);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    g_sm_attr = true;
    return 0;
}
bool al16all(std::initializer_list<const void*> ps) {
    for (const void* p : ps) if (reinterpret_cast<uintptr_t>(p) & 15) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------------
// batch norm over rows of x[R, C]

__global__ void __launch_bounds__(CG * RY)
bn_stats_k(const float* __restrict__ x, float* __restrict__ sums, int R, int C) {
    __shared__ float sm[2][RY][CG];
    const int c = blockIdx.x * CG + threadIdx.x;
    const bool ok = c < C;
    float v[2] = {0.f, 0.f};
    if (ok) for (int p = threadIdx.y; p < R; p += RY) { const float a = x[(size_t)p * C + c]; v[0] += a; v[1] = fmaf(a, a, v[1]); }
    reduce_cols<2>(v, sm);
    if (ok && threadIdx.y == 0) { sums[c] = v[0]; sums[C + c] = v[1]; }
}

__device__ __forceinline__ void bn_coeffs(const float* sums, float count, int C, int c, float eps, float& mu, float& rstd) {
    mu = sums[c] / count;
    const float var = fmaxf(sums[C + c] / count - mu * mu, 0.f);
    rstd = rsqrtf(var + eps);
}

__global__ void bn_apply_k(const float* __restrict__ x, const float* __restrict__ sums, float count,
                           const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                           long long total, int C, float eps, int act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    float mu, rstd; bn_coeffs(sums, count, C, c, eps, mu, rstd);
    y[i] = act_fwd(act, gamma[c] * (x[i] - mu) * rstd + beta[c]);
}

__global__ void __launch_bounds__(CG * RY)
bn_bwd_reduce_k(const float* __restrict__ x, const float* __restrict__ sums, float count,
                const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ gy,
                float* __restrict__ red, int R, int C, float eps, int act) {
    __shared__ float sm[2][RY][CG];
    const int c = blockIdx.x * CG + threadIdx.x;
    const bool ok = c < C;
    float mu = 0.f, rstd = 1.f, g = 1.f, b = 0.f;
    if (ok) { bn_coeffs(sums, count, C, c, eps, mu, rstd); g = gamma[c]; b = beta[c]; }
    float v[2] = {0.f, 0.f};
    if (ok) for (int p = threadIdx.y; p < R; p += RY) {
        const size_t i = (size_t)p * C + c;
        const float xh = (x[i] - mu) * rstd;
        const float gp = gy[i] * act_grad(act, g * xh + b);
        v[0] += gp; v[1] = fmaf(gp, xh, v[1]);
    }
    reduce_cols<2>(v, sm);
    if (ok && threadIdx.y == 0) { red[c] = v[0]; red[C + c] = v[1]; }
}

__global__ void bn_bwd_apply_k(const float* __restrict__ x, const float* __restrict__ sums, float count,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ gy, const float* __restrict__ red, float* __restrict__ gx,
                               long long total, int C, float eps, int act) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C);
    float mu, rstd; bn_coeffs(sums, count, C, c, eps, mu, rstd);
    const float g = gamma[c];
    const float xh = (x[i] - mu) * rstd;
    const float gp = gy[i] * act_grad(act, g * xh + beta[c]);
    gx[i] = g * rstd * (gp - red[c] / count - xh * red[C + c] / count);
}

}  // namespace

int eg_tc_scratch(cudaStream_t st, int slot, size_t bytes, float** out);

extern "C" {

/* development knob (tools/norm_time.py): pixel count from which the instance-norm backward streams in two kernels */
int eg_norm_debug(int value) {
    if (value == -2) g_in_cluster_off = 1;                   // one block per slab only
    else if (value == -3) g_in_cluster_off = 0;
    else if (value <= -10) g_sm_threads_target = -value;     // -256 (default) / -128 / -512: threads per block
    else g_in_stream = value;
    return 0;
}

int eg_instnorm_fwd(const float* x, float* y, float* stats, int N, int P, int C, float eps, int act, void* stream) {
    EG_REQUIRE(x && y && stats && N > 0 && P > 0 && C > 0 && N <= 65535);
    if (const SmPick k = al16all({x, y, stats}) ? sm_pick(P, C, 1) : SmPick{0, 1}; k.cols) {
        if (int r = sm_attrs()) return r;
        if (int r = launch_sm(instnorm_fwd_sm, C / (k.cols * 4), N, sm_threads(P / k.cl, k.cols), (size_t)(P / k.cl) * k.cols * 16, k.cl,
                              (cudaStream_t)stream, x, y, stats, P, C, eps, act, k.cols, k.cl)) return r;
        EG_CHECK_LAUNCH();
        return 0;
    }
    if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(stats)) & 15) == 0) {
        dim3 grid(eg_ceil_div(C, VQ * 4), N), block(VQ, VR);
        instnorm_fwd_v4<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, stats, P, C, eps, act);
        EG_CHECK_LAUNCH();
        return 0;
    }
    dim3 grid(eg_ceil_div(C, CG), N), block(CG, RY);
    instnorm_fwd_k<<<grid, block, 0, (cudaStream_t)stream>>>(x, y, stats, P, C, eps, act);
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_instnorm_bwd(const float* x, const float* stats, const float* gy, const float* addend, float* gx, int N,
                    int P, int C, float eps, int act, void* stream) {
    EG_REQUIRE(x && stats && gy && gx && N > 0 && P > 0 && C > 0 && N <= 65535);
    if (g_in_stream >= 0 && P >= (g_in_stream > 0 ? g_in_stream : 512) && C % 32 == 0 && al16all({x, stats, gy, addend, gx})) {
        float* sums = nullptr;
        if (int r = eg_tc_scratch((cudaStream_t)stream, 4, sizeof(float) * (size_t)N * C * 2, &sums)) return r;
        cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(float) * (size_t)N * C * 2, (cudaStream_t)stream);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
        instnorm_bwd_reduce_k<<<dim3(C / 32, eg_ceil_div(P, kRedRows), N), 256, 0, (cudaStream_t)stream>>>(x, stats, gy, sums, P, C, eps, act);
        EG_CHECK_LAUNCH();
        const long long n4 = (long long)N * P * C / 4;
        long long blocks = (n4 + 256 * 4 - 1) / (256 * 4);
        if (blocks > 148 * 16) blocks = 148 * 16;
        instnorm_bwd_apply_k<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, stats, gy, addend, sums, gx, n4, P, C, eps, act);
        EG_CHECK_LAUNCH();
        return 0;
    }
    if (const SmPick k = al16all({x, stats, gy, addend, gx}) ? sm_pick(P, C, 2) : SmPick{0, 1}; k.cols) {
        if (int r = sm_attrs()) return r;
        if (int r = launch_sm(instnorm_bwd_sm, C / (k.cols * 4), N, sm_threads(P / k.cl, k.cols), (size_t)(P / k.cl) * k.cols * 32, k.cl,
                              (cudaStream_t)stream, x, stats, gy, addend, gx, P, C, eps, act, k.cols, k.cl)) return r;
        EG_CHECK_LAUNCH();
        return 0;
    }
    if (C % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(gy) | reinterpret_cast<uintptr_t>(gx) |
                        reinterpret_cast<uintptr_t>(stats) | reinterpret_cast<uintptr_t>(addend)) & 15) == 0) {
        dim3 grid(eg_ceil_div(C, VQ * 4), N), block(VQ, VR);
        instnorm_bwd_v4<<<grid, block, 0, (cudaStream_t)stream>>>(x, stats, gy, addend, gx, P, C, eps, act);
        EG_CHECK_LAUNCH();
        return 0;
    }
    dim3 grid(eg_ceil_div(C, CG), N), block(CG, RY);
    instnorm_bwd_k<<<grid, block, 0, (cudaStream_t)stream>>>(x, stats, gy, addend, gx, P, C, eps, act);
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_instnorm_bwd2(const float* x, const float* stats, const float* gy, const float* t, float* out_gy,
                     float* out_x, int N, int P, int C, float eps, int act, void* stream) {
    EG_REQUIRE(x && stats && gy && t && out_gy && out_x && N > 0 && P > 0 && C > 0 && N <= 65535);
    EG_REQUIRE(act != EG_ACT_TANH);   // second derivative of the activation is taken as zero
    if (const SmPick k = al16all({x, stats, gy, t, out_gy, out_x}) ? sm_pick(P, C, 3) : SmPick{0, 1}; k.cols) {
        if (int r = sm_attrs()) return r;
        if (int r = launch_sm(instnorm_bwd2_sm, C / (k.cols * 4), N, sm_threads(P / k.cl, k.cols), (size_t)(P / k.cl) * k.cols * 48, k.cl,
                              (cudaStream_t)stream, x, stats, gy, t, out_gy, out_x, P, C, eps, act, k.cols, k.cl)) return r;
        EG_CHECK_LAUNCH();
        return 0;
    }
    dim3 grid(eg_ceil_div(C, CG), N), block(CG, RY);
    instnorm_bwd2_k<<<grid, block, 0, (cudaStream_t)stream>>>(x, stats, gy, t, out_gy, out_x, P, C, eps, act);
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_bn_stats(const float* x, float* sums, int R, int C, void* stream) {
    EG_REQUIRE(x && sums && R > 0 && C > 0);
    dim3 grid(eg_ceil_div(C, CG)), block(CG, RY);
    bn_stats_k<<<grid, block, 0, (cudaStream_t)stream>>>(x, sums, R, C);
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_bn_apply(const float* x, const float* sums, float count, const float* gamma, const float* beta, float* y,
                int R, int C, float eps, int act, void* stream) {
    EG_REQUIRE(x && sums && gamma && beta && y && R > 0 && C > 0 && count > 0.f);
    const long long total = (long long)R * C;
    bn_apply_k<<<eg_ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x, sums, count, gamma, beta, y, total, C, eps, act);
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_bn_bwd_reduce(const float* x, const float* sums, float count, const float* gamma, const float* beta,
                     const float* gy, float* red, int R, int C, float eps, int act, void* stream) {
    EG_REQUIRE(x && sums && gamma && beta && gy && red && R > 0 && C > 0 && count > 0.f);
    dim3 grid(eg_ceil_div(C, CG)), block(CG, RY);
    bn_bwd_reduce_k<<<grid, block, 0, (cudaStream_t)stream>>>(x, sums, count, gamma, beta, gy, red, R, C, eps, act);
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_bn_bwd_apply(const float* x, const float* sums, float count, const float* gamma, const float* beta,
                    const float* gy, const float* red, float* gx, int R, int C, float eps, int act, void* stream) {
    EG_REQUIRE(x && sums && gamma && beta && gy && red && gx && R > 0 && C > 0 && count > 0.f);
    const long long total = (long long)R * C;
    bn_bwd_apply_k<<<eg_ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(x, sums, count, gamma, beta, gy, red, gx, total, C, eps, act);
    EG_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
