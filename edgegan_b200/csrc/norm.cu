// Instance norm (+ fused activation) forward / backward / double-backward and the split batch norm of
// the generator's h0.  All kernels are HBM/L2-bound reductions over the H*W positions of one
// (sample, 32-channel group) slab; the slab is streamed twice (stats, apply) and the second pass
// hits the 126 MB L2.
//
// Reference: nn/modules/normalization.py:10-29.
#include "common.cuh"

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

extern "C" {

int eg_instnorm_fwd(const float* x, float* y, float* stats, int N, int P, int C, float eps, int act, void* stream) {
    EG_REQUIRE(x && y && stats && N > 0 && P > 0 && C > 0 && N <= 65535);
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
