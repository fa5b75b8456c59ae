// Kernels specific to the multi-class classifier D2 (reference models/classifier.py:12-119, the MRU block
// nn/modules/conv.py:133-243, spectral norm nn/modules/normalization.py:38-76, focal / CE losses
// nn/functional.py:5-16).  All HBM-bound glue; the classifier's convolutions use the shared conv trio.
#include "common.cuh"
#include <algorithm>
#include <mutex>
#include <vector>

namespace {

constexpr int TB = 256;
inline unsigned grid1d(long long n, int per_block = TB) { return (unsigned)((n + per_block - 1) / per_block); }

// float4 helpers: every elementwise kernel below streams 16 bytes per thread per access (grid-stride), with a scalar
// tail, so that the HBM-bound glue of the classifier runs near the copy bandwidth
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__host__ __device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline unsigned grid4(long long n) {              // blocks of TB threads, >= 4 float4 per thread, capped at 16 waves of 148 SMs
    long long g = (n / 4 + TB * 4 - 1) / (TB * 4);
    if (g < 1) g = 1;
    if (g > 148 * 16) g = 148 * 16;
    return (unsigned)g;
}

// ---- prelu: tf.maximum(leak * x, x), learned scalar leak (activation.py:23-27; tie -> first argument) ----
__device__ __forceinline__ float prelu1(float v, float l) { const float a = l * v; return a >= v ? a : v; }
// y = prelu(x; leak); optionally y2 = prelu(y; leak2) (h0 followed by the first unit's norm_activation_in)
__global__ void __launch_bounds__(TB)
prelu_fwd_k(const float* __restrict__ x, const float* __restrict__ leak, float* __restrict__ y,
            const float* __restrict__ leak2, float* __restrict__ y2, long long n, int vec) {
    const float l = leak[0], l2 = leak2 ? leak2[0] : 0.f;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? n / 4 : 0;
    for (long long i = tid; i < n4; i += nt) {
        const float4 v = ld4(x + 4 * i);
        const float4 o = make_float4(prelu1(v.x, l), prelu1(v.y, l), prelu1(v.z, l), prelu1(v.w, l));
        st4(y + 4 * i, o);
        if (y2) st4(y2 + 4 * i, make_float4(prelu1(o.x, l2), prelu1(o.y, l2), prelu1(o.z, l2), prelu1(o.w, l2)));
    }
    for (long long i = 4 * n4 + tid; i < n; i += nt) {
        const float o = prelu1(x[i], l);
        y[i] = o;
        if (y2) y2[i] = prelu1(o, l2);
    }
}
// gx (= or +=) gy * prelu'(x);  gleak += sum over the "leak side" of gy * x
__global__ void __launch_bounds__(TB)
prelu_bwd_k(const float* __restrict__ x, const float* __restrict__ leak, const float* __restrict__ gy,
            float* __restrict__ gx, float* __restrict__ gleak, long long n, int vec, int acc_gx) {
    __shared__ float red[33];
    const float l = leak[0];
    float acc = 0.f;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? n / 4 : 0;
    for (long long i = tid; i < n4; i += nt) {
        const float4 v = ld4(x + 4 * i), g = ld4(gy + 4 * i);
        const bool fx = l * v.x >= v.x, fy = l * v.y >= v.y, fz = l * v.z >= v.z, fw = l * v.w >= v.w;
        if (gx != nullptr) {
            float4 o = make_float4(fx ? g.x * l : g.x, fy ? g.y * l : g.y, fz ? g.z * l : g.z, fw ? g.w * l : g.w);
            if (acc_gx) { const float4 c = ld4(gx + 4 * i); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
            st4(gx + 4 * i, o);
        }
        if (fx) acc = fmaf(g.x, v.x, acc);
        if (fy) acc = fmaf(g.y, v.y, acc);
        if (fz) acc = fmaf(g.z, v.z, acc);
        if (fw) acc = fmaf(g.w, v.w, acc);
    }
    for (long long i = 4 * n4 + tid; i < n; i += nt) {
        const float v = x[i], g = gy[i];
        const bool first = (l * v >= v);
        if (gx != nullptr) { const float o = first ? g * l : g; gx[i] = acc_gx ? gx[i] + o : o; }
        if (first) acc = fmaf(g, v, acc);
    }
    if (gleak != nullptr) {
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) atomicAdd(gleak, acc);
    }
}

// ---- gate min-max normalisation (conv.py:197-198; SURVEY A11: ties share the gradient) -------------------
constexpr int CG = 32, RY = 8;

__global__ void __launch_bounds__(CG * RY)
minmax_fwd_k(const float* __restrict__ x, float* __restrict__ y, float* __restrict__ stats, int P, int C) {
    __shared__ float smn[RY][CG], smx[RY][CG];
    const int n = blockIdx.y, c = blockIdx.x * CG + threadIdx.x;
    const bool ok = c < C;
    const size_t base = (size_t)n * P * C + c;
    float mn = INFINITY, mx = -INFINITY;
    if (ok) for (int p = threadIdx.y; p < P; p += RY) { const float v = x[base + (size_t)p * C]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
    smn[threadIdx.y][threadIdx.x] = mn; smx[threadIdx.y][threadIdx.x] = mx;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RY; ++r) { mn = fminf(mn, smn[r][threadIdx.x]); mx = fmaxf(mx, smx[r][threadIdx.x]); }
    if (!ok) return;
    if (threadIdx.y == 0) { stats[((size_t)n * C + c) * 2] = mn; stats[((size_t)n * C + c) * 2 + 1] = mx; }
    const float d = mx - mn;
    for (int p = threadIdx.y; p < P; p += RY) y[base + (size_t)p * C] = (x[base + (size_t)p * C] - mn) / d;
}

__global__ void __launch_bounds__(CG * RY)
minmax_bwd_k(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gy,
             float* __restrict__ gx, int P, int C) {
    __shared__ float sm[4][RY][CG];
    const int n = blockIdx.y, c = blockIdx.x * CG + threadIdx.x;
    const bool ok = c < C;
    const size_t base = (size_t)n * P * C + c;
    float mn = 0.f, mx = 1.f;
    if (ok) { mn = stats[((size_t)n * C + c) * 2]; mx = stats[((size_t)n * C + c) * 2 + 1]; }
    const float d = mx - mn;
    float s1 = 0.f, s2 = 0.f, cmin = 0.f, cmax = 0.f;     // sum gy, sum gy*y, #argmin, #argmax
    if (ok) for (int p = threadIdx.y; p < P; p += RY) {
        const float v = x[base + (size_t)p * C], g = gy[base + (size_t)p * C];
        s1 += g; s2 = fmaf(g, (v - mn) / d, s2);
        cmin += (v == mn) ? 1.f : 0.f; cmax += (v == mx) ? 1.f : 0.f;
    }
    sm[0][threadIdx.y][threadIdx.x] = s1; sm[1][threadIdx.y][threadIdx.x] = s2;
    sm[2][threadIdx.y][threadIdx.x] = cmin; sm[3][threadIdx.y][threadIdx.x] = cmax;
    __syncthreads();
    s1 = s2 = cmin = cmax = 0.f;
#pragma unroll
    for (int r = 0; r < RY; ++r) { s1 += sm[0][r][threadIdx.x]; s2 += sm[1][r][threadIdx.x]; cmin += sm[2][r][threadIdx.x]; cmax += sm[3][r][threadIdx.x]; }
    if (!ok) return;
    const float gmin = (s2 - s1) / d / cmin, gmax = -s2 / d / cmax;
    for (int p = threadIdx.y; p < P; p += RY) {
        const size_t i = base + (size_t)p * C;
        const float v = x[i];
        float o = gy[i] / d;
        if (v == mn) o += gmin;
        if (v == mx) o += gmax;
        gx[i] = o;
    }
}

__global__ void __launch_bounds__(TB)
fma3_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
       float* __restrict__ out, long long n, int vec) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? n / 4 : 0;
    for (long long i = tid; i < n4; i += nt) {
        const float4 x = ld4(a + 4 * i), y = ld4(b + 4 * i), z = ld4(c + 4 * i);
        st4(out + 4 * i, make_float4(fmaf(y.x, z.x, x.x), fmaf(y.y, z.y, x.y), fmaf(y.z, z.z, x.z), fmaf(y.w, z.w, x.w)));
    }
    for (long long i = 4 * n4 + tid; i < n; i += nt) out[i] = fmaf(b[i], c[i], a[i]);
}
__global__ void __launch_bounds__(TB)
mul_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n, int vec) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    const long long n4 = vec ? n / 4 : 0;
    for (long long i = tid; i < n4; i += nt) {
        const float4 x = ld4(a + 4 * i), y = ld4(b + 4 * i);
        st4(out + 4 * i, make_float4(x.x * y.x, x.y * y.y, x.z * y.z, x.w * y.w));
    }
    for (long long i = 4 * n4 + tid; i < n; i += nt) out[i] = a[i] * b[i];
}

// ---- fused MRU gate (conv.py:189-206) ------------------------------------------------------------------------
//   rgl  = lrelu(cg + cg_i)          cg = SNconv(a_in) + bias, cg_i = SNconv(inp): the two parts of conv(concat(a_in, inp))
//   rg   = (rgl - min_HW) / (max_HW - min_HW)
//   plus = ht + rg * img ;  hin = prelu(plus)
// One CTA per (sample, group of GC channels), 256 threads = (GC/4 float4 columns) x rows.  Pass 1 streams cg / cg_i,
// writes rgl over cg and finds min / max; pass 2 re-reads the CTA's own rgl slab (<= 128 KB: L2) together with ht and
// img and writes plus and hin.  rg itself is never stored (the backward recomputes it from rgl and the statistics);
// stats[n][c] = {min, max, #argmin, #argmax} (tf.reduce_min / max split their gradient between ties).
// Replaces axpby + act_fwd + minmax_fwd + fma3 + prelu_fwd: 17 -> 10 tensor passes.
__device__ __forceinline__ float4 f4min(float4 a, float4 b) { return make_float4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)); }
__device__ __forceinline__ float4 f4max(float4 a, float4 b) { return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }
__device__ __forceinline__ float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float lrelu_gate(float v) { return v > 0.f ? v : 0.2f * v; }          // tf.maximum(0.2x, x)

// tree reduction over the row index of a [rows][tpp] thread layout; result broadcast to every thread of the column
template <int OP>   // 0 min, 1 max, 2 sum
__device__ __forceinline__ float4 col_reduce(float4 v, float4* sm, int tpp, int rows, int tx, int ty) {
    const int tid = ty * tpp + tx;
    __syncthreads();
    sm[tid] = v;
    __syncthreads();
    for (int s = rows >> 1; s > 0; s >>= 1) {
        if (ty < s) {
            const float4 o = sm[tid + s * tpp];
            sm[tid] = OP == 0 ? f4min(sm[tid], o) : (OP == 1 ? f4max(sm[tid], o) : f4add(sm[tid], o));
        }
        __syncthreads();
    }
    return sm[tx];
}

__global__ void __launch_bounds__(256)
mru_gate_fwd_k(float* __restrict__ cg, const float* __restrict__ cgi, const float* __restrict__ ht,
               const float* __restrict__ img, const float* __restrict__ leak, float* __restrict__ stats,
               float* __restrict__ plus, float* __restrict__ hin, int P, int C, int GC) {
    __shared__ float4 sm[256];
    const int tpp = GC >> 2, rows = 256 / tpp;
    const int tx = threadIdx.x % tpp, ty = threadIdx.x / tpp;
    const int n = blockIdx.y, c = blockIdx.x * GC + tx * 4;
    const size_t base = (size_t)n * P * C + c;
    float4 mn = make_float4(INFINITY, INFINITY, INFINITY, INFINITY), mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int p = ty; p < P; p += rows) {
        const size_t i = base + (size_t)p * C;
        const float4 a = ld4(cg + i), b = ld4(cgi + i);
        const float4 r = make_float4(lrelu_gate(a.x + b.x), lrelu_gate(a.y + b.y), lrelu_gate(a.z + b.z), lrelu_gate(a.w + b.w));
        st4(cg + i, r);
        mn = f4min(mn, r); mx = f4max(mx, r);
    }
    mn = col_reduce<0>(mn, sm, tpp, rows, tx, ty);
    mx = col_reduce<1>(mx, sm, tpp, rows, tx, ty);
    const float l = leak[0];
    const float4 d = make_float4(mx.x - mn.x, mx.y - mn.y, mx.z - mn.z, mx.w - mn.w);
    float4 cmin = make_float4(0.f, 0.f, 0.f, 0.f), cmax = cmin;
    for (int p = ty; p < P; p += rows) {
        const size_t i = base + (size_t)p * C;
        const float4 r = ld4(cg + i), h = ld4(ht + i), im = ld4(img + i);
        const float4 pl = make_float4(fmaf((r.x - mn.x) / d.x, im.x, h.x), fmaf((r.y - mn.y) / d.y, im.y, h.y),
                                      fmaf((r.z - mn.z) / d.z, im.z, h.z), fmaf((r.w - mn.w) / d.w, im.w, h.w));
        st4(plus + i, pl);
        st4(hin + i, make_float4(prelu1(pl.x, l), prelu1(pl.y, l), prelu1(pl.z, l), prelu1(pl.w, l)));
        cmin.x += r.x == mn.x; cmin.y += r.y == mn.y; cmin.z += r.z == mn.z; cmin.w += r.w == mn.w;
        cmax.x += r.x == mx.x; cmax.y += r.y == mx.y; cmax.z += r.z == mx.z; cmax.w += r.w == mx.w;
    }
    cmin = col_reduce<2>(cmin, sm, tpp, rows, tx, ty);
    cmax = col_reduce<2>(cmax, sm, tpp, rows, tx, ty);
    if (ty == 0) {
        float* sp = stats + ((size_t)n * C + c) * 4;
        st4(sp, make_float4(mn.x, mx.x, cmin.x, cmax.x)); st4(sp + 4, make_float4(mn.y, mx.y, cmin.y, cmax.y));
        st4(sp + 8, make_float4(mn.z, mx.z, cmin.z, cmax.z)); st4(sp + 12, make_float4(mn.w, mx.w, cmin.w, cmax.w));
    }
}

// Backward of the same block, given g_hin (cotangent of hin):
//   g_plus = g_hin * prelu'(plus) (+ leak gradient);  g_ht += g_plus;  g_img = g_plus * rg;  g_rg = g_plus * img
//   g_rgl  = min-max backward of g_rg (two per-channel sums, ties share);  g_cg = g_rgl * lrelu'(rgl)
// g_cg doubles as the scratch for g_rg between the two passes (same thread, same index).
// Replaces prelu_bwd + axpby + 2 x mul + minmax_bwd + act_bwd: 22 -> 10 tensor passes.
__global__ void __launch_bounds__(256)
mru_gate_bwd_k(const float* __restrict__ plus, const float* __restrict__ g_hin, const float* __restrict__ img,
               const float* __restrict__ rgl, const float* __restrict__ stats, const float* __restrict__ leak,
               float* __restrict__ g_ht, float* __restrict__ g_img, float* __restrict__ g_cg, float* __restrict__ gleak,
               int P, int C, int GC) {
    __shared__ float4 sm[256];
    __shared__ float red[33];
    const int tpp = GC >> 2, rows = 256 / tpp;
    const int tx = threadIdx.x % tpp, ty = threadIdx.x / tpp;
    const int n = blockIdx.y, c = blockIdx.x * GC + tx * 4;
    const size_t base = (size_t)n * P * C + c;
    const float* sp = stats + ((size_t)n * C + c) * 4;
    const float4 s0 = ld4(sp), s1_ = ld4(sp + 4), s2_ = ld4(sp + 8), s3 = ld4(sp + 12);
    const float4 mn = make_float4(s0.x, s1_.x, s2_.x, s3.x), mx = make_float4(s0.y, s1_.y, s2_.y, s3.y);
    const float4 cmin = make_float4(s0.z, s1_.z, s2_.z, s3.z), cmax = make_float4(s0.w, s1_.w, s2_.w, s3.w);
    const float4 d = make_float4(mx.x - mn.x, mx.y - mn.y, mx.z - mn.z, mx.w - mn.w);
    const float l = leak[0];
    float4 a1 = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a1;      // sum g_rg, sum g_rg * rg
    float accl = 0.f;
    for (int p = ty; p < P; p += rows) {
        const size_t i = base + (size_t)p * C;
        const float4 pl = ld4(plus + i), g = ld4(g_hin + i), im = ld4(img + i), r = ld4(rgl + i);
        float4 gh = ld4(g_ht + i);
        float4 gp;
        { const bool f = l * pl.x >= pl.x; gp.x = f ? g.x * l : g.x; if (f) accl = fmaf(g.x, pl.x, accl); }
        { const bool f = l * pl.y >= pl.y; gp.y = f ? g.y * l : g.y; if (f) accl = fmaf(g.y, pl.y, accl); }
        { const bool f = l * pl.z >= pl.z; gp.z = f ? g.z * l : g.z; if (f) accl = fmaf(g.z, pl.z, accl); }
        { const bool f = l * pl.w >= pl.w; gp.w = f ? g.w * l : g.w; if (f) accl = fmaf(g.w, pl.w, accl); }
        gh.x += gp.x; gh.y += gp.y; gh.z += gp.z; gh.w += gp.w;
        st4(g_ht + i, gh);
        const float4 rg = make_float4((r.x - mn.x) / d.x, (r.y - mn.y) / d.y, (r.z - mn.z) / d.z, (r.w - mn.w) / d.w);
        st4(g_img + i, make_float4(gp.x * rg.x, gp.y * rg.y, gp.z * rg.z, gp.w * rg.w));
        const float4 grg = make_float4(gp.x * im.x, gp.y * im.y, gp.z * im.z, gp.w * im.w);
        st4(g_cg + i, grg);
        a1 = f4add(a1, grg);
        a2.x = fmaf(grg.x, rg.x, a2.x); a2.y = fmaf(grg.y, rg.y, a2.y); a2.z = fmaf(grg.z, rg.z, a2.z); a2.w = fmaf(grg.w, rg.w, a2.w);
    }
    a1 = col_reduce<2>(a1, sm, tpp, rows, tx, ty);
    a2 = col_reduce<2>(a2, sm, tpp, rows, tx, ty);
    if (gleak != nullptr) {
        accl = block_sum(accl, red);
        if (threadIdx.x == 0) atomicAdd(gleak, accl);
    }
    const float4 gmin = make_float4((a2.x - a1.x) / d.x / cmin.x, (a2.y - a1.y) / d.y / cmin.y, (a2.z - a1.z) / d.z / cmin.z, (a2.w - a1.w) / d.w / cmin.w);
    const float4 gmax = make_float4(-a2.x / d.x / cmax.x, -a2.y / d.y / cmax.y, -a2.z / d.z / cmax.z, -a2.w / d.w / cmax.w);
    for (int p = ty; p < P; p += rows) {
        const size_t i = base + (size_t)p * C;
        const float4 r = ld4(rgl + i), grg = ld4(g_cg + i);
        float4 o = make_float4(grg.x / d.x, grg.y / d.y, grg.z / d.z, grg.w / d.w);
        if (r.x == mn.x) o.x += gmin.x; if (r.x == mx.x) o.x += gmax.x;
        if (r.y == mn.y) o.y += gmin.y; if (r.y == mx.y) o.y += gmax.y;
        if (r.z == mn.z) o.z += gmin.z; if (r.z == mx.z) o.z += gmax.z;
        if (r.w == mn.w) o.w += gmin.w; if (r.w == mx.w) o.w += gmax.w;
        o.x *= r.x > 0.f ? 1.f : 0.2f; o.y *= r.y > 0.f ? 1.f : 0.2f; o.z *= r.z > 0.f ? 1.f : 0.2f; o.w *= r.w > 0.f ? 1.f : 0.2f;
        st4(g_cg + i, o);
    }
}

// ---- spectral norm, one power iteration from a frozen u (normalization.py:38-76) -----------
// ws layout: [0,K) p = W u^T ; [K,K+C) r = v W ; [K+C, K+C+8) scalars {np, nr, sigma, s, -, -, -, -} ; [K+C+8, 2K+C+8) gv
constexpr float SN_EPS = 1e-12f;

__global__ void __launch_bounds__(256)
sn_p_k(const float* __restrict__ W, const float* __restrict__ u, float* __restrict__ p, int K, int C) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= K) return;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(W[(size_t)row * C + c], u[c], s);
    s = warp_sum(s);
    if (lane == 0) p[row] = s;
}
// r[c] += sum over this block's k-slab of v[k] W[k,c]
__global__ void __launch_bounds__(256)
sn_r_k(const float* __restrict__ W, const float* __restrict__ p, float* __restrict__ r, int K, int C, int kslab) {
    __shared__ float red[33];
    float s = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) s = fmaf(p[k], p[k], s);
    const float np = sqrtf(block_sum(s, red));
    const float inv = 1.f / (np + SN_EPS);
    const int k0 = blockIdx.y * kslab, k1 = min(K, k0 + kslab);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(p[k] * inv, W[(size_t)k * C + c], acc);
    atomicAdd(r + c, acc);
}
__global__ void __launch_bounds__(256)
sn_scale_k(const float* __restrict__ W, const float* __restrict__ p, const float* __restrict__ r, float* __restrict__ scal,
           float* __restrict__ Wbar, long long n, int K, int C) {
    __shared__ float red[33];
    float s = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) s = fmaf(r[c], r[c], s);
    const float nr = sqrtf(block_sum(s, red));
    const float sigma = nr * nr / (nr + SN_EPS);
    if (blockIdx.x == 0) {
        float t = 0.f;
        for (int k = threadIdx.x; k < K; k += blockDim.x) t = fmaf(p[k], p[k], t);
        t = block_sum(t, red);
        if (threadIdx.x == 0) { scal[0] = sqrtf(t); scal[1] = nr; scal[2] = sigma; }
    }
    const float inv = 1.f / sigma;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        Wbar[i] = W[i] * inv;
}
__global__ void __launch_bounds__(256)
sn_dot_k(const float* __restrict__ G, const float* __restrict__ W, float* __restrict__ out, long long n) {
    __shared__ float red[33];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        s = fmaf(G[i], W[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(out, s);
}
// gv[k] = sum_c W[k,c] * gr[c],  gr = dsigma/dr = coef * r / nr
__global__ void __launch_bounds__(256)
sn_gv_k(const float* __restrict__ W, const float* __restrict__ r, const float* __restrict__ scal, float* __restrict__ gv, int K, int C) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= K) return;
    const float nr = scal[1];
    const float coef = (nr * nr + 2.f * SN_EPS * nr) / ((nr + SN_EPS) * (nr + SN_EPS)) / nr;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(W[(size_t)row * C + c], coef * r[c], s);
    s = warp_sum(s);
    if (lane == 0) gv[row] = s;
}
__global__ void __launch_bounds__(256)
sn_gw_k(const float* __restrict__ W, const float* __restrict__ u, const float* __restrict__ p, const float* __restrict__ r,
        const float* __restrict__ scal, const float* __restrict__ gv, const float* __restrict__ G, float* __restrict__ gW,
        int K, int C) {
    __shared__ float red[33];
    float t = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) t = fmaf(gv[k], p[k], t);
    const float gvp = block_sum(t, red);
    const float np = scal[0], nr = scal[1], sigma = scal[2], sdot = scal[3];
    const float coef = (nr * nr + 2.f * SN_EPS * nr) / ((nr + SN_EPS) * (nr + SN_EPS)) / nr;
    const float a = 1.f / (np + SN_EPS), b = gvp / (np * (np + SN_EPS) * (np + SN_EPS));
    const float f = sdot / (sigma * sigma), isg = 1.f / sigma;
    const long long n = (long long)K * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i / C), c = (int)(i % C);
        const float v = p[k] * a;
        const float gp = gv[k] * a - b * p[k];
        const float dsig = v * (coef * r[c]) + gp * u[c];
        gW[i] = G[i] * isg - f * dsig;
    }
}

// ---- the same for a whole network in three launches per direction (eg_spectral_norm_set_*) ------------------------------
// One table entry per weight tensor; blockIdx.y (z for the r pass) selects the tensor and blocks beyond a tensor's own
// extent leave.  A tensor may be SPLIT along its input-channel axis (cin = hd + rest): Wbar is then also written as two
// contiguous filters (Wa [taps, hd, C], Wi [taps, cin - hd, C]) and dL/dWbar is read from two such parts (Ga, Gi) --
// the classifier's update gate convolves concat(a, image) as two convs (classifier.py:69-76 of the reference builds the concat).
struct SnDesc {
    const float* W; const float* u; float* Wbar; float* ws;
    const float* G; float* gW;
    float* Wa; float* Wi;
    const float* Ga; const float* Gi;
    int K, C, cin, hd;
};
struct SnSet { SnDesc* tab; float* dots; int n, Kmax, Cmax; long long nmax; bool bwd_ok; };
static std::mutex g_sn_mu;
static std::vector<SnSet*> g_sn_sets;

__device__ __forceinline__ unsigned sn_blocks(long long n) {
    long long g = (n + 1023) / 1024;
    return (unsigned)(g > 1184 ? 1184 : (g < 1 ? 1 : g));
}
// offset of element (k, c) inside the split part it belongs to; part = 0 (a) or 1 (i)
__device__ __forceinline__ long long sn_split(const SnDesc& d, int k, int c, int& part) {
    const int tap = k / d.cin, ci = k - tap * d.cin;
    if (ci < d.hd) { part = 0; return ((long long)tap * d.hd + ci) * d.C + c; }
    part = 1;
    return ((long long)tap * (d.cin - d.hd) + (ci - d.hd)) * d.C + c;
}
__device__ __forceinline__ float sn_G(const SnDesc& d, long long i) {
    if (d.Ga == nullptr) return d.G[i];
    int part;
    const int k = (int)(i / d.C), c = (int)(i - (long long)k * d.C);
    const long long o = sn_split(d, k, c, part);
    return part == 0 ? d.Ga[o] : d.Gi[o];
}
__global__ void __launch_bounds__(256)
sn_set_p_k(const SnDesc* __restrict__ tab) {
    const SnDesc d = tab[blockIdx.y];
    if (blockIdx.x == 0)                                     // r and the scalars are accumulated by the passes that follow
        for (int i = threadIdx.x; i < d.C + 8; i += blockDim.x) d.ws[d.K + i] = 0.f;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= d.K) return;
    float s = 0.f;
    for (int c = lane; c < d.C; c += 32) s = fmaf(d.W[(size_t)row * d.C + c], d.u[c], s);
    s = warp_sum(s);
    if (lane == 0) d.ws[row] = s;
}
__global__ void __launch_bounds__(256)
sn_set_r_k(const SnDesc* __restrict__ tab, int kslab) {
    __shared__ float red[33];
    const SnDesc d = tab[blockIdx.z];
    const int k0 = blockIdx.y * kslab;
    if (k0 >= d.K || blockIdx.x * blockDim.x >= d.C) return;
    const float* p = d.ws;
    float s = 0.f;
    for (int k = threadIdx.x; k < d.K; k += blockDim.x) s = fmaf(p[k], p[k], s);
    const float np = sqrtf(block_sum(s, red));
    const float inv = 1.f / (np + SN_EPS);
    const int k1 = min(d.K, k0 + kslab);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d.C) return;
    float acc = 0.f;
    for (int k = k0; k < k1; ++k) acc = fmaf(p[k] * inv, d.W[(size_t)k * d.C + c], acc);
    atomicAdd(d.ws + d.K + c, acc);
}
__global__ void __launch_bounds__(256)
sn_set_scale_k(const SnDesc* __restrict__ tab) {
    __shared__ float red[33];
    const SnDesc d = tab[blockIdx.y];
    const long long n = (long long)d.K * d.C;
    const unsigned g = sn_blocks(n);
    if (blockIdx.x >= g) return;
    const float *p = d.ws, *r = d.ws + d.K;
    float* scal = d.ws + d.K + d.C;
    float s = 0.f;
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) s = fmaf(r[c], r[c], s);
    const float nr = sqrtf(block_sum(s, red));
    const float sigma = nr * nr / (nr + SN_EPS);
    if (blockIdx.x == 0) {
        float t = 0.f;
        for (int k = threadIdx.x; k < d.K; k += blockDim.x) t = fmaf(p[k], p[k], t);
        t = block_sum(t, red);
        if (threadIdx.x == 0) { scal[0] = sqrtf(t); scal[1] = nr; scal[2] = sigma; }
    }
    const float inv = 1.f / sigma;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)g * blockDim.x) {
        const float v = d.W[i] * inv;
        d.Wbar[i] = v;
        if (d.Wa != nullptr) {
            int part;
            const int k = (int)(i / d.C), c = (int)(i - (long long)k * d.C);
            const long long o = sn_split(d, k, c, part);
            (part == 0 ? d.Wa : d.Wi)[o] = v;
        }
    }
}
__global__ void __launch_bounds__(256)
sn_set_dot_k(const SnDesc* __restrict__ tab, float* __restrict__ dots) {
    __shared__ float red[33];
    const SnDesc d = tab[blockIdx.y];
    const long long n = (long long)d.K * d.C;
    const unsigned g = sn_blocks(n);
    if (blockIdx.x >= g) return;
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)g * blockDim.x)
        s = fmaf(sn_G(d, i), d.W[i], s);
    s = block_sum(s, red);
    if (threadIdx.x == 0) atomicAdd(dots + blockIdx.y, s);
}
__global__ void __launch_bounds__(256)
sn_set_gv_k(const SnDesc* __restrict__ tab) {
    const SnDesc d = tab[blockIdx.y];
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= d.K) return;
    const float *r = d.ws + d.K, *scal = d.ws + d.K + d.C;
    const float nr = scal[1];
    const float coef = (nr * nr + 2.f * SN_EPS * nr) / ((nr + SN_EPS) * (nr + SN_EPS)) / nr;
    float s = 0.f;
    for (int c = lane; c < d.C; c += 32) s = fmaf(d.W[(size_t)row * d.C + c], coef * r[c], s);
    s = warp_sum(s);
    if (lane == 0) d.ws[d.K + d.C + 8 + row] = s;
}
__global__ void __launch_bounds__(256)
sn_set_gw_k(const SnDesc* __restrict__ tab, const float* __restrict__ dots) {
    __shared__ float red[33];
    const SnDesc d = tab[blockIdx.y];
    const long long n = (long long)d.K * d.C;
    const unsigned g = sn_blocks(n);
    if (blockIdx.x >= g) return;
    const float *p = d.ws, *r = d.ws + d.K, *scal = d.ws + d.K + d.C, *gv = d.ws + d.K + d.C + 8;
    float t = 0.f;
    for (int k = threadIdx.x; k < d.K; k += blockDim.x) t = fmaf(gv[k], p[k], t);
    const float gvp = block_sum(t, red);
    const float np = scal[0], nr = scal[1], sigma = scal[2], sdot = dots[blockIdx.y];
    const float coef = (nr * nr + 2.f * SN_EPS * nr) / ((nr + SN_EPS) * (nr + SN_EPS)) / nr;
    const float a = 1.f / (np + SN_EPS), b = gvp / (np * (np + SN_EPS) * (np + SN_EPS));
    const float f = sdot / (sigma * sigma), isg = 1.f / sigma;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)g * blockDim.x) {
        const int k = (int)(i / d.C), c = (int)(i - (long long)k * d.C);
        const float v = p[k] * a;
        const float gp = gv[k] * a - b * p[k];
        const float dsig = v * (coef * r[c]) + gp * d.u[c];
        d.gW[i] = sn_G(d, i) * isg - f * dsig;
    }
}

// ---- softmax cross-entropy losses (functional.py:5-16) -------------------------------------------------
// one thread per sample; labels are the float class ids stored in column `label_col` of z
__global__ void softmax_ce_bwd_k(const float* __restrict__ logits, const float* __restrict__ z, int zstride, int label_col,
                                 int B, int C, int focal, float weight, float inv_b, float* __restrict__ glogits,
                                 float* __restrict__ loss) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float* l = logits + (size_t)b * C;
    const int y = (int)z[(size_t)b * zstride + label_col];
    float mx = -INFINITY;
    for (int j = 0; j < C; ++j) mx = fmaxf(mx, l[j]);
    float se = 0.f;
    for (int j = 0; j < C; ++j) se += expf(l[j] - mx);
    const float lse = logf(se) + mx;
    const float ce = lse - l[y];
    const float py = expf(l[y] - lse);
    float fac = 1.f, lb = ce;
    if (focal) { fac = (1.f - py) * (1.f - py) + 2.f * ce * (1.f - py) * py; lb = (1.f - py) * (1.f - py) * ce; }
    for (int j = 0; j < C; ++j) {
        const float pj = expf(l[j] - lse);
        glogits[(size_t)b * C + j] = weight * inv_b * fac * (pj - (j == y ? 1.f : 0.f));
    }
    atomicAdd(loss, weight * inv_b * lb);
}

// ---- mean_pool of a sum (pooling.py:4-8 after conv.py:236) and global mean (classifier.py:111) ---------
// y = mean_2x2(a + b); optionally y_act = prelu(y; leak) (the next unit's norm_activation_in).  V = 4: float4 over channels.
template <int V>
__global__ void __launch_bounds__(TB)
add_pool2_fwd_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                const float* __restrict__ leak, float* __restrict__ y_act, int N, int H, int W, int C) {
    const int OH = H / 2, OW = W / 2, CV = C / V;
    const long long total = (long long)N * OH * OW * CV;
    const float l = leak ? leak[0] : 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % CV) * V; long long t = i / CV;
        const int ox = (int)(t % OW); t /= OW;
        const int oy = (int)(t % OH); const int n = (int)(t / OH);
        float acc[V];
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const size_t j = (((size_t)n * H + 2 * oy + dy) * W + 2 * ox + dx) * C + c;
                if (V == 4) {
                    const float4 u = ld4(a + j);
                    acc[0] += u.x; acc[1 % V] += u.y; acc[2 % V] += u.z; acc[3 % V] += u.w;
                    if (b) { const float4 w = ld4(b + j); acc[0] += w.x; acc[1 % V] += w.y; acc[2 % V] += w.z; acc[3 % V] += w.w; }
                } else {
                    acc[0] += a[j] + (b ? b[j] : 0.f);
                }
            }
        const size_t o = (((size_t)n * OH + oy) * OW + ox) * C + c;
#pragma unroll
        for (int k = 0; k < V; ++k) acc[k] *= 0.25f;
        if (V == 4) {
            st4(y + o, make_float4(acc[0], acc[1 % V], acc[2 % V], acc[3 % V]));
            if (y_act) st4(y_act + o, make_float4(prelu1(acc[0], l), prelu1(acc[1 % V], l), prelu1(acc[2 % V], l), prelu1(acc[3 % V], l)));
        } else {
            y[o] = acc[0];
            if (y_act) y_act[o] = prelu1(acc[0], l);
        }
    }
}
// gx (= or +=) gy / 4 broadcast over each 2x2 window; one thread per gy vector -> four gx vectors
template <int V>
__global__ void __launch_bounds__(TB)
pool2_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, int N, int H, int W, int C, int accumulate) {
    const int OH = H / 2, OW = W / 2, CV = C / V;
    const long long total = (long long)N * OH * OW * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % CV) * V; long long t = i / CV;
        const int ox = (int)(t % OW); t /= OW;
        const int oy = (int)(t % OH); const int n = (int)(t / OH);
        const size_t o = (((size_t)n * OH + oy) * OW + ox) * C + c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (V == 4) { v = ld4(gy + o); v.x *= 0.25f; v.y *= 0.25f; v.z *= 0.25f; v.w *= 0.25f; } else v.x = 0.25f * gy[o];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const size_t j = (((size_t)n * H + 2 * oy + dy) * W + 2 * ox + dx) * C + c;
                if (V == 4) {
                    float4 w = v;
                    if (accumulate) { const float4 u = ld4(gx + j); w.x += u.x; w.y += u.y; w.z += u.z; w.w += u.w; }
                    st4(gx + j, w);
                } else {
                    gx[j] = accumulate ? gx[j] + v.x : v.x;
                }
            }
    }
}
__global__ void globalmean_fwd_k(const float* __restrict__ x, float* __restrict__ y, int N, int P, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int c = i % C, n = i / C;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += x[((size_t)n * P + p) * C + c];
    y[i] = s / P;
}
__global__ void globalmean_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, long long total, int P, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); const long long n = i / ((long long)P * C);
    gx[i] = gy[n * C + c] / P;
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" {

int eg_prelu_fwd(const float* x, const float* leak, float* y, long long n, void* stream) {
    EG_REQUIRE(x && leak && y && n > 0);
    prelu_fwd_k<<<grid4(n), TB, 0, ST>>>(x, leak, y, nullptr, nullptr, n, al16(x) && al16(y));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_prelu_fwd2(const float* x, const float* leak, float* y, const float* leak2, float* y2, long long n, void* stream) {
    EG_REQUIRE(x && leak && y && leak2 && y2 && n > 0);
    prelu_fwd_k<<<grid4(n), TB, 0, ST>>>(x, leak, y, leak2, y2, n, al16(x) && al16(y) && al16(y2));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_prelu_bwd_ex(const float* x, const float* leak, const float* gy, float* gx, float* gleak, long long n,
                    int accumulate_leak, int accumulate_gx, void* stream) {
    EG_REQUIRE(x && leak && gy && n > 0 && (gx || gleak));
    if (gleak && !accumulate_leak) {
        cudaError_t e = cudaMemsetAsync(gleak, 0, sizeof(float), ST);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    prelu_bwd_k<<<grid4(n), TB, 0, ST>>>(x, leak, gy, gx, gleak, n, al16(x) && al16(gy) && (!gx || al16(gx)), accumulate_gx);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_prelu_bwd(const float* x, const float* leak, const float* gy, float* gx, float* gleak, long long n,
                 int accumulate_leak, void* stream) {
    return eg_prelu_bwd_ex(x, leak, gy, gx, gleak, n, accumulate_leak, 0, stream);
}
static int gate_group(int C) { return C % 32 == 0 ? 32 : (C % 8 == 0 ? 8 : 4); }
int eg_mru_gate_fwd(float* cg_rgl, const float* cg_i, const float* ht, const float* img, const float* leak, float* stats,
                    float* plus, float* hin, int N, int P, int C, void* stream) {
    EG_REQUIRE(cg_rgl && cg_i && ht && img && leak && stats && plus && hin && N > 0 && P > 0 && C > 0 && C % 4 == 0 && N <= 65535);
    EG_REQUIRE(al16(cg_rgl) && al16(cg_i) && al16(ht) && al16(img) && al16(stats) && al16(plus) && al16(hin));
    const int GC = gate_group(C);
    mru_gate_fwd_k<<<dim3(C / GC, N), 256, 0, ST>>>(cg_rgl, cg_i, ht, img, leak, stats, plus, hin, P, C, GC);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_mru_gate_bwd(const float* plus, const float* g_hin, const float* img, const float* rgl, const float* stats,
                    const float* leak, float* g_ht, float* g_img, float* g_cg, float* gleak, int accumulate_leak,
                    int N, int P, int C, void* stream) {
    EG_REQUIRE(plus && g_hin && img && rgl && stats && leak && g_ht && g_img && g_cg && N > 0 && P > 0 && C > 0 && C % 4 == 0 && N <= 65535);
    EG_REQUIRE(al16(plus) && al16(g_hin) && al16(img) && al16(rgl) && al16(stats) && al16(g_ht) && al16(g_img) && al16(g_cg));
    if (gleak && !accumulate_leak) {
        cudaError_t e = cudaMemsetAsync(gleak, 0, sizeof(float), ST);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    const int GC = gate_group(C);
    mru_gate_bwd_k<<<dim3(C / GC, N), 256, 0, ST>>>(plus, g_hin, img, rgl, stats, leak, g_ht, g_img, g_cg, gleak, P, C, GC);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_minmax_fwd(const float* x, float* y, float* stats, int N, int P, int C, void* stream) {
    EG_REQUIRE(x && y && stats && N > 0 && P > 0 && C > 0 && N <= 65535);
    dim3 grid(eg_ceil_div(C, CG), N), block(CG, RY);
    minmax_fwd_k<<<grid, block, 0, ST>>>(x, y, stats, P, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_minmax_bwd(const float* x, const float* stats, const float* gy, float* gx, int N, int P, int C, void* stream) {
    EG_REQUIRE(x && stats && gy && gx && N > 0 && P > 0 && C > 0 && N <= 65535);
    dim3 grid(eg_ceil_div(C, CG), N), block(CG, RY);
    minmax_bwd_k<<<grid, block, 0, ST>>>(x, stats, gy, gx, P, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_fma3(const float* a, const float* b, const float* c, float* out, long long n, void* stream) {
    EG_REQUIRE(a && b && c && out && n > 0);
    fma3_k<<<grid4(n), TB, 0, ST>>>(a, b, c, out, n, al16(a) && al16(b) && al16(c) && al16(out));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_mul(const float* a, const float* b, float* out, long long n, void* stream) {
    EG_REQUIRE(a && b && out && n > 0);
    mul_k<<<grid4(n), TB, 0, ST>>>(a, b, out, n, al16(a) && al16(b) && al16(out));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_add_pool2_prelu_fwd(const float* a, const float* b, float* y, const float* leak, float* y_act, int N, int H, int W, int C,
                           void* stream) {
    EG_REQUIRE(a && y && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0 && (!y_act || leak));
    const long long outs = (long long)N * (H / 2) * (W / 2) * C;
    if (C % 4 == 0 && al16(a) && (!b || al16(b)) && al16(y) && (!y_act || al16(y_act)))
        add_pool2_fwd_k<4><<<grid4(outs), TB, 0, ST>>>(a, b, y, leak, y_act, N, H, W, C);
    else
        add_pool2_fwd_k<1><<<grid4(outs * 4), TB, 0, ST>>>(a, b, y, leak, y_act, N, H, W, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_add_pool2_fwd(const float* a, const float* b, float* y, int N, int H, int W, int C, void* stream) {
    return eg_add_pool2_prelu_fwd(a, b, y, nullptr, nullptr, N, H, W, C, stream);
}
int eg_pool2_bwd(const float* gy, float* gx, int N, int H, int W, int C, int accumulate, void* stream) {
    EG_REQUIRE(gy && gx && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0);
    const long long outs = (long long)N * (H / 2) * (W / 2) * C;
    if (C % 4 == 0 && al16(gy) && al16(gx)) pool2_bwd_k<4><<<grid4(outs), TB, 0, ST>>>(gy, gx, N, H, W, C, accumulate);
    else pool2_bwd_k<1><<<grid4(outs * 4), TB, 0, ST>>>(gy, gx, N, H, W, C, accumulate);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_globalmean_fwd(const float* x, float* y, int N, int P, int C, void* stream) {
    EG_REQUIRE(x && y && N > 0 && P > 0 && C > 0);
    globalmean_fwd_k<<<grid1d((long long)N * C), TB, 0, ST>>>(x, y, N, P, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_globalmean_bwd(const float* gy, float* gx, int N, int P, int C, void* stream) {
    EG_REQUIRE(gy && gx && N > 0 && P > 0 && C > 0);
    globalmean_bwd_k<<<grid1d((long long)N * P * C), TB, 0, ST>>>(gy, gx, (long long)N * P * C, P, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_spectral_norm_ws_floats(int K, int C) { return 2 * K + C + 8; }
int eg_spectral_norm_fwd(const float* W, const float* u, float* Wbar, float* ws, int K, int C, void* stream) {
    EG_REQUIRE(W && u && Wbar && ws && K > 0 && C > 0);
    float *p = ws, *r = ws + K, *scal = ws + K + C;
    cudaError_t e = cudaMemsetAsync(r, 0, sizeof(float) * (C + 8), ST);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    sn_p_k<<<eg_ceil_div(K, 8), 256, 0, ST>>>(W, u, p, K, C);
    EG_CHECK_LAUNCH();
    const int kslab = 64;
    dim3 grid(eg_ceil_div(C, 256), eg_ceil_div(K, kslab));
    sn_r_k<<<grid, 256, 0, ST>>>(W, p, r, K, C, kslab);
    EG_CHECK_LAUNCH();
    const long long n = (long long)K * C;
    unsigned g = grid1d(n, 256 * 4);
    if (g > 1184) g = 1184;
    sn_scale_k<<<g, 256, 0, ST>>>(W, p, r, scal, Wbar, n, K, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_spectral_norm_bwd(const float* W, const float* u, float* ws, const float* Gbar, float* gW, int K, int C,
                         void* stream) {
    EG_REQUIRE(W && u && ws && Gbar && gW && K > 0 && C > 0);
    float *p = ws, *r = ws + K, *scal = ws + K + C, *gv = ws + K + C + 8;
    cudaError_t e = cudaMemsetAsync(scal + 3, 0, sizeof(float), ST);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    const long long n = (long long)K * C;
    unsigned g = grid1d(n, 256 * 4);
    if (g > 1184) g = 1184;
    sn_dot_k<<<g, 256, 0, ST>>>(Gbar, W, scal + 3, n);
    EG_CHECK_LAUNCH();
    sn_gv_k<<<eg_ceil_div(K, 8), 256, 0, ST>>>(W, r, scal, gv, K, C);
    EG_CHECK_LAUNCH();
    sn_gw_k<<<g, 256, 0, ST>>>(W, u, p, r, scal, gv, Gbar, gW, K, C);
    EG_CHECK_LAUNCH(); return 0;
}
// ---- spectral norm of a whole network: table on the device, 3 launches forward, memset + 3 launches backward ----------
int eg_spectral_norm_set_create(const eg_sn_desc* descs, int n, long long* handle) {
    EG_REQUIRE(descs && handle && n > 0);
    static_assert(sizeof(eg_sn_desc) == sizeof(SnDesc), "eg_sn_desc and the device table entry must have the same layout");
    SnSet* s = new SnSet();
    s->n = n; s->Kmax = 0; s->Cmax = 0; s->nmax = 0; s->bwd_ok = true;
    for (int i = 0; i < n; ++i) {
        const eg_sn_desc& d = descs[i];
        const bool split = d.Wa || d.Wi || d.Ga || d.Gi;
        if (!(d.W && d.u && d.Wbar && d.ws && d.K > 0 && d.C > 0) ||
            (split && !(d.Wa && d.Wi && d.cin > 0 && d.hd > 0 && d.hd < d.cin && d.K % d.cin == 0))) {
            delete s;
            return eg_fail_arg("eg_spectral_norm_set_create: bad descriptor", __FILE__, __LINE__);
        }
        if (!(d.gW && (d.G || (d.Ga && d.Gi && split)))) s->bwd_ok = false;
        s->Kmax = std::max(s->Kmax, d.K); s->Cmax = std::max(s->Cmax, d.C);
        s->nmax = std::max(s->nmax, (long long)d.K * d.C);
    }
    cudaError_t e = cudaMalloc(&s->tab, sizeof(SnDesc) * n);
    if (e == cudaSuccess) e = cudaMalloc(&s->dots, sizeof(float) * n);
    if (e == cudaSuccess) e = cudaMemcpy(s->tab, descs, sizeof(SnDesc) * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete s; return eg_fail(e, __FILE__, __LINE__); }
    std::lock_guard<std::mutex> lk(g_sn_mu);
    g_sn_sets.push_back(s);
    *handle = (long long)g_sn_sets.size() - 1;
    return 0;
}
static SnSet* sn_get(long long h) {
    std::lock_guard<std::mutex> lk(g_sn_mu);
    return (h >= 0 && h < (long long)g_sn_sets.size()) ? g_sn_sets[(size_t)h] : nullptr;
}
int eg_spectral_norm_set_destroy(long long handle) {
    std::lock_guard<std::mutex> lk(g_sn_mu);
    if (handle < 0 || handle >= (long long)g_sn_sets.size() || !g_sn_sets[(size_t)handle]) return 0;
    SnSet* s = g_sn_sets[(size_t)handle];
    g_sn_sets[(size_t)handle] = nullptr;
    cudaFree(s->tab); cudaFree(s->dots);
    delete s;
    return 0;
}
static unsigned sn_host_blocks(long long n) {
    long long g = (n + 1023) / 1024;
    return (unsigned)(g > 1184 ? 1184 : (g < 1 ? 1 : g));
}
int eg_spectral_norm_set_fwd(long long handle, void* stream) {
    SnSet* s = sn_get(handle);
    EG_REQUIRE(s != nullptr);
    sn_set_p_k<<<dim3(eg_ceil_div(s->Kmax, 8), s->n), 256, 0, ST>>>(s->tab);
    EG_CHECK_LAUNCH();
    const int kslab = 64;
    sn_set_r_k<<<dim3(eg_ceil_div(s->Cmax, 256), eg_ceil_div(s->Kmax, kslab), s->n), 256, 0, ST>>>(s->tab, kslab);
    EG_CHECK_LAUNCH();
    sn_set_scale_k<<<dim3(sn_host_blocks(s->nmax), s->n), 256, 0, ST>>>(s->tab);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_spectral_norm_set_bwd(long long handle, void* stream) {
    SnSet* s = sn_get(handle);
    EG_REQUIRE(s != nullptr && s->bwd_ok);
    cudaError_t e = cudaMemsetAsync(s->dots, 0, sizeof(float) * s->n, ST);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    const dim3 g(sn_host_blocks(s->nmax), s->n);
    sn_set_dot_k<<<g, 256, 0, ST>>>(s->tab, s->dots);
    EG_CHECK_LAUNCH();
    sn_set_gv_k<<<dim3(eg_ceil_div(s->Kmax, 8), s->n), 256, 0, ST>>>(s->tab);
    EG_CHECK_LAUNCH();
    sn_set_gw_k<<<g, 256, 0, ST>>>(s->tab, s->dots);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_softmax_ce_bwd(const float* logits, const float* z, int z_stride, int label_col, int B, int C, int focal,
                      float weight, float inv_global_batch, float* glogits, float* loss, void* stream) {
    EG_REQUIRE(logits && z && glogits && loss && B > 0 && C > 0 && label_col < z_stride);
    softmax_ce_bwd_k<<<grid1d(B, 64), 64, 0, ST>>>(logits, z, z_stride, label_col, B, C, focal, weight, inv_global_batch, glogits, loss);
    EG_CHECK_LAUNCH(); return 0;
}

}  // extern "C"
