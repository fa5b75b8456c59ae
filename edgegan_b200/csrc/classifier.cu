// Kernels specific to the multi-class classifier D2 (reference models/classifier.py:12-119, the MRU block
// nn/modules/conv.py:133-243, spectral norm nn/modules/normalization.py:38-76, focal / CE losses
// nn/functional.py:5-16).  All HBM-bound glue; the classifier's convolutions use the shared conv trio.
#include "common.cuh"

namespace {

constexpr int TB = 256;
inline unsigned grid1d(long long n, int per_block = TB) { return (unsigned)((n + per_block - 1) / per_block); }

// ---- prelu: tf.maximum(leak * x, x), learned scalar leak (activation.py:23-27; tie -> first argument) ----
__global__ void prelu_fwd_k(const float* __restrict__ x, const float* __restrict__ leak, float* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float l = leak[0], v = x[i], a = l * v;
    y[i] = a >= v ? a : v;
}
__global__ void __launch_bounds__(TB)
prelu_bwd_k(const float* __restrict__ x, const float* __restrict__ leak, const float* __restrict__ gy,
            float* __restrict__ gx, float* __restrict__ gleak, long long n) {
    __shared__ float red[33];
    const float l = leak[0];
    float acc = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i], g = gy[i];
        const bool first = (l * v >= v);
        if (gx != nullptr) gx[i] = first ? g * l : g;
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

__global__ void fma3_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                       float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fmaf(b[i], c[i], a[i]);
}
__global__ void mul_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
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
__global__ void add_pool2_fwd_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
                                int N, int H, int W, int C) {
    const int OH = H / 2, OW = W / 2;
    const long long total = (long long)N * OH * OW * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int ox = (int)(t % OW); t /= OW;
    const int oy = (int)(t % OH); const int n = (int)(t / OH);
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const size_t j = (((size_t)n * H + 2 * oy + dy) * W + 2 * ox + dx) * C + c;
            acc += a[j] + (b ? b[j] : 0.f);
        }
    y[i] = acc * 0.25f;
}
// gx (= or +=) gy / 4 broadcast over each 2x2 window
__global__ void pool2_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, int N, int H, int W, int C, int accumulate) {
    const long long total = (long long)N * H * W * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int ix = (int)(t % W); t /= W;
    const int iy = (int)(t % H); const int n = (int)(t / H);
    const float v = 0.25f * gy[(((size_t)n * (H / 2) + iy / 2) * (W / 2) + ix / 2) * C + c];
    gx[i] = accumulate ? gx[i] + v : v;
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
    prelu_fwd_k<<<grid1d(n), TB, 0, ST>>>(x, leak, y, n);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_prelu_bwd(const float* x, const float* leak, const float* gy, float* gx, float* gleak, long long n,
                 int accumulate_leak, void* stream) {
    EG_REQUIRE(x && leak && gy && n > 0 && (gx || gleak));
    if (gleak && !accumulate_leak) {
        cudaError_t e = cudaMemsetAsync(gleak, 0, sizeof(float), ST);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    unsigned g = grid1d(n, TB * 8);
    if (g > 2048) g = 2048;
    prelu_bwd_k<<<g, TB, 0, ST>>>(x, leak, gy, gx, gleak, n);
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
    fma3_k<<<grid1d(n), TB, 0, ST>>>(a, b, c, out, n);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_mul(const float* a, const float* b, float* out, long long n, void* stream) {
    EG_REQUIRE(a && b && out && n > 0);
    mul_k<<<grid1d(n), TB, 0, ST>>>(a, b, out, n);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_add_pool2_fwd(const float* a, const float* b, float* y, int N, int H, int W, int C, void* stream) {
    EG_REQUIRE(a && y && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0);
    add_pool2_fwd_k<<<grid1d((long long)N * (H / 2) * (W / 2) * C), TB, 0, ST>>>(a, b, y, N, H, W, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_pool2_bwd(const float* gy, float* gx, int N, int H, int W, int C, int accumulate, void* stream) {
    EG_REQUIRE(gy && gx && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0);
    pool2_bwd_k<<<grid1d((long long)N * H * W * C), TB, 0, ST>>>(gy, gx, N, H, W, C, accumulate);
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
int eg_softmax_ce_bwd(const float* logits, const float* z, int z_stride, int label_col, int B, int C, int focal,
                      float weight, float inv_global_batch, float* glogits, float* loss, void* stream) {
    EG_REQUIRE(logits && z && glogits && loss && B > 0 && C > 0 && label_col < z_stride);
    softmax_ce_bwd_k<<<grid1d(B, 64), 64, 0, ST>>>(logits, z, z_stride, label_col, B, C, focal, weight, inv_global_batch, glogits, loss);
    EG_CHECK_LAUNCH(); return 0;
}

}  // extern "C"
