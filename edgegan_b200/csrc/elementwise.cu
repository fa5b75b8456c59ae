// HBM-bound glue kernels of the EdgeGAN step: activations, legacy bicubic 2x resize, width slices /
// concat, the WGAN-GP interpolate / norm / seed kernels, encoder pooling + reflect padding, the
// discriminator's [F,1] linear head, loss reductions and the RMSProp update.
//
// Reference call sites are cited per entry point in include/edgegan_b200.h.
#include "common.cuh"

namespace {

constexpr int TB = 256;

inline unsigned grid1d(long long n, int per_block = TB) {
    long long g = (n + per_block - 1) / per_block;
    return (unsigned)(g < 1 ? 1 : g);
}
// float4 streaming for the flat element-wise kernels: 16 bytes per access, grid-stride, scalar tail
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__host__ __device__ __forceinline__ bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline unsigned grid4(long long n) {
    long long g = (n / 4 + TB * 4 - 1) / (TB * 4);
    if (g < 1) g = 1;
    if (g > 148 * 16) g = 148 * 16;
    return (unsigned)g;
}
#define EG_FLAT_LOOP(n, vec)                                                                                       \
    const long long tid_ = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt_ = (long long)gridDim.x * blockDim.x; \
    const long long n4_ = (vec) ? (n) / 4 : 0

__global__ void __launch_bounds__(TB) act_fwd_k(const float* __restrict__ x, float* __restrict__ y, long long n, int act, int vec) {
    EG_FLAT_LOOP(n, vec);
    for (long long i = tid_; i < n4_; i += nt_) {
        const float4 v = ld4(x + 4 * i);
        st4(y + 4 * i, make_float4(act_fwd(act, v.x), act_fwd(act, v.y), act_fwd(act, v.z), act_fwd(act, v.w)));
    }
    for (long long i = 4 * n4_ + tid_; i < n; i += nt_) y[i] = act_fwd(act, x[i]);
}
__global__ void __launch_bounds__(TB) act_bwd_k(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                                long long n, int act, int vec) {
    EG_FLAT_LOOP(n, vec);
    for (long long i = tid_; i < n4_; i += nt_) {
        const float4 v = ld4(x + 4 * i), g = ld4(gy + 4 * i);
        st4(gx + 4 * i, make_float4(g.x * act_grad(act, v.x), g.y * act_grad(act, v.y), g.z * act_grad(act, v.z), g.w * act_grad(act, v.w)));
    }
    for (long long i = 4 * n4_ + tid_; i < n; i += nt_) gx[i] = gy[i] * act_grad(act, x[i]);
}

// ---- legacy bicubic 2x (A = -0.75, no half-pixel offset, clamped borders) ---------------
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// forward taps of output index o: up to 4 (index, weight)
__device__ __forceinline__ int up2_taps(int o, int n, int* idx, float* w) {
    const int i = o >> 1;
    if ((o & 1) == 0) { idx[0] = i; w[0] = 1.f; return 1; }
    idx[0] = clampi(i - 1, 0, n - 1); w[0] = -0.09375f;
    idx[1] = clampi(i, 0, n - 1);     w[1] = 0.59375f;
    idx[2] = clampi(i + 1, 0, n - 1); w[2] = 0.59375f;
    idx[3] = clampi(i + 2, 0, n - 1); w[3] = -0.09375f;
    return 4;
}
// transposed taps of input index i: 5 candidate outputs (index, accumulated weight; weight 0 if unused)
__device__ __forceinline__ void up2_taps_t(int i, int n, int* o, float* w) {
    o[0] = 2 * i; w[0] = 1.f;
    const float tw[4] = {-0.09375f, 0.59375f, 0.59375f, -0.09375f};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const int j = i - 2 + d;      // odd output 2j+1 reads j-1 .. j+2 (clamped)
        float acc = 0.f;
        if (j >= 0 && j < n) {
#pragma unroll
            for (int a = -1; a <= 2; ++a) if (clampi(j + a, 0, n - 1) == i) acc += tw[a + 1];
        }
        o[1 + d] = (j >= 0 && j < n) ? 2 * j + 1 : 0;
        w[1 + d] = acc;
    }
}

__global__ void bicubic_up2_fwd_k(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C) {
    const long long total = (long long)N * 2 * H * 2 * W * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int ox = (int)(t % (2 * W)); t /= 2 * W;
    const int oy = (int)(t % (2 * H)); const int n = (int)(t / (2 * H));
    int iy[4], ix[4]; float wy[4], wx[4];
    const int ny = up2_taps(oy, H, iy, wy), nx = up2_taps(ox, W, ix, wx);
    // rows first, then columns -- same association as the separable reference
    float acc = 0.f;
    for (int b = 0; b < nx; ++b) {
        float col = 0.f;
        for (int a = 0; a < ny; ++a) col = fmaf(wy[a], x[(((size_t)n * H + iy[a]) * W + ix[b]) * C + c], col);
        acc = fmaf(wx[b], col, acc);
    }
    y[i] = acc;
}

// the same, one thread per output PIXEL (C <= 4 channels: the image tensors in front of the patch critics): the tap
// indices / weights and the 32-bit index arithmetic are shared by the channels
template <int C>
__global__ void __launch_bounds__(256) bicubic_up2_fwd_px_k(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W) {
    const int total = N * 2 * H * 2 * W;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ox = i % (2 * W); int t = i / (2 * W);
    const int oy = t % (2 * H); const int n = t / (2 * H);
    int iy[4], ix[4]; float wy[4], wx[4];
    const int ny = up2_taps(oy, H, iy, wy), nx = up2_taps(ox, W, ix, wx);
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = 0.f;
    for (int b = 0; b < nx; ++b) {
        float col[C];
#pragma unroll
        for (int c = 0; c < C; ++c) col[c] = 0.f;
        for (int a = 0; a < ny; ++a) {
            const float* src = x + ((size_t)(n * H + iy[a]) * W + ix[b]) * C;
#pragma unroll
            for (int c = 0; c < C; ++c) col[c] = fmaf(wy[a], __ldg(src + c), col[c]);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(wx[b], col[c], acc[c]);
    }
    float* dst = y + (size_t)i * C;
#pragma unroll
    for (int c = 0; c < C; ++c) dst[c] = acc[c];
}

__global__ void bicubic_up2_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, int N, int H, int W, int C) {
    const long long total = (long long)N * H * W * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int ix = (int)(t % W); t /= W;
    const int iy = (int)(t % H); const int n = (int)(t / H);
    int oy[5], ox[5]; float wy[5], wx[5];
    up2_taps_t(iy, H, oy, wy); up2_taps_t(ix, W, ox, wx);
    float acc = 0.f;
    for (int a = 0; a < 5; ++a) {
        if (wy[a] == 0.f) continue;
        float row = 0.f;
        for (int b = 0; b < 5; ++b) {
            if (wx[b] == 0.f) continue;
            row = fmaf(wx[b], gy[(((size_t)n * 2 * H + oy[a]) * 2 * W + ox[b]) * C + c], row);
        }
        acc = fmaf(wy[a], row, acc);
    }
    gx[i] = acc;
}

__global__ void copy2d_k(const float* __restrict__ src, long long ss, float* __restrict__ dst, long long ds,
                         long long rows, long long cols) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const long long r = i / cols, c = i % cols;
    dst[r * ds + c] = src[r * ss + c];
}
__global__ void __launch_bounds__(TB) fill_k(float* dst, long long n, float v, int vec) {
    EG_FLAT_LOOP(n, vec);
    for (long long i = tid_; i < n4_; i += nt_) st4(dst + 4 * i, make_float4(v, v, v, v));
    for (long long i = 4 * n4_ + tid_; i < n; i += nt_) dst[i] = v;
}
// dst[i] = lut[src[i]]: the loader uploads image bytes and the byte -> [-1, 1] table of utils.transform
// (edgegan/utils/utils.py:160: x / 127.5 - 1); 16 bytes in, 64 bytes out per thread, table in shared memory
__global__ void u8_lut_f32_k(const uint8_t* __restrict__ src, const float* __restrict__ lut, float* __restrict__ dst, long long n) {
    __shared__ float tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tab[i] = lut[i];
    __syncthreads();
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (i0 >= n) return;
    if (i0 + 16 <= n && (reinterpret_cast<uintptr_t>(src) & 15u) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(src + i0);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            *reinterpret_cast<float4*>(dst + i0 + 4 * k) =
                make_float4(tab[w[k] & 255u], tab[(w[k] >> 8) & 255u], tab[(w[k] >> 16) & 255u], tab[w[k] >> 24]);
    } else {
        for (long long i = i0; i < n && i < i0 + 16; ++i) dst[i] = tab[src[i]];
    }
}
__global__ void __launch_bounds__(TB) axpby_k(const float* __restrict__ x, float* __restrict__ y, long long n, float a, float b, int vec) {
    EG_FLAT_LOOP(n, vec);
    for (long long i = tid_; i < n4_; i += nt_) {
        const float4 u = ld4(x + 4 * i);
        float4 o = make_float4(a * u.x, a * u.y, a * u.z, a * u.w);
        if (b != 0.f) { const float4 w = ld4(y + 4 * i); o.x += b * w.x; o.y += b * w.y; o.z += b * w.z; o.w += b * w.w; }
        st4(y + 4 * i, o);
    }
    for (long long i = 4 * n4_ + tid_; i < n; i += nt_) y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}

// ---- WGAN-GP ---------------------------------------------------------------------------------------
__global__ void gp_interpolate_k(const float* __restrict__ real, const float* __restrict__ fake,
                                 const float* __restrict__ alpha, float* __restrict__ xhat, long long total,
                                 long long per) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float a = alpha[i / per];
    const float r = real[i];
    xhat[i] = r + a * (fake[i] - r);
}
__device__ __forceinline__ float sigmoidf_(float d) { return 1.f / (1.f + expf(-d)); }
__global__ void gp_seed_k(const float* __restrict__ d, float* __restrict__ dd, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { const float s = sigmoidf_(d[i]); dd[i] = 1.f + s * (1.f - s); }
}
__global__ void gp_seed_bwd_k(const float* __restrict__ d, const float* __restrict__ ddbar, float* __restrict__ dbar, int B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { const float s = sigmoidf_(d[i]); dbar[i] = ddbar[i] * s * (1.f - s) * (1.f - 2.f * s); }
}
__global__ void __launch_bounds__(1024)
gp_penalty_k(const float* __restrict__ g, float* __restrict__ gbar, float* __restrict__ norms, float* __restrict__ loss,
             long long per, float weight, float inv_b) {
    __shared__ float red[33];
    const int b = blockIdx.x;
    const float* gp = g + (size_t)b * per;
    float s = 0.f;
    for (long long i = threadIdx.x; i < per; i += blockDim.x) { const float v = gp[i]; s = fmaf(v, v, s); }
    s = block_sum(s, red);
    const float nrm = sqrtf(s);
    const float coef = weight * 2.f * (nrm - 1.f) * inv_b / nrm;
    float* op = gbar + (size_t)b * per;
    for (long long i = threadIdx.x; i < per; i += blockDim.x) op[i] = coef * gp[i];
    if (threadIdx.x == 0) {
        norms[b] = nrm;
        atomicAdd(loss, weight * (nrm - 1.f) * (nrm - 1.f) * inv_b);
    }
}

__global__ void __launch_bounds__(TB)
sum_scaled_k(const float* __restrict__ x, long long n, float scale, float* __restrict__ out, int accumulate) {
    __shared__ float red[33];
    float s = 0.f;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) out[0] = (accumulate ? out[0] : 0.f) + scale * s;
}

// ---- discriminator head ----------------------------------------------------------------------------
__global__ void __launch_bounds__(512)
rowdot_fwd_k(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ bias,
             float* __restrict__ d, int F) {
    __shared__ float red[33];
    const int b = blockIdx.x;
    const float* hp = h + (size_t)b * F;
    float s = 0.f;
    if ((F & 3) == 0 && ((reinterpret_cast<uintptr_t>(hp) | reinterpret_cast<uintptr_t>(w)) & 15) == 0) {
        const float4* h4 = reinterpret_cast<const float4*>(hp);
        const float4* w4 = reinterpret_cast<const float4*>(w);
        for (int i = threadIdx.x; i < F / 4; i += blockDim.x) {
            const float4 a = __ldg(h4 + i), c = __ldg(w4 + i);
            s = fmaf(a.x, c.x, s); s = fmaf(a.y, c.y, s); s = fmaf(a.z, c.z, s); s = fmaf(a.w, c.w, s);
        }
    } else {
        for (int i = threadIdx.x; i < F; i += blockDim.x) s = fmaf(hp[i], w[i], s);
    }
    s = block_sum(s, red);
    if (threadIdx.x == 0) d[b] = s + (bias ? bias[0] : 0.f);
}
__global__ void rowdot_bwd_input_k(const float* __restrict__ gd, const float* __restrict__ w, float* __restrict__ gh,
                                   long long total, int F) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) gh[i] = gd[i / F] * w[i % F];
}
__global__ void rowdot_bwd_weight_k(const float* __restrict__ gd, const float* __restrict__ h, float* __restrict__ gw,
                                    float* __restrict__ gb, int B, int F, int accumulate) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < F) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s = fmaf(gd[b], h[(size_t)b * F + f], s);
        gw[f] = (accumulate ? gw[f] : 0.f) + s;
    }
    if (f == 0 && gb != nullptr) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += gd[b];
        gb[0] = (accumulate ? gb[0] : 0.f) + s;
    }
}

// db[c] += sum over a slab of rows
__global__ void __launch_bounds__(256)
bias_grad_k(const float* __restrict__ dy, long long rows, int C, float* __restrict__ db, int rows_per_block) {
    __shared__ float sm[8][32];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    const long long r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    if (c < C) for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += dy[r * C + c];
    sm[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][threadIdx.x];
        atomicAdd(db + c, t);
    }
}

// ---- encoder pieces --------------------------------------------------------------------------------
__device__ __forceinline__ int reflecti(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * (n - 1) - i : i); }

__global__ void reflect_pad_fwd_k(const float* __restrict__ x, float* __restrict__ y, int N, int H, int W, int C, int p) {
    const int HP = H + 2 * p, WP = W + 2 * p;
    const long long total = (long long)N * HP * WP * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int xx = (int)(t % WP); t /= WP;
    const int yy = (int)(t % HP); const int n = (int)(t / HP);
    y[i] = x[(((size_t)n * H + reflecti(yy - p, H)) * W + reflecti(xx - p, W)) * C + c];
}
// candidates of padded coordinates that read source index i
__device__ __forceinline__ int reflect_sources(int i, int n, int p, int* out) {
    int k = 0;
    out[k++] = i + p;
    if (i >= 1 && i <= p) out[k++] = p - i;
    if (i <= n - 2 && i >= n - 1 - p) out[k++] = 2 * (n - 1) - i + p;
    return k;
}
__global__ void reflect_pad_bwd_k(const float* __restrict__ gy, float* __restrict__ gx, int N, int H, int W, int C, int p) {
    const int HP = H + 2 * p, WP = W + 2 * p;
    const long long total = (long long)N * H * W * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int ix = (int)(t % W); t /= W;
    const int iy = (int)(t % H); const int n = (int)(t / H);
    int ys[3], xs[3];
    const int ny = reflect_sources(iy, H, p, ys), nx = reflect_sources(ix, W, p, xs);
    float acc = 0.f;
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) acc += gy[(((size_t)n * HP + ys[a]) * WP + xs[b]) * C + c];
    gx[i] = acc;
}

// 16-byte versions (C % 4 == 0, < 2^31 float4 elements): the encoder's maps have 64+ channels
__global__ void __launch_bounds__(256) reflect_pad_fwd_v4_k(const float4* __restrict__ x, float4* __restrict__ y, int N, int H, int W, int C4, int p) {
    const int HP = H + 2 * p, WP = W + 2 * p;
    const unsigned total = (unsigned)N * HP * WP * C4;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned c = i % C4; unsigned t = i / C4;
    const int xx = (int)(t % WP); t /= WP;
    const int yy = (int)(t % HP); const int n = (int)(t / HP);
    y[i] = __ldg(x + ((size_t)(n * H + reflecti(yy - p, H)) * W + reflecti(xx - p, W)) * C4 + c);
}
__global__ void __launch_bounds__(256) reflect_pad_bwd_v4_k(const float4* __restrict__ gy, float4* __restrict__ gx, int N, int H, int W, int C4, int p) {
    const int HP = H + 2 * p, WP = W + 2 * p;
    const unsigned total = (unsigned)N * H * W * C4;
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const unsigned c = i % C4; unsigned t = i / C4;
    const int ix = (int)(t % W); t /= W;
    const int iy = (int)(t % H); const int n = (int)(t / H);
    int ys[3], xs[3];
    const int ny = reflect_sources(iy, H, p, ys), nx = reflect_sources(ix, W, p, xs);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) {
            const float4 v = __ldg(gy + ((size_t)(n * HP + ys[a]) * WP + xs[b]) * C4 + c);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    gx[i] = acc;
}

__global__ void addrelu_pool2_fwd_k(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y,
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
            float v = a[j]; if (b) v += b[j];
            acc += fmaxf(v, 0.f);
        }
    y[i] = acc * 0.25f;
}
__global__ void addrelu_pool2_bwd_k(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gy,
                                    float* __restrict__ g, int N, int H, int W, int C) {
    const long long total = (long long)N * H * W * C;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); long long t = i / C;
    const int ix = (int)(t % W); t /= W;
    const int iy = (int)(t % H); const int n = (int)(t / H);
    float v = a[i]; if (b) v += b[i];
    g[i] = v > 0.f ? 0.25f * gy[(((size_t)n * (H / 2) + iy / 2) * (W / 2) + ix / 2) * C + c] : 0.f;
}
__global__ void relu_globalmean_fwd_k(const float* __restrict__ x, float* __restrict__ y, int N, int P, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * C) return;
    const int c = i % C, n = i / C;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += fmaxf(x[((size_t)n * P + p) * C + c], 0.f);
    y[i] = s / P;
}
__global__ void relu_globalmean_bwd_k(const float* __restrict__ x, const float* __restrict__ gy, float* __restrict__ gx,
                                      long long total, int P, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i % C); const long long n = i / ((long long)P * C);
    gx[i] = x[i] > 0.f ? gy[n * C + c] / P : 0.f;
}
__global__ void reparam_fwd_k(const float* __restrict__ mu, const float* __restrict__ ls, float eps, const float* __restrict__ eps_dev,
                              float* __restrict__ z, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (eps_dev != nullptr) eps = eps_dev[0];
    if (i < n) z[i] = mu[i] + eps * expf(ls[i]);
}
__global__ void __launch_bounds__(TB)
zl1_loss_bwd_k(const float* __restrict__ mu, const float* __restrict__ ls, float eps, const float* __restrict__ eps_dev, const float* __restrict__ target,
               int tstride, int B, int Z, float weight, float inv_count, float* __restrict__ gmu,
               float* __restrict__ gls, float* __restrict__ loss) {
    __shared__ float red[33];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float l = 0.f;
    if (eps_dev != nullptr) eps = eps_dev[0];
    if (i < B * Z) {
        const int b = i / Z, j = i % Z;
        const float e = expf(ls[i]);
        const float z = mu[i] + eps * e;
        const float diff = target[(size_t)b * tstride + j] - z;
        l = fabsf(diff);
        const float sg = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
        const float gz = -weight * inv_count * sg;
        gmu[i] = gz;
        gls[i] = gz * eps * e;
    }
    l = block_sum(l, red);
    if (threadIdx.x == 0) atomicAdd(loss, weight * inv_count * l);
}

__global__ void onehot_concat_k(const float* __restrict__ z, int zdim, int classes, float* __restrict__ out, int n) {
    const int W = zdim + classes;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * W) return;
    const int b = i / W, j = i % W;
    if (j < zdim) out[i] = z[(size_t)b * (zdim + 1) + j];
    else out[i] = ((int)z[(size_t)b * (zdim + 1) + zdim] == j - zdim) ? 1.f : 0.f;   // tf.cast(float->int32) truncates
}

__device__ __forceinline__ void rmsprop1(float& v, float g, float& m, float lr, float decay, float eps) {
    m = decay * m + (1.f - decay) * g * g;
    v -= lr * g / sqrtf(m + eps);
}
__global__ void __launch_bounds__(TB) rmsprop_k(float* __restrict__ var, const float* __restrict__ grad, float* __restrict__ ms, long long n,
                                                float lr, float decay, float eps, int vec) {
    EG_FLAT_LOOP(n, vec);
    for (long long i = tid_; i < n4_; i += nt_) {
        const float4 g = ld4(grad + 4 * i);
        float4 m = ld4(ms + 4 * i), v = ld4(var + 4 * i);
        rmsprop1(v.x, g.x, m.x, lr, decay, eps); rmsprop1(v.y, g.y, m.y, lr, decay, eps);
        rmsprop1(v.z, g.z, m.z, lr, decay, eps); rmsprop1(v.w, g.w, m.w, lr, decay, eps);
        st4(ms + 4 * i, m); st4(var + 4 * i, v);
    }
    for (long long i = 4 * n4_ + tid_; i < n; i += nt_) rmsprop1(var[i], grad[i], ms[i], lr, decay, eps);
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" {

int eg_act_fwd(const float* x, float* y, long long n, int act, void* stream) {
    EG_REQUIRE(x && y && n > 0);
    act_fwd_k<<<grid4(n), TB, 0, ST>>>(x, y, n, act, al16(x) && al16(y));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_act_bwd(const float* x_pre, const float* gy, float* gx, long long n, int act, void* stream) {
    EG_REQUIRE(x_pre && gy && gx && n > 0);
    act_bwd_k<<<grid4(n), TB, 0, ST>>>(x_pre, gy, gx, n, act, al16(x_pre) && al16(gy) && al16(gx));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_bicubic_up2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream) {
    EG_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0);
    const long long px = (long long)N * 4 * H * W;
    if (C == 3 && px < (1ll << 31) && px * C < (1ll << 31)) bicubic_up2_fwd_px_k<3><<<grid1d(px, 256), 256, 0, ST>>>(x, y, N, H, W);
    else if (C == 1 && px < (1ll << 31)) bicubic_up2_fwd_px_k<1><<<grid1d(px, 256), 256, 0, ST>>>(x, y, N, H, W);
    else bicubic_up2_fwd_k<<<grid1d((long long)N * 4 * H * W * C), TB, 0, ST>>>(x, y, N, H, W, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_bicubic_up2_bwd(const float* gy, float* gx, int N, int H, int W, int C, void* stream) {
    EG_REQUIRE(gy && gx && N > 0 && H > 0 && W > 0 && C > 0);
    bicubic_up2_bwd_k<<<grid1d((long long)N * H * W * C), TB, 0, ST>>>(gy, gx, N, H, W, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_copy2d(const float* src, long long src_stride, float* dst, long long dst_stride, long long rows,
              long long cols, void* stream) {
    EG_REQUIRE(src && dst && rows > 0 && cols > 0);
    copy2d_k<<<grid1d(rows * cols), TB, 0, ST>>>(src, src_stride, dst, dst_stride, rows, cols);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_fill(float* dst, long long n, float value, void* stream) {
    EG_REQUIRE(dst && n > 0);
    fill_k<<<grid4(n), TB, 0, ST>>>(dst, n, value, al16(dst));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_u8_lut_f32(const void* src, const float* lut, float* dst, long long n, void* stream) {
    EG_REQUIRE(src && lut && dst && n > 0);
    u8_lut_f32_k<<<grid1d((n + 15) / 16), TB, 0, ST>>>(static_cast<const uint8_t*>(src), lut, dst, n);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_axpby(const float* x, float* y, long long n, float a, float b, void* stream) {
    EG_REQUIRE(x && y && n > 0);
    axpby_k<<<grid4(n), TB, 0, ST>>>(x, y, n, a, b, al16(x) && al16(y));
    EG_CHECK_LAUNCH(); return 0;
}
int eg_gp_interpolate(const float* real, const float* fake, const float* alpha, float* xhat, int B, long long per,
                      void* stream) {
    EG_REQUIRE(real && fake && alpha && xhat && B > 0 && per > 0);
    gp_interpolate_k<<<grid1d((long long)B * per), TB, 0, ST>>>(real, fake, alpha, xhat, (long long)B * per, per);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_gp_seed(const float* d, float* dd, int B, void* stream) {
    EG_REQUIRE(d && dd && B > 0);
    gp_seed_k<<<grid1d(B), TB, 0, ST>>>(d, dd, B);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_gp_seed_bwd(const float* d, const float* ddbar, float* dbar, int B, void* stream) {
    EG_REQUIRE(d && ddbar && dbar && B > 0);
    gp_seed_bwd_k<<<grid1d(B), TB, 0, ST>>>(d, ddbar, dbar, B);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_gp_penalty(const float* g, float* gbar, float* norms, float* loss, int B, long long per, float weight,
                  float inv_global_batch, void* stream) {
    EG_REQUIRE(g && gbar && norms && loss && B > 0 && per > 0);
    gp_penalty_k<<<B, 1024, 0, ST>>>(g, gbar, norms, loss, per, weight, inv_global_batch);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_sum_scaled(const float* x, long long n, float scale, float* out, int accumulate, void* stream) {
    EG_REQUIRE(x && out && n > 0);
    sum_scaled_k<<<1, TB, 0, ST>>>(x, n, scale, out, accumulate);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_rowdot_fwd(const float* h, const float* w, const float* bias, float* d, int B, int F, void* stream) {
    EG_REQUIRE(h && w && d && B > 0 && F > 0);
    rowdot_fwd_k<<<B, 512, 0, ST>>>(h, w, bias, d, F);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_rowdot_bwd_input(const float* gd, const float* w, float* gh, int B, int F, void* stream) {
    EG_REQUIRE(gd && w && gh && B > 0 && F > 0);
    rowdot_bwd_input_k<<<grid1d((long long)B * F), TB, 0, ST>>>(gd, w, gh, (long long)B * F, F);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_rowdot_bwd_weight(const float* gd, const float* h, float* gw, float* gb, int B, int F, int accumulate,
                         void* stream) {
    EG_REQUIRE(gd && h && gw && B > 0 && F > 0);
    rowdot_bwd_weight_k<<<grid1d(F, 128), 128, 0, ST>>>(gd, h, gw, gb, B, F, accumulate);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_bias_grad(const float* dy, long long rows, int C, float* db, int accumulate, void* stream) {
    EG_REQUIRE(dy && db && rows > 0 && C > 0);
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(db, 0, sizeof(float) * C, ST);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    const int rpb = 512;
    dim3 grid(eg_ceil_div(C, 32), eg_ceil_div(rows, rpb)), block(32, 8);
    bias_grad_k<<<grid, block, 0, ST>>>(dy, rows, C, db, rpb);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_reflect_pad_fwd(const float* x, float* y, int N, int H, int W, int C, int p, void* stream) {
    EG_REQUIRE(x && y && N > 0 && H > p && W > p && C > 0 && p >= 0);
    const long long n4 = (long long)N * (H + 2 * p) * (W + 2 * p) * (C / 4);
    if (C % 4 == 0 && n4 < (1ll << 31) && al16(x) && al16(y))
        reflect_pad_fwd_v4_k<<<grid1d(n4, 256), 256, 0, ST>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), N, H, W, C / 4, p);
    else
        reflect_pad_fwd_k<<<grid1d((long long)N * (H + 2 * p) * (W + 2 * p) * C), TB, 0, ST>>>(x, y, N, H, W, C, p);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_reflect_pad_bwd(const float* gy, float* gx, int N, int H, int W, int C, int p, void* stream) {
    EG_REQUIRE(gy && gx && N > 0 && H > p && W > p && C > 0 && p >= 0);
    const long long n4 = (long long)N * (H + 2 * p) * (W + 2 * p) * (C / 4);
    if (C % 4 == 0 && n4 < (1ll << 31) && al16(gy) && al16(gx))
        reflect_pad_bwd_v4_k<<<grid1d((long long)N * H * W * (C / 4), 256), 256, 0, ST>>>(reinterpret_cast<const float4*>(gy), reinterpret_cast<float4*>(gx), N, H, W, C / 4, p);
    else
        reflect_pad_bwd_k<<<grid1d((long long)N * H * W * C), TB, 0, ST>>>(gy, gx, N, H, W, C, p);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_addrelu_pool2_fwd(const float* a, const float* b, float* y, int N, int H, int W, int C, void* stream) {
    EG_REQUIRE(a && y && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0);
    addrelu_pool2_fwd_k<<<grid1d((long long)N * (H / 2) * (W / 2) * C), TB, 0, ST>>>(a, b, y, N, H, W, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_addrelu_pool2_bwd(const float* a, const float* b, const float* gy, float* g, int N, int H, int W, int C,
                         void* stream) {
    EG_REQUIRE(a && gy && g && N > 0 && H > 0 && W > 0 && C > 0 && H % 2 == 0 && W % 2 == 0);
    addrelu_pool2_bwd_k<<<grid1d((long long)N * H * W * C), TB, 0, ST>>>(a, b, gy, g, N, H, W, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_relu_globalmean_fwd(const float* x, float* y, int N, int P, int C, void* stream) {
    EG_REQUIRE(x && y && N > 0 && P > 0 && C > 0);
    relu_globalmean_fwd_k<<<grid1d((long long)N * C), TB, 0, ST>>>(x, y, N, P, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_relu_globalmean_bwd(const float* x, const float* gy, float* gx, int N, int P, int C, void* stream) {
    EG_REQUIRE(x && gy && gx && N > 0 && P > 0 && C > 0);
    relu_globalmean_bwd_k<<<grid1d((long long)N * P * C), TB, 0, ST>>>(x, gy, gx, (long long)N * P * C, P, C);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_reparam_fwd(const float* mu, const float* ls, float eps, const float* eps_dev, float* z, long long n, void* stream) {
    EG_REQUIRE(mu && ls && z && n > 0);
    reparam_fwd_k<<<grid1d(n), TB, 0, ST>>>(mu, ls, eps, eps_dev, z, n);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_zl1_loss_bwd(const float* mu, const float* ls, float eps, const float* eps_dev, const float* target, int target_stride,
                    int B, int Z, float weight, float inv_global_count, float* gmu, float* gls, float* loss, void* stream) {
    EG_REQUIRE(mu && ls && target && gmu && gls && loss && B > 0 && Z > 0 && target_stride >= Z);
    zl1_loss_bwd_k<<<grid1d((long long)B * Z), TB, 0, ST>>>(mu, ls, eps, eps_dev, target, target_stride, B, Z, weight,
                                                           inv_global_count, gmu, gls, loss);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_onehot_concat(const float* z, int n, int zdim, int classes, float* out, void* stream) {
    EG_REQUIRE(z && out && n > 0 && zdim > 0 && classes > 0);
    onehot_concat_k<<<grid1d((long long)n * (zdim + classes)), TB, 0, ST>>>(z, zdim, classes, out, n);
    EG_CHECK_LAUNCH(); return 0;
}
int eg_rmsprop(float* var, const float* grad, float* ms, long long n, float lr, float decay, float eps, void* stream) {
    EG_REQUIRE(var && grad && ms && n > 0);
    rmsprop_k<<<grid4(n), TB, 0, ST>>>(var, grad, ms, n, lr, decay, eps, al16(var) && al16(grad) && al16(ms));
    EG_CHECK_LAUNCH(); return 0;
}

}  // extern "C"
