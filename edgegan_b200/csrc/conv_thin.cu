// Thin-channel convolutions (Ci <= 8: the image-side layers of the critics, the encoder and the generator) on the
// tensor-core path.
//
// Replaces for those layers: tf.nn.conv2d / tf.nn.conv2d_transpose (edgegan/nn/modules/conv.py:29,49-53) and their
// gradients.  With 3 input channels the implicit-GEMM K dimension of one filter tap is 12 bytes -- below TMA's
// 16-byte granularity and far below a 32-channel K slab -- so the tcgen05 conv kernels cannot read the image
// directly, and on the FFMA path these layers ran at ~10 % of their HBM roofline.  Here the K dimension is the
// whole receptive field (KH*KW*Ci = 48 or 75 values, padded to a multiple of 32):
//
//   forward       y  = im2col(x)[P, Kpad] . w[Kpad, Co]                 (P = N*OH*OW output pixels)
//   input grad    dx = col2im( dy[P, Co] . w^T[Co, Npad] ) + bias       (Npad = 64 or 128 >= KH*KW*Ci)
//   filter grad   dw = im2col(x)^T[Kpad, P] . dy[P, Co]
//
// The dense products run on the kernels of conv_tc.cu as 1x1 convolutions over P "images" of one pixel; the
// patch matrices live in a per-stream scratch buffer.  im2col / col2im are one pass over the patch matrix each.
#include "common.cuh"

int eg_tc_scratch(cudaStream_t st, int slot, size_t bytes, float** out);
float* eg_tc_managed_thin(const float* w, size_t n, int mode, int which);
int eg_tc_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int three_x, cudaStream_t st, const EgEpi* epi = nullptr, const eg_conv_shape* scatter = nullptr);
int eg_tc_scatter_supported(const eg_conv_shape* c);
int eg_tc_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int three_x, cudaStream_t st);

namespace {

constexpr int kThinMaxCi = 8;
constexpr int kThinMaxK = 128;

struct ThinP {
    int N, H, W, Ci, OH, OW, KH, KW, stride, pad_t, pad_l;
    int K, Kpad;
};

ThinP make_p(const eg_conv_shape* s) {
    ThinP p;
    p.N = s->N; p.H = s->H; p.W = s->W; p.Ci = s->Ci; p.OH = s->OH; p.OW = s->OW; p.KH = s->KH; p.KW = s->KW;
    p.stride = s->stride; p.pad_t = s->pad_t; p.pad_l = s->pad_l;
    p.K = s->KH * s->KW * s->Ci;
    p.Kpad = (p.K + 31) / 32 * 32;
    return p;
}

// A[q][j] = x[n, oh*s - pad_t + kh, ow*s - pad_l + kw, ci],  j = (kh*KW + kw)*Ci + ci ; zero outside the image and
// for j >= K.  One thread writes 4 consecutive columns (16 B).
__global__ void thin_im2col_k(const float* __restrict__ x, float* __restrict__ A, ThinP p, long long total4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const int g4 = p.Kpad >> 2;
    const long long q = i / g4;
    const int j0 = (int)(i - q * g4) * 4;
    const int ow = (int)(q % p.OW);
    const long long t = q / p.OW;
    const int oh = (int)(t % p.OH), n = (int)(t / p.OH);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int j = j0 + e;
        float val = 0.f;
        if (j < p.K) {
            const int tap = j / p.Ci, ci = j - tap * p.Ci;
            const int kh = tap / p.KW, kw = tap - kh * p.KW;
            const int h = oh * p.stride - p.pad_t + kh, w = ow * p.stride - p.pad_l + kw;
            if (h >= 0 && h < p.H && w >= 0 && w < p.W) val = __ldg(x + (((long long)n * p.H + h) * p.W + w) * p.Ci + ci);
        }
        v[e] = val;
    }
    *reinterpret_cast<float4*>(A + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
}

// dx[n, h, w, ci] = bias[ci] + sum over taps (kh, kw) with h = oh*s - pad_t + kh, w = ow*s - pad_l + kw of
// C[(n, oh, ow)][(kh*KW + kw)*Ci + ci].  One thread per input pixel, all (<= 4) channels.
__global__ void thin_col2im_k(const float* __restrict__ C, const float* __restrict__ bias, float* __restrict__ dx,
                              ThinP p, int Npad, long long pixels) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pixels) return;
    const int w = (int)(i % p.W);
    const long long t = i / p.W;
    const int h = (int)(t % p.H), n = (int)(t / p.H);
    float acc[kThinMaxCi];
#pragma unroll
    for (int c = 0; c < kThinMaxCi; ++c) acc[c] = (bias != nullptr && c < p.Ci) ? __ldg(bias + c) : 0.f;
    for (int kh = 0; kh < p.KH; ++kh) {
        const int hh = h + p.pad_t - kh;
        if (hh < 0 || hh % p.stride) continue;
        const int oh = hh / p.stride;
        if (oh >= p.OH) continue;
        for (int kw = 0; kw < p.KW; ++kw) {
            const int ww = w + p.pad_l - kw;
            if (ww < 0 || ww % p.stride) continue;
            const int ow = ww / p.stride;
            if (ow >= p.OW) continue;
            const float* row = C + (((long long)n * p.OH + oh) * p.OW + ow) * Npad + (kh * p.KW + kw) * p.Ci;
#pragma unroll
            for (int c = 0; c < kThinMaxCi; ++c)
                if (c < p.Ci) acc[c] += __ldg(row + c);
        }
    }
#pragma unroll
    for (int c = 0; c < kThinMaxCi; ++c)
        if (c < p.Ci) dx[i * p.Ci + c] = acc[c];
}

// Same patch matrix, one block per output row (n, oh): the KH input rows it reads are staged in shared memory and
// the column -> (kh, kw, ci) decomposition comes from a table, so the inner loop has no integer division.
template <int KPAD>
__global__ void thin_im2col_row_k(const float* __restrict__ x, float* __restrict__ A, ThinP p) {
    extern __shared__ float xs[];                            // [KH][W * Ci]
    __shared__ int tab[KPAD];                                // j -> kh << 24 | kw << 16 | (kw * Ci + ci), -1 for j >= K
    const int row = blockIdx.x, n = row / p.OH, oh = row - n * p.OH;
    const int rowlen = p.W * p.Ci;
    for (int kh = 0; kh < p.KH; ++kh) {
        const int h = oh * p.stride - p.pad_t + kh;
        const bool in = h >= 0 && h < p.H;
        const float* src = x + ((long long)n * p.H + h) * rowlen;
        for (int i = threadIdx.x; i < rowlen; i += blockDim.x) xs[kh * rowlen + i] = in ? __ldg(src + i) : 0.f;
    }
    for (int j = threadIdx.x; j < KPAD; j += blockDim.x) {
        int t = -1;
        if (j < p.K) {
            const int tap = j / p.Ci, ci = j - tap * p.Ci, kh = tap / p.KW, kw = tap - kh * p.KW;
            t = (kh << 24) | (kw << 16) | (kw * p.Ci + ci);
        }
        tab[j] = t;
    }
    __syncthreads();
    constexpr int G4 = KPAD / 4;
    float* Arow = A + (long long)row * p.OW * KPAD;
    for (int idx = threadIdx.x; idx < p.OW * G4; idx += blockDim.x) {
        const int ow = idx / G4, j0 = (idx - ow * G4) * 4;
        const int wbase = ow * p.stride - p.pad_l;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int t = tab[j0 + e];
            const int w = wbase + ((t >> 16) & 255);
            v[e] = (t >= 0 && w >= 0 && w < p.W) ? xs[(t >> 24) * rowlen + wbase * p.Ci + (t & 0xffff)] : 0.f;
        }
        *reinterpret_cast<float4*>(Arow + (long long)idx * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// col2im with the stride known at compile time: one block (256 threads) per input row (n, h).  The filter rows kh that
// reach row h (at most ceil(KH / S)) each read one row of C pixels; their KW*Ci-column segments are staged in shared
// memory with coalesced (16-byte when KW*Ci % 4 == 0) loads, then the threads -- pixel w x a slice of the channels, so that
// a 64-pixel row still keeps 256 threads busy -- sum their taps from there in a fixed order (deterministic).
template <int S>
__global__ void __launch_bounds__(256)
thin_col2im_row_k(const float* __restrict__ C, const float* __restrict__ bias, float* __restrict__ dx, ThinP p, int Npad) {
    extern __shared__ float4 cs4[];                          // [valid kh][OW][KW * Ci]
    float* cs = reinterpret_cast<float*>(cs4);
    const int row = blockIdx.x, n = row / p.H, h = row - n * p.H;
    const int kwc = p.KW * p.Ci, seg = p.OW * kwc;
    const bool vec = (kwc & 3) == 0 && (Npad & 3) == 0;
    int nv = 0;
    for (int kh = 0; kh < p.KH; ++kh) {
        const int hh = h + p.pad_t - kh;
        if (hh < 0 || hh % S) continue;
        const int oh = hh / S;
        if (oh >= p.OH) continue;
        const float* src = C + ((long long)n * p.OH + oh) * p.OW * Npad + kh * kwc;
        if (vec) {
            const int kq = kwc >> 2;
            float4* dst = reinterpret_cast<float4*>(cs + nv * seg);
            for (int i = threadIdx.x; i < p.OW * kq; i += blockDim.x) {
                const int ow = i / kq, e = i - ow * kq;
                dst[i] = __ldg(reinterpret_cast<const float4*>(src + (long long)ow * Npad) + e);
            }
        } else {
            for (int i = threadIdx.x; i < seg; i += blockDim.x) {
                const int ow = i / kwc, e = i - ow * kwc;
                cs[nv * seg + i] = __ldg(src + (long long)ow * Npad + e);
            }
        }
        ++nv;
    }
    __syncthreads();
    const int Wr = (p.W + 31) & ~31;
    const int parts = (int)blockDim.x >= Wr ? (int)blockDim.x / Wr : 1;
    const int cpt = (p.Ci + parts - 1) / parts;              // channels per thread
    for (int t = threadIdx.x; t < Wr * parts; t += blockDim.x) {
        const int w = t % Wr, c0 = (t / Wr) * cpt;
        const int cnt = min(p.Ci - c0, cpt);
        if (w >= p.W || cnt <= 0) continue;
        float acc[kThinMaxCi];
#pragma unroll
        for (int j = 0; j < kThinMaxCi; ++j) acc[j] = (bias != nullptr && j < cnt) ? __ldg(bias + c0 + j) : 0.f;
        for (int v = 0; v < nv; ++v) {
            for (int kw = 0; kw < p.KW; ++kw) {
                const int ww = w + p.pad_l - kw;
                if (ww < 0 || ww % S) continue;
                const int ow = ww / S;
                if (ow >= p.OW) continue;
                const float* r = cs + v * seg + ow * kwc + kw * p.Ci + c0;
#pragma unroll
                for (int j = 0; j < kThinMaxCi; ++j)
                    if (j < cnt) acc[j] += r[j];
            }
        }
        float* o = dx + ((long long)row * p.W + w) * p.Ci + c0;
#pragma unroll
        for (int j = 0; j < kThinMaxCi; ++j)
            if (j < cnt) o[j] = acc[j];
    }
}

// out[j][co] = j < K ? w[j][co] : 0   (the HWIO filter is already [K][Co])
__global__ void thin_pad_rows_k(const float* __restrict__ w, float* __restrict__ out, int K, int Kpad, int Co) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Kpad * Co) return;
    out[i] = i < K * Co ? __ldg(w + i) : 0.f;
}

// out[co][j] = j < K ? w[j][co] : 0   (1x1 filter [Ci' = Co][Co' = Npad] of the input-gradient product)
__global__ void thin_transpose_w_k(const float* __restrict__ w, float* __restrict__ out, int K, int Npad, int Co) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Npad * Co) return;
    const int co = i / Npad, j = i - co * Npad;
    out[i] = j < K ? __ldg(w + (long long)j * Co + co) : 0.f;
}

// dx[i] = bias[i % C] (or 0): the initial value the scatter epilogue adds to
__global__ void thin_bias_fill_k(float* __restrict__ dx, const float* __restrict__ bias, long long n, int C) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dx[i] = bias != nullptr ? __ldg(bias + (int)(i % C)) : 0.f;
}

__global__ void thin_dw_out_k(const float* __restrict__ dW, float* __restrict__ dw, int n, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dw[i] = accumulate ? dw[i] + dW[i] : dW[i];
}

eg_conv_shape gemm_shape(long long P, int Ci, int Co) {
    eg_conv_shape g{};
    g.N = (int)P; g.H = 1; g.W = 1; g.Ci = Ci; g.OH = 1; g.OW = 1; g.Co = Co; g.KH = 1; g.KW = 1; g.stride = 1;
    g.pad_t = 0; g.pad_l = 0;
    return g;
}

int launch_im2col(const float* x, float* A, const ThinP& p, long long P, cudaStream_t st) {
    const size_t smem = sizeof(float) * (size_t)p.KH * p.W * p.Ci;
    if (smem <= 40 * 1024 && p.KW < 256 && p.KW * p.Ci < 65536) {
        const int rows = p.N * p.OH;
        switch (p.Kpad) {
            case 32: thin_im2col_row_k<32><<<rows, 256, smem, st>>>(x, A, p); break;
            case 64: thin_im2col_row_k<64><<<rows, 256, smem, st>>>(x, A, p); break;
            case 96: thin_im2col_row_k<96><<<rows, 256, smem, st>>>(x, A, p); break;
            default: thin_im2col_row_k<128><<<rows, 256, smem, st>>>(x, A, p); break;
        }
    } else {
        const long long total4 = P * (p.Kpad / 4);
        thin_im2col_k<<<eg_ceil_div(total4, 256), 256, 0, st>>>(x, A, p, total4);
    }
    EG_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// Filter gradient of the thin layers on the CUDA cores: dw[K, Co] += patch(x)[P, K]^T . dy[P, Co], K = KH*KW*Ci <= 160.
// The tensor-core formulations of this product either pay a patch-matrix round trip or leave most TMEM lanes empty
// (conv_tc.cu, gather_wgrad), and the generic FFMA implicit GEMM reaches 21 TFLOP/s on it.  Here one block owns ALL K rows
// of a 64-channel column slab: 32 pixels of dy (8 KB) and of the patch matrix (<= 20 KB) are staged in shared memory,
// and every thread keeps a 4-channel x KC-row register tile (rows tk, tk + 16, ...): per pixel 1 + KC shared-memory
// reads (conflict-free: 16 distinct float4 of dy, 2 distinct patch words per warp) feed 4 * KC FMAs.  Pixels are split
// across blocks; red.global.add.v4 epilogue.
constexpr int kWgPix = 32;

template <int KC>
__global__ void __launch_bounds__(256)
thin_wgrad_ffma_k(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw, ThinP p, int Co,
                  long long P, int chunks_per_block) {
    constexpr int KP = KC * 16;                              // padded K
    __shared__ __align__(16) float sdy[kWgPix][64];
    __shared__ float spatch[kWgPix][KP + 1];
    __shared__ int tab[KP];                                  // k -> kh << 24 | kw << 16 | (kh * W + kw) * Ci + ci ; -1 beyond K
    const int tid = threadIdx.x, tk = tid >> 4, tc = tid & 15;
    const int co0 = blockIdx.x * 64;
    for (int j = tid; j < KP; j += 256) {
        int t = -1;
        if (j < p.K) {
            const int tap = j / p.Ci, ci = j - tap * p.Ci, kh = tap / p.KW, kw = tap - kh * p.KW;
            t = (kh << 24) | (kw << 16) | ((kh * p.W + kw) * p.Ci + ci);
        }
        tab[j] = t;
    }
    float acc[KC][4];
#pragma unroll
    for (int i = 0; i < KC; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
    const long long chunk0 = (long long)blockIdx.y * chunks_per_block;
    for (int ch = 0; ch < chunks_per_block; ++ch) {
        const long long q0 = (chunk0 + ch) * kWgPix;
        if (q0 >= P) break;
        __syncthreads();
        // dy rows: 32 pixels x 64 channels, 16 bytes per thread x 2
        for (int i = tid; i < kWgPix * 16; i += 256) {
            const int pp = i >> 4, c4 = (i & 15) * 4;
            const long long q = q0 + pp;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (q < P) v = __ldg(reinterpret_cast<const float4*>(dy + q * Co + co0 + c4));
            *reinterpret_cast<float4*>(&sdy[pp][c4]) = v;
        }
        // patch rows: thread -> (pixel = tid / 8, columns tid % 8, +8, ...)
        {
            const int pp = tid >> 3;
            const long long q = q0 + pp;
            const bool pv = q < P;
            int ow = 0, oh = 0; long long n = 0;
            if (pv) { ow = (int)(q % p.OW); const long long t2 = q / p.OW; oh = (int)(t2 % p.OH); n = t2 / p.OH; }
            const int h0 = oh * p.stride - p.pad_t, w0 = ow * p.stride - p.pad_l;
            const float* base = x + ((n * p.H + h0) * p.W + w0) * p.Ci;
            for (int j = tid & 7; j < KP; j += 8) {
                const int t = tab[j];
                float v = 0.f;
                if (pv && t >= 0) {
                    const int h = h0 + (t >> 24), w = w0 + ((t >> 16) & 255);
                    if (h >= 0 && h < p.H && w >= 0 && w < p.W) v = __ldg(base + (t & 0xffff));
                }
                spatch[pp][j] = v;
            }
        }
        __syncthreads();
#pragma unroll 4
        for (int pp = 0; pp < kWgPix; ++pp) {
            const float4 d = *reinterpret_cast<const float4*>(&sdy[pp][tc * 4]);
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                const float a = spatch[pp][tk + 16 * i];
                acc[i][0] = fmaf(a, d.x, acc[i][0]); acc[i][1] = fmaf(a, d.y, acc[i][1]);
                acc[i][2] = fmaf(a, d.z, acc[i][2]); acc[i][3] = fmaf(a, d.w, acc[i][3]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KC; ++i) {
        const int k = tk + 16 * i;
        if (k < p.K) {
            float* o = dw + (size_t)k * Co + co0 + tc * 4;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(acc[i][0]), "f"(acc[i][1]), "f"(acc[i][2]), "f"(acc[i][3]) : "memory");
        }
    }
}

bool thin_common(const eg_conv_shape* s) {
    if (s->Ci < 1 || s->Ci > kThinMaxCi) return false;
    if (s->KH * s->KW * s->Ci > kThinMaxK) return false;
    if ((long long)s->N * s->OH * s->OW >= (1ll << 31)) return false;
    return s->stride >= 1;
}

}  // namespace

int eg_thin_supported_fwd(const eg_conv_shape* s) { return thin_common(s) && s->Co % 64 == 0; }
int eg_thin_supported_bwd_data(const eg_conv_shape* s) { return thin_common(s) && s->Co % 32 == 0; }
int eg_thin_supported_bwd_weight(const eg_conv_shape* s) { return thin_common(s) && s->Co % 32 == 0; }

int eg_thin_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int three_x,
                       cudaStream_t st) {
    const ThinP p = make_p(s);
    const long long P = (long long)s->N * s->OH * s->OW;
    float *A = nullptr, *wp = nullptr;
    if (int r = eg_tc_scratch(st, 1, sizeof(float) * (size_t)P * p.Kpad, &A)) return r;
    if (int r = eg_tc_scratch(st, 2, sizeof(float) * (size_t)p.Kpad * s->Co, &wp)) return r;
    if (int r = launch_im2col(x, A, p, P, st)) return r;
    thin_pad_rows_k<<<eg_ceil_div((long long)p.Kpad * s->Co, 256), 256, 0, st>>>(w, wp, p.K, p.Kpad, s->Co);
    EG_CHECK_LAUNCH();
    const eg_conv_shape g = gemm_shape(P, p.Kpad, s->Co);
    return eg_tc_conv2d_fwd(&g, A, wp, bias, y, three_x, st);
}

int eg_thin_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx,
                            int three_x, cudaStream_t st) {
    const ThinP p = make_p(s);
    const long long P = (long long)s->N * s->OH * s->OW;
    const int Npad = p.K <= 64 ? 64 : 128;
    float *C = nullptr;
    // filter of a prepared-filter set: its [Npad][Co] hi/lo copy IS the dense product's prepared operand (no launch here)
    float* wt = eg_tc_managed_thin(w, (size_t)p.K * s->Co, three_x ? 3 : 1, 1);
    if (wt == nullptr) {
        if (int r = eg_tc_scratch(st, 2, sizeof(float) * (size_t)Npad * s->Co, &wt)) return r;
        thin_transpose_w_k<<<eg_ceil_div((long long)Npad * s->Co, 256), 256, 0, st>>>(w, wt, p.K, Npad, s->Co);
        EG_CHECK_LAUNCH();
    }
    const eg_conv_shape g = gemm_shape(P, s->Co, Npad);
    if (eg_tc_scatter_supported(s)) {
        // dx = bias (or 0), then the dense product's epilogue adds every im2col column where it belongs: no [P, Npad]
        // product matrix, no col2im pass
        const long long pixels = (long long)s->N * s->H * s->W;
        thin_bias_fill_k<<<eg_ceil_div(pixels * s->Ci, 256 * 4), 256, 0, st>>>(dx, bias, pixels * s->Ci, s->Ci);
        EG_CHECK_LAUNCH();
        return eg_tc_conv2d_fwd(&g, dy, wt, nullptr, dx, three_x, st, nullptr, s);
    }
    if (int r = eg_tc_scratch(st, 1, sizeof(float) * (size_t)P * Npad, &C)) return r;
    {
        if (int r = eg_tc_conv2d_fwd(&g, dy, wt, nullptr, C, three_x, st)) return r;
    }
    const int threads = 256;
    const size_t smem = sizeof(float) * (size_t)((s->KH + s->stride - 1) / s->stride) * s->OW * s->KW * s->Ci;
    if (s->stride == 1 && smem <= 40 * 1024) thin_col2im_row_k<1><<<s->N * s->H, threads, smem, st>>>(C, bias, dx, p, Npad);
    else if (s->stride == 2 && smem <= 40 * 1024) thin_col2im_row_k<2><<<s->N * s->H, threads, smem, st>>>(C, bias, dx, p, Npad);
    else {
        const long long pixels = (long long)s->N * s->H * s->W;
        thin_col2im_k<<<eg_ceil_div(pixels, 256), 256, 0, st>>>(C, bias, dx, p, Npad, pixels);
    }
    EG_CHECK_LAUNCH();
    return 0;
}

int eg_thin_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate,
                              int three_x, cudaStream_t st) {
    const ThinP p = make_p(s);
    const long long P = (long long)s->N * s->OH * s->OW;
    float *A = nullptr, *dW = nullptr;
    if (int r = eg_tc_scratch(st, 1, sizeof(float) * (size_t)P * p.Kpad, &A)) return r;
    if (int r = eg_tc_scratch(st, 2, sizeof(float) * (size_t)p.Kpad * s->Co, &dW)) return r;
    if (int r = launch_im2col(x, A, p, P, st)) return r;
    const eg_conv_shape g = gemm_shape(P, p.Kpad, s->Co);
    if (int r = eg_tc_conv2d_bwd_weight(&g, A, dy, dW, 0, three_x, st)) return r;
    const int n = p.K * s->Co;
    thin_dw_out_k<<<eg_ceil_div(n, 256), 256, 0, st>>>(dW, dw, n, accumulate);
    EG_CHECK_LAUNCH();
    return 0;
}

// FFMA filter gradient of a thin layer (see thin_wgrad_ffma_k); returns -100 when the shape is not covered
int eg_thin_wgrad_ffma(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int sms,
                       cudaStream_t st) {
    const ThinP p = make_p(s);
    if (s->Ci > 8 || p.K > 160 || s->Co % 64 || s->KW > 255 || ((s->KH - 1) * s->W + s->KW) * s->Ci >= 65536) return -100;
    if ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dw)) & 15) return -100;
    const long long P = (long long)s->N * s->OH * s->OW;
    const long long chunks = (P + kWgPix - 1) / kWgPix;
    const int slabs = s->Co / 64;
    long long blocks_y = (8ll * sms + slabs - 1) / slabs;                 // ~8 blocks per SM in flight (30 KB smem, 256 threads)
    if (blocks_y > chunks) blocks_y = chunks;
    if (blocks_y > 65535) blocks_y = 65535;
    const int cpb = (int)((chunks + blocks_y - 1) / blocks_y);
    blocks_y = (chunks + cpb - 1) / cpb;
    if (!accumulate) {
        cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)p.K * s->Co, st);
        if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    }
    const dim3 grid(slabs, (unsigned)blocks_y);
    const int KC = (p.K + 15) / 16;
#define EG_TW(kc) thin_wgrad_ffma_k<kc><<<grid, 256, 0, st>>>(x, dy, dw, p, s->Co, P, cpb)
    switch (KC) {
        case 1: EG_TW(1); break;
        case 2: EG_TW(2); break;
        case 3: EG_TW(3); break;
        case 4: EG_TW(4); break;
        case 5: EG_TW(5); break;
        case 6: case 7: EG_TW(7); break;
        case 8: case 9: case 10: EG_TW(10); break;
        default: return -100;
    }
#undef EG_TW
    EG_CHECK_LAUNCH();
    return 0;
}
