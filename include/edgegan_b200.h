/* edgegan_b200 -- C ABI of the B200 (sm_100a) EdgeGAN hot path.
 *
 * The reference (sysu-imsl/EdgeGAN) has no FFI: its arithmetic is TensorFlow-1.14 library calls made by the
 * graph-building functions in edgegan/nn/modules/{conv,linear,normalization,activation}.py and edgegan/nn/functional.py.  Every entry point below
 * replaces one of those TF call sites (cited as reference file:line, relative to /root/reference/edgegan/),
 * plus the backward / double-backward passes TF derived implicitly through tf.gradients / minimize().
 *
 * Conventions
 *   - all tensors are float32, NHWC, dense, device pointers owned by the caller (torch CUDA tensors);
 *   - `stream` is a cudaStream_t passed as void*; every call is stream-ordered and asynchronous;
 *   - return 0 on success, a negative value on error (eg_last_error() gives the text); nothing throws;
 *   - no call allocates device memory; calls that need scratch take it as an explicit argument.
 */
#ifndef EDGEGAN_B200_H
#define EDGEGAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define EG_ABI_VERSION 1

/* activation codes (nn/modules/activation.py:4-15; generator.py:74) */
#define EG_ACT_NONE 0
#define EG_ACT_RELU 1
#define EG_ACT_LRELU_BLOCK 2 /* tf.maximum(x, 0.2x): slope 1 at x == 0 (activation.py:9) */
#define EG_ACT_TANH 3
#define EG_ACT_SIGMOID 4     /* discriminator.py:81 (the `prob` output) */
#define EG_ACT_LRELU 5       /* nn.lrelu = tf.maximum(0.2x, x): slope 0.2 at x == 0 (activation.py:30-32) */

/* conv algorithm selector */
#define EG_ALGO_AUTO 0
#define EG_ALGO_SIMT 1   /* fp32 FFMA implicit GEMM (exact fp32; thin-channel layers, on-GPU checker) */
#define EG_ALGO_TC 2     /* tcgen05 kind::tf32, operands rounded to TF32 by the tensor core            */
#define EG_ALGO_TC3X 3   /* tcgen05 3xTF32 split (hi*hi + hi*lo + lo*hi), fp32-class accuracy          */

const char* eg_last_error(void);
int eg_abi_version(void);
int eg_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* What EG_ALGO_AUTO resolves to for layers the tensor-core kernels support (default EG_ALGO_TC3X);
 * layers they do not support (channel counts that are not multiples of 32 / 64; the forward and filter-gradient
 * passes of the 3-channel image layers) take the fp32 SIMT kernel.  The input gradient of the 3-channel layers runs
 * on the tensor cores as a dense product plus col2im (csrc/conv_thin.cu). */
int eg_set_default_algo(int algo);
int eg_get_default_algo(void);

/* development knobs (kernel layout / route experiments from tools/); not part of the stable surface */
int eg_debug_set(int key, int value);
int eg_norm_debug(int value);

/* Prepared-filter sets.  The tensor-core conv kernels read a re-laid-out copy of the filter (forward: [tap][Co][Ci],
 * input gradient: [tap][Ci][Co]; in the 3xTF32 mode each with its low-order split next to it).  A set keeps those
 * copies for ALL conv filters of one network in library-owned device memory and refreshes them with ONE kernel launch:
 * the owner of the weights calls eg_filter_set_prepare on the stream right after every write to them (the reference's
 * counterpart of that moment is the end of an RMSPropOptimizer.minimize run, edgegan/models/edgegan.py:105-124, or a
 * Saver.restore, :641-657).  eg_conv2d_fwd / eg_conv2d_bwd_data look the filter pointer up and launch no preparation
 * kernel; freshness follows from stream order, also when the launches are replayed from a CUDA graph.  Filters that
 * are in no set are prepared per call.  `algo`: EG_ALGO_TC or EG_ALGO_TC3X (the copies are mode specific; a conv call
 * in another mode ignores the set).  A pointer belongs to at most one set (the newest); destroy a set before the
 * memory of its filters is released. */
typedef struct { const float* w; int taps, Ci, Co; } eg_filter_desc;    /* filter [taps][Ci][Co] (HWIO, taps = KH*KW) */
int eg_filter_set_create(const eg_filter_desc* descs, int n, int algo, long long* handle);
int eg_filter_set_prepare(long long handle, cudaStream_t stream);
int eg_filter_set_destroy(long long handle);
long long eg_filter_set_hits(void);     /* conv calls served from a set so far */

/* number of CUDA kernels the library has launched so far in this process (host-side counter, one host thread per
 * rank; launches recorded into a CUDA graph count once, at capture) -- bench.py's gpu_launches */
long long eg_kernel_launches(void);

/* CRC-32C (Castagnoli) of `n` bytes of HOST memory continuing from `crc` (0 to start): the checksum of the TensorFlow
 * tensor-bundle files tf.train.Saver writes (edgegan/models/edgegan.py:421,547,635-657); returns the checksum */
unsigned int eg_crc32c(const void* data, long long n, unsigned int crc);

/* A strided convolution y[N,OH,OW,Co] = conv(x[N,H,W,Ci], w[KH,KW,Ci,Co]) with zero padding pad_t/pad_l before
 * the first row/column (whatever is needed after the last one is implied by OH/OW):
 *   y[n,oh,ow,co] = sum_{r,q,ci} x[n, oh*stride - pad_t + r, ow*stride - pad_l + q, ci] * w[r,q,ci,co]
 * This one descriptor serves tf.nn.conv2d SAME/VALID (conv.py:26,29,279) and, with the roles of
 * x and y swapped, tf.nn.conv2d_transpose (conv.py:49; SURVEY A2: SAME s2 k5 <=> pad_t = pad_l = 1). */
typedef struct {
    int N, H, W, Ci;
    int OH, OW, Co;
    int KH, KW, stride, pad_t, pad_l;
} eg_conv_shape;

/* the algorithm a call with (shape, pass, algo) will actually run: pass 0 fwd, 1 bwd_data, 2 bwd_weight */
int eg_conv2d_algo_for(const eg_conv_shape* s, int pass, int algo);
/* replaces tf.nn.conv2d (+ tf.nn.bias_add) -- conv.py:29,34 ; bias[Co] may be NULL */
int eg_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int algo,
                  void* stream);
/* input gradient of the conv == tf.nn.conv2d_transpose forward (conv.py:49-53); bias[Ci] may be NULL */
int eg_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx,
                       int algo, void* stream);
/* The same two with a fused epilogue, for a conv_block without a norm (conv.py:61-67 with norm=None: conv ->
 * activation_fn) and its backward, which tf.gradients emits as separate elementwise kernels:
 *   EG_EPI_ACT   out = act(conv + bias)
 *   EG_EPI_MASK  out = (conv + bias) * act'(mask_src[i])   mask_src shaped like out; for relu / lrelu the derivative
 *                depends only on the sign, so the post-activation tensor serves as mask_src
 * (EG_EPI_NONE = the plain calls above). */
#define EG_EPI_NONE 0
#define EG_EPI_ACT 1
#define EG_EPI_MASK 2
int eg_conv2d_fwd_ex(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int epi,
                     int act, const float* mask_src, int algo, void* stream);
int eg_conv2d_bwd_data_ex(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx, int epi,
                          int act, const float* mask_src, int algo, void* stream);
/* filter gradient dw[KH,KW,Ci,Co] (= or +=) sum_pixels x (x) dy */
int eg_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate,
                         int algo, void* stream);
/* column sums: db[c] (= or +=) sum_r dy[r, c]   (gradient of tf.nn.bias_add, conv.py:34,53; linear.py:29) */
int eg_bias_grad(const float* dy, long long rows, int C, float* db, int accumulate, void* stream);

/* instance norm (normalization.py:13-18) fused with the activation that follows it in
 * conv_block / deconv_block / residual (conv.py:61-85,124-130).  x,y: [N,P,C] with P = H*W.
 * stats[N,C,2] = (mean, sqrt(biased var)).  n = (x-mean)/(sd+eps); y = act(n). */
int eg_instnorm_fwd(const float* x, float* y, float* stats, int N, int P, int C, float eps, int act, void* stream);
/* gx = J^T (gy * act'(n)) [+ addend];  addend may be NULL */
int eg_instnorm_bwd(const float* x, const float* stats, const float* gy, const float* addend, float* gx, int N,
                    int P, int C, float eps, int act, void* stream);
/* second-order pass used by the WGAN-GP penalty (edgegan.py:38-42, functional.py:26-29).  With
 * gx = F(x, gy) the first-order backward above and `t` the cotangent arriving on gx:
 *   out_gy = act'(n) * J t          (cotangent on gy)
 *   out_x  = d<t, F(x, gy)>/dx      (cotangent on x through mean / sd; act' treated as locally constant) */
int eg_instnorm_bwd2(const float* x, const float* stats, const float* gy, const float* t, float* out_gy,
                     float* out_x, int N, int P, int C, float eps, int act, void* stream);

/* plain activations (activation.py:4-15, generator.py:74) */
int eg_act_fwd(const float* x, float* y, long long n, int act, void* stream);
int eg_act_bwd(const float* x_pre, const float* gy, float* gx, long long n, int act, void* stream);

/* batch norm of the generator's h0 (normalization.py:19-25 via generator.py:51-52), split so
 * the [2C] sums can be all-reduced between ranks (sync-BN): x [R,C].
 *   sums = (sum x, sum x^2) per channel                      -> eg_bn_stats
 *   y = act(gamma * (x-mu)/sqrt(var+eps) + beta)              -> eg_bn_apply   (count = global rows)
 *   red = (sum gpre, sum gpre*xhat) with gpre = gy*act'(pre)  -> eg_bn_bwd_reduce
 *   gx = gamma*rstd*(gpre - red0/count - xhat*red1/count)     -> eg_bn_bwd_apply  */
int eg_bn_stats(const float* x, float* sums, int R, int C, void* stream);
int eg_bn_apply(const float* x, const float* sums, float count, const float* gamma, const float* beta, float* y,
                int R, int C, float eps, int act, void* stream);
int eg_bn_bwd_reduce(const float* x, const float* sums, float count, const float* gamma, const float* beta,
                     const float* gy, float* red, int R, int C, float eps, int act, void* stream);
int eg_bn_bwd_apply(const float* x, const float* sums, float count, const float* gamma, const float* beta,
                    const float* gy, const float* red, float* gx, int R, int C, float eps, int act, void* stream);

/* discriminator head d = h @ Matrix + bias with Matrix [F,1] (linear.py:29-31 via discriminator.py:76) */
int eg_rowdot_fwd(const float* h, const float* w, const float* bias, float* d, int B, int F, void* stream);
int eg_rowdot_bwd_input(const float* gd, const float* w, float* gh, int B, int F, void* stream);
int eg_rowdot_bwd_weight(const float* gd, const float* h, float* gw, float* gb, int B, int F, int accumulate,
                         void* stream);

/* tf.image.resize_images(method=BICUBIC) 2x, legacy kernel (edgegan.py:211-213) */
int eg_bicubic_up2_fwd(const float* x, float* y, int N, int H, int W, int C, void* stream);
int eg_bicubic_up2_bwd(const float* gy, float* gx, int N, int H, int W, int C, void* stream);

/* strided 2-D copy: tf.concat(axis=2) / width slices (edgegan.py:203-209,243-247) */
int eg_copy2d(const float* src, long long src_stride, float* dst, long long dst_stride, long long rows,
              long long cols, void* stream);
int eg_fill(float* dst, long long n, float value, void* stream);
/* dst[i] = lut[src[i]] for n bytes (lut: 256 floats in device memory): image bytes -> [-1, 1] on the device with the
 * table of utils.transform (edgegan/utils/utils.py:160), so the loader uploads 1 byte per value */
int eg_u8_lut_f32(const void* src, const float* lut, float* dst, long long n, void* stream);
/* y = a*x + b*y */
int eg_axpby(const float* x, float* y, long long n, float a, float b, void* stream);

/* WGAN-GP (edgegan.py:32-42, functional.py:26-29) */
/* xhat = real + alpha[b] * (fake - real) */
int eg_gp_interpolate(const float* real, const float* fake, const float* alpha, float* xhat, int B, long long per,
                      void* stream);
/* seed of the first-order backward: dd[b] = d/dd (sigmoid(d) + d) = 1 + s(1-s) */
int eg_gp_seed(const float* d, float* dd, int B, void* stream);
/* norms[b] = ||g_b||; gbar = weight * 2 (norm-1) / (Bglobal * norm) * g; loss[0] += weight * sum_b (norm-1)^2 / Bglobal */
int eg_gp_penalty(const float* g, float* gbar, float* norms, float* loss, int B, long long per, float weight,
                  float inv_global_batch, void* stream);
/* dbar[b] = ddbar[b] * s(1-s)(1-2s)  (cotangent on the logit through the seed) */
int eg_gp_seed_bwd(const float* d, const float* ddbar, float* dbar, int B, void* stream);

/* out[0] (= or +=) scale * sum(x)  -- the reduce_mean's of functional.py:32-41 */
int eg_sum_scaled(const float* x, long long n, float scale, float* out, int accumulate, void* stream);

/* encoder pieces (encoder.py:54-84, conv.py:25,70-85) */
int eg_reflect_pad_fwd(const float* x, float* y, int N, int H, int W, int C, int p, void* stream);
int eg_reflect_pad_bwd(const float* gy, float* gx, int N, int H, int W, int C, int p, void* stream);
/* y = avgpool2x2(relu(a + b))   (conv.py:85 + encoder.py:68); b may be NULL */
int eg_addrelu_pool2_fwd(const float* a, const float* b, float* y, int N, int H, int W, int C, void* stream);
/* g = relu'(a + b) * gy / 4 (same gradient for a and b) */
int eg_addrelu_pool2_bwd(const float* a, const float* b, const float* gy, float* g, int N, int H, int W, int C,
                         void* stream);
/* y[n,c] = mean_p relu(x[n,p,c])   (encoder.py:69-71: relu, 8x8 SAME avg pool over a <=8x8 map, flatten) */
int eg_relu_globalmean_fwd(const float* x, float* y, int N, int P, int C, void* stream);
int eg_relu_globalmean_bwd(const float* x, const float* gy, float* gx, int N, int P, int C, void* stream);
/* z = mu + eps * exp(log_sigma)   (encoder.py:78-82; eps is ONE scalar for the whole batch).
 * eps_dev, when not NULL, is a device scalar that overrides `eps` (keeps a captured CUDA graph re-playable) */
int eg_reparam_fwd(const float* mu, const float* ls, float eps, const float* eps_dev, float* z, long long n,
                   void* stream);
/* loss[0] += weight * mean|target - z| ; gmu = dloss/dmu ; gls = dloss/dls   (functional.py:40-41, edgegan.py:337-342)
 * target has row stride `target_stride` (z carries a trailing class-id column in multi-class mode) */
int eg_zl1_loss_bwd(const float* mu, const float* ls, float eps, const float* eps_dev, const float* target,
                    int target_stride, int B, int Z, float weight, float inv_global_count, float* gmu, float* gls,
                    float* loss, void* stream);

/* ---- multi-class classifier D2 (models/classifier.py:12-119, nn/modules/conv.py:133-357) ---------------- */
/* prelu: tf.maximum(leak*x, x) with a learned scalar leak held in device memory (activation.py:23-27).
 * bwd: gx = gy * (leak*x >= x ? leak : 1) (gx may be NULL); gleak[0] (= or +=) sum gy*x*[leak*x >= x] (may be NULL) */
int eg_prelu_fwd(const float* x, const float* leak, float* y, long long n, void* stream);
int eg_prelu_bwd(const float* x, const float* leak, const float* gy, float* gx, float* gleak, long long n,
                 int accumulate_leak, void* stream);
/* y = prelu(x; leak), y2 = prelu(y; leak2): classifier.py:43-45 (h0) followed by the first MRU unit's
 * norm_activation_in (conv.py:160) in one pass.  _ex: gx (= or +=) ... when accumulate_gx (cotangents that meet at ht) */
int eg_prelu_fwd2(const float* x, const float* leak, float* y, const float* leak2, float* y2, long long n, void* stream);
int eg_prelu_bwd_ex(const float* x, const float* leak, const float* gy, float* gx, float* gleak, long long n,
                    int accumulate_leak, int accumulate_gx, void* stream);
/* The element-wise middle of one MRU unit (nn/modules/conv.py:189-209) as ONE kernel per direction, tensors [N,P,C]:
 *   fwd:  rgl = lrelu(cg + cg_i) written over cg (cg = conv over prelu(ht) + bias, cg_i = conv over the image part of
 *         the concat, conv.py:189-196);  rg = (rgl - min_P)/(max_P - min_P) (:197-198);  plus = ht + rg*img (:209);
 *         hin = prelu(plus; leak) (:212).  stats[N,C,4] = (min, max, #argmin, #argmax).  rg is not stored.
 *   bwd:  from g_hin: g_ht += g_plus, g_img = g_plus*rg, g_cg = lrelu'(rgl) * minmax_bwd(g_plus*img); gleak (may be
 *         NULL) (= or +=) the prelu leak gradient.  C % 4 == 0. */
int eg_mru_gate_fwd(float* cg_rgl, const float* cg_i, const float* ht, const float* img, const float* leak, float* stats,
                    float* plus, float* hin, int N, int P, int C, void* stream);
int eg_mru_gate_bwd(const float* plus, const float* g_hin, const float* img, const float* rgl, const float* stats,
                    const float* leak, float* g_ht, float* g_img, float* g_cg, float* gleak, int accumulate_leak,
                    int N, int P, int C, void* stream);
/* update-gate normalisation (conv.py:197-198): y = (x - min)/(max - min) over H,W per (n,c); stats[N,C,2] =
 * (min, max); bwd sends the min / max terms to the arg-extrema, split evenly between ties */
int eg_minmax_fwd(const float* x, float* y, float* stats, int N, int P, int C, void* stream);
int eg_minmax_bwd(const float* x, const float* stats, const float* gy, float* gx, int N, int P, int C, void* stream);
/* out = a + b*c  (ht + rg*img_new, conv.py:209) ; out = a*b */
int eg_fma3(const float* a, const float* b, const float* c, float* out, long long n, void* stream);
int eg_mul(const float* a, const float* b, float* out, long long n, void* stream);
/* y = mean_pool2x2(a + b) (b may be NULL; pooling.py:4-8 after conv.py:236); gx (= or +=) pooled-gradient ;
 * y[n,c] = mean_p x[n,p,c] (classifier.py:111) and its gradient */
int eg_add_pool2_fwd(const float* a, const float* b, float* y, int N, int H, int W, int C, void* stream);
/* same, and y_act = prelu(y; leak) (the next unit's norm_activation_in, conv.py:160) when y_act != NULL */
int eg_add_pool2_prelu_fwd(const float* a, const float* b, float* y, const float* leak, float* y_act, int N, int H, int W,
                           int C, void* stream);
int eg_pool2_bwd(const float* gy, float* gx, int N, int H, int W, int C, int accumulate, void* stream);
int eg_globalmean_fwd(const float* x, float* y, int N, int P, int C, void* stream);
int eg_globalmean_bwd(const float* gy, float* gx, int N, int P, int C, void* stream);
/* spectral normalisation with ONE power iteration from a frozen u (normalization.py:38-76):
 * W viewed as [K, C]; v = l2n(u W^T), u' = l2n(v W), sigma = v W u'^T, Wbar = W / sigma.  `ws` holds
 * eg_spectral_norm_ws_floats(K, C) floats and carries the forward's intermediates to the backward, which maps
 * Gbar = dL/dWbar to gW = dL/dW THROUGH sigma, v and u' (no stop_gradient in the reference). */
int eg_spectral_norm_ws_floats(int K, int C);
int eg_spectral_norm_fwd(const float* W, const float* u, float* Wbar, float* ws, int K, int C, void* stream);
int eg_spectral_norm_bwd(const float* W, const float* u, float* ws, const float* Gbar, float* gW, int K, int C,
                         void* stream);
/* The same for every weight tensor of a network at once (the classifier normalises 22 filters per step,
 * classifier.py:12-119 through conv.py's spectral_normed_weight): a device-side table of descriptors, forward = 3
 * launches, backward = 1 memset + 3 launches, instead of 4 nodes per tensor and direction.  G / gW are read / written by
 * the backward only.  A tensor [taps, cin, C] may be split along cin at `hd` (all of Wa, Wi set): the forward also writes
 * Wbar as the two contiguous filters Wa [taps, hd, C] and Wi [taps, cin-hd, C]; with Ga / Gi set the backward reads
 * dL/dWbar from two such parts instead of G.  Pointers must stay valid for the life of the set. */
typedef struct {
    const float* W; const float* u; float* Wbar; float* ws;
    const float* G; float* gW;
    float* Wa; float* Wi;
    const float* Ga; const float* Gi;
    int K, C, cin, hd;
} eg_sn_desc;
int eg_spectral_norm_set_create(const eg_sn_desc* descs, int n, long long* handle);
int eg_spectral_norm_set_fwd(long long handle, void* stream);
int eg_spectral_norm_set_bwd(long long handle, void* stream);
int eg_spectral_norm_set_destroy(long long handle);
/* softmax cross-entropy heads (functional.py:5-16): labels = (int) z[b, label_col];
 * focal = 0: loss += weight * mean CE ; focal = 1: loss += weight * mean (1-p_y)^2 CE ; glogits = dloss/dlogits */
int eg_softmax_ce_bwd(const float* logits, const float* z, int z_stride, int label_col, int B, int C, int focal,
                      float weight, float inv_global_batch, float* glogits, float* loss, void* stream);

/* generator input of the multi-class model (edgegan.py:188-197): out[n, zdim+classes] =
 * z[:, :zdim] ++ one_hot(int(z[:, zdim]), classes);  z has zdim+1 columns */
int eg_onehot_concat(const float* z, int n, int zdim, int classes, float* out, void* stream);

/* tf.train.RMSPropOptimizer step on a flat parameter buffer (edgegan.py:105,117-120):
 *   ms = decay*ms + (1-decay)*g^2 ; var -= lr * g / sqrt(ms + eps) */
int eg_rmsprop(float* var, const float* grad, float* ms, long long n, float lr, float decay, float eps,
               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EDGEGAN_B200_H */
