// C-ABI plumbing: error text, device info and the conv algorithm dispatch.
#include "common.cuh"

#include <stdio.h>
#include <string.h>

// conv_simt.cu
int eg_conv_shape_check(const eg_conv_shape* s);
int eg_simt_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, const EgEpi* epi, cudaStream_t st);
int eg_simt_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx, const EgEpi* epi, cudaStream_t st);
int eg_simt_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int sm_count, cudaStream_t st);
// conv_tc.cu
int eg_tc_supported_fwd(const eg_conv_shape* s);
int eg_tc_supported_bwd_data(const eg_conv_shape* s);
int eg_tc_supported_bwd_weight(const eg_conv_shape* s);
int eg_tc_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int three_x, cudaStream_t st, const EgEpi* epi, const eg_conv_shape* scatter = nullptr);
int eg_tc_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx, int three_x, cudaStream_t st, const EgEpi* epi);
int eg_tc_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int three_x, cudaStream_t st);

unsigned long long g_eg_kernel_launches = 0;
static thread_local char g_err[512] = "";

int eg_fail(cudaError_t e, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
    return -1;
}
int eg_fail_arg(const char* what, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "invalid argument: !(%s) at %s:%d", what, file, line);
    return -2;
}

static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sm_count = n;
        else
            g_sm_count = 148;
    }
    return g_sm_count;
}

static int g_default_algo = EG_ALGO_AUTO;

extern "C" {

const char* eg_last_error(void) { return g_err; }
int eg_abi_version(void) { return EG_ABI_VERSION; }

int eg_device_info(int* sms, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return eg_fail(e, __FILE__, __LINE__);
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return 0;
}

/* what EG_ALGO_AUTO resolves to for layers the tensor-core path supports (EG_ALGO_TC or EG_ALGO_TC3X);
 * EG_ALGO_SIMT forces the fp32 path everywhere. */
int eg_set_default_algo(int algo) {
    EG_REQUIRE(algo >= EG_ALGO_AUTO && algo <= EG_ALGO_TC3X);
    g_default_algo = algo;
    return 0;
}
int eg_get_default_algo(void) { return g_default_algo; }
long long eg_kernel_launches(void) { return (long long)g_eg_kernel_launches; }

static int resolve(int algo, int supported) {
    if (algo == EG_ALGO_AUTO) algo = g_default_algo;
    if (algo == EG_ALGO_AUTO) algo = EG_ALGO_TC3X;    // parity first: fp32-class accuracy on the tensor cores
    if (algo != EG_ALGO_SIMT && !supported) algo = EG_ALGO_SIMT;
    return algo;
}

int eg_conv2d_algo_for(const eg_conv_shape* s, int pass, int algo) {
    if (eg_conv_shape_check(s)) return -2;
    int sup = pass == 0 ? eg_tc_supported_fwd(s) : (pass == 1 ? eg_tc_supported_bwd_data(s) : eg_tc_supported_bwd_weight(s));
    return resolve(algo, sup);
}

// the epilogue as separate in-place passes, for the routes that have no fused one (rc == 1 from the implementation)
static int run_epilogue(const EgEpi& e, float* out, long long n, void* stream) {
    if (e.mode == EG_EPI_ACT) return eg_act_fwd(out, out, n, e.act, stream);
    if (e.mode == EG_EPI_MASK) return eg_act_bwd(e.mask, out, out, n, e.act, stream);
    return 0;
}

int eg_conv2d_fwd_ex(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int epi,
                     int act, const float* mask_src, int algo, void* stream) {
    if (int r = eg_conv_shape_check(s)) return r;
    EG_REQUIRE(x && w && y);
    EG_REQUIRE(epi >= EG_EPI_NONE && epi <= EG_EPI_MASK && (epi != EG_EPI_MASK || mask_src));
    const EgEpi e{epi, act, mask_src};
    const int a = resolve(algo, eg_tc_supported_fwd(s));
    int rc = a == EG_ALGO_SIMT ? eg_simt_conv2d_fwd(s, x, w, bias, y, &e, (cudaStream_t)stream)
                               : eg_tc_conv2d_fwd(s, x, w, bias, y, a == EG_ALGO_TC3X, (cudaStream_t)stream, &e);
    if (rc == 1) rc = run_epilogue(e, y, (long long)s->N * s->OH * s->OW * s->Co, stream);
    return rc;
}

int eg_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int algo,
                  void* stream) {
    return eg_conv2d_fwd_ex(s, x, w, bias, y, EG_EPI_NONE, 0, nullptr, algo, stream);
}

int eg_conv2d_bwd_data_ex(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx, int epi,
                          int act, const float* mask_src, int algo, void* stream) {
    if (int r = eg_conv_shape_check(s)) return r;
    EG_REQUIRE(dy && w && dx);
    EG_REQUIRE(epi >= EG_EPI_NONE && epi <= EG_EPI_MASK && (epi != EG_EPI_MASK || mask_src));
    const EgEpi e{epi, act, mask_src};
    const int a = resolve(algo, eg_tc_supported_bwd_data(s));
    int rc = a == EG_ALGO_SIMT ? eg_simt_conv2d_bwd_data(s, dy, w, bias, dx, &e, (cudaStream_t)stream)
                               : eg_tc_conv2d_bwd_data(s, dy, w, bias, dx, a == EG_ALGO_TC3X, (cudaStream_t)stream, &e);
    if (rc == 1) rc = run_epilogue(e, dx, (long long)s->N * s->H * s->W * s->Ci, stream);
    return rc;
}

int eg_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx,
                       int algo, void* stream) {
    return eg_conv2d_bwd_data_ex(s, dy, w, bias, dx, EG_EPI_NONE, 0, nullptr, algo, stream);
}

int eg_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate,
                         int algo, void* stream) {
    if (int r = eg_conv_shape_check(s)) return r;
    EG_REQUIRE(x && dy && dw);
    const int a = resolve(algo, eg_tc_supported_bwd_weight(s));
    if (a == EG_ALGO_SIMT)
        return eg_simt_conv2d_bwd_weight(s, x, dy, dw, accumulate, sm_count(), (cudaStream_t)stream);
    return eg_tc_conv2d_bwd_weight(s, x, dy, dw, accumulate, a == EG_ALGO_TC3X, (cudaStream_t)stream);
}

}  // extern "C"
