// C-ABI plumbing: error text, device info and the conv algorithm dispatch.
#include "common.cuh"

#include <stdio.h>
#include <string.h>

// conv_simt.cu
int eg_conv_shape_check(const eg_conv_shape* s);
int eg_simt_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, cudaStream_t st);
int eg_simt_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx, cudaStream_t st);
int eg_simt_conv2d_bwd_weight(const eg_conv_shape* s, const float* x, const float* dy, float* dw, int accumulate, int sm_count, cudaStream_t st);
// conv_tc.cu
int eg_tc_supported_fwd(const eg_conv_shape* s);
int eg_tc_supported_bwd_data(const eg_conv_shape* s);
int eg_tc_supported_bwd_weight(const eg_conv_shape* s);
int eg_tc_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int three_x, cudaStream_t st);
int eg_tc_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx, int three_x, cudaStream_t st);
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

int eg_conv2d_fwd(const eg_conv_shape* s, const float* x, const float* w, const float* bias, float* y, int algo,
                  void* stream) {
    if (int r = eg_conv_shape_check(s)) return r;
    EG_REQUIRE(x && w && y);
    const int a = resolve(algo, eg_tc_supported_fwd(s));
    if (a == EG_ALGO_SIMT) return eg_simt_conv2d_fwd(s, x, w, bias, y, (cudaStream_t)stream);
    return eg_tc_conv2d_fwd(s, x, w, bias, y, a == EG_ALGO_TC3X, (cudaStream_t)stream);
}

int eg_conv2d_bwd_data(const eg_conv_shape* s, const float* dy, const float* w, const float* bias, float* dx,
                       int algo, void* stream) {
    if (int r = eg_conv_shape_check(s)) return r;
    EG_REQUIRE(dy && w && dx);
    const int a = resolve(algo, eg_tc_supported_bwd_data(s));
    if (a == EG_ALGO_SIMT) return eg_simt_conv2d_bwd_data(s, dy, w, bias, dx, (cudaStream_t)stream);
    return eg_tc_conv2d_bwd_data(s, dy, w, bias, dx, a == EG_ALGO_TC3X, (cudaStream_t)stream);
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
