// placeholder until the tcgen05 kernels land
#include "common.cuh"
int eg_tc_supported_fwd(const eg_conv_shape*) { return 0; }
int eg_tc_supported_bwd_data(const eg_conv_shape*) { return 0; }
int eg_tc_supported_bwd_weight(const eg_conv_shape*) { return 0; }
int eg_tc_conv2d_fwd(const eg_conv_shape*, const float*, const float*, const float*, float*, int, cudaStream_t) { return -3; }
int eg_tc_conv2d_bwd_data(const eg_conv_shape*, const float*, const float*, const float*, float*, int, cudaStream_t) { return -3; }
int eg_tc_conv2d_bwd_weight(const eg_conv_shape*, const float*, const float*, float*, int, int, cudaStream_t) { return -3; }
