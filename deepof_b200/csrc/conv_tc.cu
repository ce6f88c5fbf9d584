// tcgen05 (kind::tf32) implicit-GEMM convolution family -- placeholder until the kernels land.
#include "common.cuh"
namespace dofb {
int tc_conv_fwd(const dofb_conv_geom *, const float *, int, const float *, const float *, float *, int, int, cudaStream_t) {
    return set_error("dofb_conv_fwd: math=TF32 not available in this build");
}
int tc_conv_dgrad(const dofb_conv_geom *, const float *, int, const float *, const float *, float *, int, int, int, cudaStream_t) {
    return set_error("dofb_conv_dgrad: math=TF32 not available in this build");
}
int tc_conv_wgrad(const dofb_conv_geom *, const float *, int, const float *, int, float *, cudaStream_t) {
    return set_error("dofb_conv_wgrad: math=TF32 not available in this build");
}
}  // namespace dofb
