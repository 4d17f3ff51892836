// One translation unit per covariance family (-DFUSED_DFN=.. -DFUSED_WFN=..) holding the fused
// per-unit kernel, so that the four instantiations compile in parallel (see build.sh).
#include <cuda_runtime.h>
#define GPRF_FUSED_ONLY   // skip the non-template __global__ wrappers (defined in gprf_lib.cu)
#include "gprf_kernels.cuh"

#ifndef FUSED_DFN
#error "compile with -DFUSED_DFN=0|1 -DFUSED_WFN=0|1"
#endif

namespace gprf {

template <int DFN, int WFN> void fused_set_attr();
template <int DFN, int WFN>
void fused_launch(const EvalParams& P, double* ll_u, double* gth_u, int want_grad, int nunits, cudaStream_t st);

template <>
void fused_set_attr<FUSED_DFN, FUSED_WFN>() {
  cudaFuncSetAttribute(k_unit_fused<FUSED_DFN, FUSED_WFN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                       (int)FUSED_SMEM_BYTES);
}

template <>
void fused_launch<FUSED_DFN, FUSED_WFN>(const EvalParams& P, double* ll_u, double* gth_u, int want_grad,
                                        int nunits, cudaStream_t st) {
  k_unit_fused<FUSED_DFN, FUSED_WFN><<<nunits, NTHREADS, FUSED_SMEM_BYTES, st>>>(P, ll_u, gth_u, want_grad);
}

}  // namespace gprf
