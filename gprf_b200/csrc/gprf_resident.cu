// One translation unit per covariance family (-DRES_DFN=.. -DRES_WFN=..) holding the resident
// (shared-memory) unit kernel of resident.cuh, so that the four instantiations compile in parallel
// with the main translation unit (see build.sh).
#include <cuda_runtime.h>
#define GPRF_RES_KERNEL_ONLY   // the plan / combine kernels are defined in gprf_lib.cu
#include "resident.cuh"

#ifndef RES_DFN
#error "compile with -DRES_DFN=0|1 -DRES_WFN=0|1"
#endif

namespace gprf {
namespace res {

template <int DFN, int WFN> int resident_set_attr();
template <int DFN, int WFN> void resident_launch(const ResParams& P, int grid, cudaStream_t st);

template <>
int resident_set_attr<RES_DFN, RES_WFN>() {
  return (int)cudaFuncSetAttribute(k_resident<RES_DFN, RES_WFN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   R_SMEM_BYTES);
}

template <>
void resident_launch<RES_DFN, RES_WFN>(const ResParams& P, int grid, cudaStream_t st) {
  k_resident<RES_DFN, RES_WFN><<<grid, RNT, R_SMEM_BYTES, st>>>(P);
}

}  // namespace res
}  // namespace gprf
