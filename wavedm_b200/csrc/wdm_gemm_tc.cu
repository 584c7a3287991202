// wdm_gemm_tc.cu -- tcgen05 / TMA implicit-GEMM (bf16 in, fp32 accumulate in TMEM). Placeholder until the
// kernel lands: reports every shape as unsupported so the executor uses the CUDA-core path.
#include "wdm_common.cuh"
#include "wdm_engine.h"

namespace wdm {
bool gemm_tc_supported(const GemmParams&) { return false; }
int launch_gemm_tc(const GemmParams&, cudaStream_t) { return WDM_ERR_UNSUPPORTED; }
}  // namespace wdm
