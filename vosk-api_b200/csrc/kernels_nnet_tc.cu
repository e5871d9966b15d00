// kernels_nnet_tc.cu — placeholder until the tcgen05 kernel lands (next commit): routes to the fp32 kernel.
#include "vb_kernels.h"
namespace vb {
extern "C" cudaError_t vbk_gemm_tc(const GemmArgs *a, cudaStream_t s) { return vbk_gemm_fp32(a, s); }
}
