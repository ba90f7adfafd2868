#include "internal.h"
namespace asrb {
bool launch_gemm_tc(const GemmA&, const bf16*, int, const GemmEpi&, cudaStream_t) { return false; }
}
