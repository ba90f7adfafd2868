#include "internal.h"
namespace asrb {
bool decode_mega_supported(const Model&, int) { return false; }
void launch_decode_step_mega(const Model&, const DecodeBufs&, int, float*, float*, size_t, size_t, int, cudaStream_t, int64_t*) {
    throw Error(ASRB_ERR_STATE, "fused decode step not available");
}
}
