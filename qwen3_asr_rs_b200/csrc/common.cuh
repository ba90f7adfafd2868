// common.cuh -- shared device helpers for the sm_100a kernels of the Qwen3-ASR hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdexcept>
#include <string>

namespace asrb {

struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define ASRB_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            throw ::asrb::Error(2, std::string(#expr) + " failed: " + cudaGetErrorString(_e) + \
                                       " (" __FILE__ ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define ASRB_REQUIRE(cond, code, msg)                                          \
    do {                                                                       \
        if (!(cond)) throw ::asrb::Error((code), std::string(msg));            \
    } while (0)

typedef __nv_bfloat16 bf16;

// ---- bf16 <-> fp32 (bit-exact up-cast, as src/weights.rs:134-142) --------------------
__device__ __forceinline__ float bf16_bits_to_f32(uint32_t bits16) { return __uint_as_float(bits16 << 16); }
__device__ __forceinline__ float bf16_lo(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// ---- exact 3-way bf16 split of an fp32 activation: x == hi + mid + lo (to 1 ulp) -----
// GEMM A-operands are stored as three bf16 planes so tcgen05 kind::f16 MMAs with exact
// bf16 weights reproduce fp32 products (bf16*bf16 is exact in fp32; fp32 accumulate).
struct Split3 { bf16 hi, mid, lo; };
__device__ __forceinline__ Split3 split3(float x) {
    Split3 s;
    s.hi = __float2bfloat16_rn(x);
    float r = x - __bfloat162float(s.hi);
    s.mid = __float2bfloat16_rn(r);
    r = r - __bfloat162float(s.mid);
    s.lo = __float2bfloat16_rn(r);
    return s;
}
__device__ __forceinline__ void store_split3(bf16* base, size_t plane_stride, size_t idx, float x) {
    Split3 s = split3(x);
    base[idx] = s.hi;
    base[plane_stride + idx] = s.mid;
    base[2 * plane_stride + idx] = s.lo;
}
__device__ __forceinline__ float load_split3(const bf16* base, size_t plane_stride, size_t idx, int nplanes) {
    float v = __bfloat162float(base[idx]);
    if (nplanes > 1) v += __bfloat162float(base[plane_stride + idx]);
    if (nplanes > 2) v += __bfloat162float(base[2 * plane_stride + idx]);
    return v;
}

// ---- math matching the ATen CPU ops the reference's tch arm dispatches to --------------
__device__ __forceinline__ float gelu_erf(float x) {            // gelu("none"), src/tensor.rs:350
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }   // src/tensor.rs:354

// ---- warp / block reductions -----------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// blockDim.x multiple of 32, <= 1024; scratch >= 32 floats; result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : 0.f;
    r = warp_sum(r);
    return r;
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) scratch[w] = v;
    __syncthreads();
    float r = (lane < nw) ? scratch[lane] : -INFINITY;
    r = warp_max(r);
    return r;
}

// order-preserving float <-> int key (for atomicMax on floats of either sign)
__device__ __host__ __forceinline__ int float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    int i = __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4);
#endif
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int k) {
    int i = k >= 0 ? k : k ^ 0x7fffffff;
    return __int_as_float(i);
}

}  // namespace asrb
