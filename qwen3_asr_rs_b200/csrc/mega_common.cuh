// mega_common.cuh -- device helpers shared by the fused decode steps (decode_mega.cu: one sequence per launch,
// decode_batch.cu: NB sequences per launch): mbarrier / bulk-copy PTX, the tagged-word exchange, the weight ring.
#pragma once
#include "internal.h"

namespace asrb {
namespace mega {

static constexpr int NCONS_WARPS = 8;
static constexpr int NCONS = NCONS_WARPS * 32;          // 256 consumer threads
static constexpr int NTHREADS = NCONS + 32;             // + 1 producer warp
static constexpr int SLOT_BYTES = 32 * 1024;           // 16 rows of K = 1024: one row per half-warp and pass
static constexpr int NSLOT_MAX = 4;                     // weight ring: 4 x 32 KB in flight per SM (3 for the 1.7B dims: larger vectors)
static constexpr int KV_KEYS = 64;                      // keys per attention split (K and V tiles staged in smem)
static constexpr int KV_TILE_BYTES = KV_KEYS * 128 * 4; // 32 KB each for K and V (fp32 cache)
static constexpr int XS_MIN = 3072;                     // activation vector / attention scratch: max(I, XS_MIN) + 64 floats
static constexpr int XRES_MAX = 64;
static constexpr int MAX_LAYERS = 32;                   // layer table staged in shared memory                     // residual rows owned by one CTA (H / gridDim.x, rounded up)
static constexpr int HD = 128;
static constexpr int PSTRIDE = HD + 2;                  // partial record: o[128], m, l
static constexpr int DBG_SLOTS = 1024;
static constexpr int MAX_SPLITS = 18;                   // 64-key attention splits per kv head: contexts up to 1152 keys (148 SMs / 8 kv heads = 18)

// phases (3 bits of the tag)
enum { PH_QKV = 1, PH_PART = 2, PH_ATTN = 3, PH_XO = 4, PH_ACT = 5, PH_XD = 6 };

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// ---- tagged exchange ({fp32 value, tag} in one 64-bit word) ------------------------------------
// Publication is a fire-and-forget 64-bit `red.max`: the tag sits in the upper 32 bits and grows monotonically for
// every word (epoch, then layer, then phase), so the new word is always the maximum, i.e. the reduction acts as an
// exchange.  Reductions are performed at L2 as soon as they are issued and return nothing: plain/volatile stores
// were measured to linger for microseconds, a release fence costs ~1 us, and `atom.exch` (which returns the old
// value) serialised each warp's publications on the atomic round trip (~1000 cycles per GEMV row).
__device__ __forceinline__ void ll_store(uint2* p, float v, uint32_t tag) {
    const unsigned long long val = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(val) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const uint2* p) {      // two consecutive words (16-byte aligned)
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ll_poll1(const uint2* p, uint32_t tag) {
    uint2 v;
    do {
        asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    } while (v.y != tag);
    return __uint_as_float(v.x);
}
// four words at p, p+stride, ... : all loads are issued before any tag is examined (one round trip when ready)
__device__ __forceinline__ void ll_poll4(const uint2* p, int stride, uint32_t tag, float (&out)[4]) {
    uint2 v[4];
    bool ok;
    do {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[i].x), "=r"(v[i].y) : "l"(p + (size_t)i * stride) : "memory");
        ok = (v[0].y == tag) && (v[1].y == tag) && (v[2].y == tag) && (v[3].y == tag);
    } while (!ok);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = __uint_as_float(v[i].x);
}
// Activation vectors in shared memory are stored with 16-byte group k at k ^ ((k >> 3) & 1).  The GEMV register loads
// read, per lane, the two groups of 8 consecutive elements (32-byte lane stride): unswizzled, lanes i and i+4 of every
// quarter warp hit the same banks (2-way conflict on every LDS.128, ~1000 cycles per phase for 8 warps x 8 KB).
__device__ __forceinline__ int xs_swz(int e) { const int k = e >> 2; return ((k ^ ((k >> 3) & 1)) << 2) | (e & 3); }
// all consumer threads: gather n (even) tagged values into shared memory; returns this thread's sum of squares
__device__ __forceinline__ float ll_gather(const uint2* buf, int n, uint32_t tag, float* xs) {
    float ss = 0.f;
    const int pairs = n >> 1;
    constexpr int U = 8;                                           // independent 16-byte loads in flight per thread
    for (int i0 = threadIdx.x; i0 < pairs; i0 += U * NCONS) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = make_uint4(0u, 0u, 0u, 0u);      // tag 0 is never published (epochs start at 1)
        bool ok;
        do {
            ok = true;
            // only the pairs still missing are re-read (the registers themselves say which): every CTA reads every word, so
            // a full re-poll costs n x 8 B x gridDim.x of L2 bandwidth per round -- 3.6 MB for the 3072 SwiGLU activations
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * NCONS;
                if (i < pairs && !(v[u].y == tag && v[u].w == tag)) v[u] = ll_load2(buf + 2 * i);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * NCONS;
                if (i < pairs) ok = ok && (v[u].y == tag) && (v[u].w == tag);
            }
        } while (!ok);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * NCONS;
            if (i < pairs) {
                const float a = __uint_as_float(v[u].x), b = __uint_as_float(v[u].z);
                *reinterpret_cast<float2*>(xs + xs_swz(2 * i)) = make_float2(a, b);
                ss = fmaf(a, a, ss); ss = fmaf(b, b, ss);
            }
        }
    }
    return ss;
}

struct Ring {
    uint8_t* slots; uint64_t* full; uint64_t* empty;
    uint32_t nslot;         // ring depth (compile-time constant of the instantiation, propagated through inlining)
};

// one weight phase as seen by a CTA: rows [r0, r1) of W[N][K]
struct Slice { const bf16* W; int K, r0, r1, rpc; };
__device__ __forceinline__ Slice make_slice(const bf16* W, int N, int K, int rstep) {
    Slice s; s.W = W; s.K = K;
    const unsigned units = (unsigned)(N / rstep);      // units * gridDim.x < 2^32 for every matrix of the model
    const int u0 = (int)((blockIdx.x * units) / gridDim.x), u1 = (int)(((blockIdx.x + 1) * units) / gridDim.x);
    s.r0 = u0 * rstep; s.r1 = u1 * rstep;
    s.rpc = SLOT_BYTES / (K * 2);
    s.rpc &= ~1;                        // keep (gate, up) pairs together
    return s;
}

__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

}  // namespace mega
}  // namespace asrb
