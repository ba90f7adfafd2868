// decode_mega.cu -- one greedy decode iteration (inference.rs:160-200) as ONE persistent kernel.
//
// Why: at batch 1 a decoder forward is 1.19 GB of weights read once (HBM-bound, ~190 us at the
// measured 6.5 TB/s) but has ~141 dependent phases (28 layers x {qkv, attention, o_proj, gate/up,
// down} + lm_head); launched as separate kernels each phase pays launch latency + a cold start of
// the weight stream.  Here one CTA per SM stays resident for the whole step:
//   * a producer warp streams this CTA's slice of EVERY weight matrix, in phase order, into a
//     shared-memory ring with cp.async.bulk (TMA bulk copy) + mbarrier transaction counts.  Weight
//     addresses do not depend on activations, so the producer runs AHEAD across phase boundaries:
//     HBM stays busy while consumers wait at a grid barrier.
//   * 8 consumer warps do the fp32 GEMV from shared memory (activation vector held in registers,
//     bf16 -> fp32 up-cast is exact), phases separated by a device-wide barrier (one atomic + spin).
//   * attention (QK-RMSNorm + RoPE + KV append + softmax.V) is split over kv-heads x ctx splits;
//     partials are merged by the consumers of the o_proj phase.
//   * the last CTA to finish the lm_head performs the greedy bookkeeping (argmax, EOS, append,
//     embedding of the next token), so no host sync and no extra launch per token.
// Reference semantics per phase: see decode.cu.  Batch 1 only; other batches use decode.cu.
#include "internal.h"

namespace asrb {

namespace mega {

static constexpr int NCONS_WARPS = 8;
static constexpr int NCONS = NCONS_WARPS * 32;          // 256 consumer threads
static constexpr int NTHREADS = NCONS + 32;             // + 1 producer warp
static constexpr int SLOT_BYTES = 24 * 1024;
static constexpr int NSLOT = 5;                         // weight ring: 120 KB in flight per SM
static constexpr int KV_KEYS = 64;                      // keys per attention split (K and V tiles staged in smem)
static constexpr int KV_TILE_BYTES = KV_KEYS * 128 * 4; // 32 KB each for K and V (fp32 cache)
static constexpr int XS_FLOATS = 3072 + 64;             // activation vector / attention scratch

struct Params {
    const DecLayerW* layers;     // device array [L]
    const bf16* lm_head;
    const bf16* embed;
    const float* final_norm;
    const float* rope_cos; const float* rope_sin;
    float eps;
    int L, H, QD, KVD, I, V, nq, nkv, group;
    // state
    float* x; float* qkv; float* act;
    float* attn_part;            // [nkv*nsplit][group][HD+2]
    float* kcache; float* vcache; size_t cache_layer_stride; int max_ctx;
    int nsplit;
    float* part_val; int* part_idx;     // [gridDim.x]
    int* pos; int* done; int* next_id; int* ids_out; int* n_out; int max_new;
    unsigned* bar;               // [0] grid barrier counter, [1] finish ticket
    long long* dbg;              // optional timeline [2][DBG_SLOTS] of clock64 (CTA 0 and CTA G-1), else null
};
static constexpr int DBG_SLOTS = 512;

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
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// device-wide barrier among the consumer threads of all CTAs (all CTAs are co-resident:
// cooperative launch, one CTA per SM)
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned target) {
    cons_sync();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (ld_acquire(ctr) < target) { }
        __threadfence();
    }
    cons_sync();
}

struct Ring {
    uint8_t* slots; uint64_t* full; uint64_t* empty;
};

// one weight phase as seen by a CTA: rows [r0, r1) of W[N][K]
struct Slice { const bf16* W; int K, r0, r1, rpc; };
__device__ __forceinline__ Slice make_slice(const bf16* W, int N, int K, int rstep) {
    Slice s; s.W = W; s.K = K;
    int units = N / rstep;
    int u0 = (int)(((long long)blockIdx.x * units) / gridDim.x), u1 = (int)(((long long)(blockIdx.x + 1) * units) / gridDim.x);
    s.r0 = u0 * rstep; s.r1 = u1 * rstep;
    s.rpc = SLOT_BYTES / (K * 2);
    s.rpc &= ~1;                        // keep (gate, up) pairs together
    return s;
}
__device__ __forceinline__ int n_chunks(const Slice& s) { return (s.r1 - s.r0 + s.rpc - 1) / s.rpc; }

// producer: issue all chunks of a slice
__device__ __forceinline__ void produce(const Slice& s, const Ring& ring, uint32_t& q) {
    for (int r = s.r0; r < s.r1; r += s.rpc, ++q) {
        int rows = min(s.rpc, s.r1 - r);
        uint32_t slot = q % NSLOT, par = (q / NSLOT) & 1;
        mbar_wait(&ring.empty[slot], par ^ 1);
        uint32_t bytes = (uint32_t)rows * s.K * 2;
        mbar_expect_tx(&ring.full[slot], bytes);
        bulk_g2s(ring.slots + (size_t)slot * SLOT_BYTES, s.W + (size_t)r * s.K, bytes, &ring.full[slot]);
    }
}

template <int K>
__device__ __forceinline__ void load_xr(const float* xs, float (&xr)[K / 32], int lane) {
#pragma unroll
    for (int c = 0; c < K / 256; ++c) {
        const float4 a = *reinterpret_cast<const float4*>(xs + (c * 32 + lane) * 8);
        const float4 b = *reinterpret_cast<const float4*>(xs + (c * 32 + lane) * 8 + 4);
        xr[c * 8 + 0] = a.x; xr[c * 8 + 1] = a.y; xr[c * 8 + 2] = a.z; xr[c * 8 + 3] = a.w;
        xr[c * 8 + 4] = b.x; xr[c * 8 + 5] = b.y; xr[c * 8 + 6] = b.z; xr[c * 8 + 7] = b.w;
    }
}
template <int K>
__device__ __forceinline__ float row_dot(const uint4* wrow, const float (&xr)[K / 32], int lane) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < K / 256; ++c) {
        const uint4 w = wrow[c * 32 + lane];
        a0 = fmaf(bf16_lo(w.x), xr[c * 8 + 0], a0); a1 = fmaf(bf16_hi(w.x), xr[c * 8 + 1], a1);
        a0 = fmaf(bf16_lo(w.y), xr[c * 8 + 2], a0); a1 = fmaf(bf16_hi(w.y), xr[c * 8 + 3], a1);
        a0 = fmaf(bf16_lo(w.z), xr[c * 8 + 4], a0); a1 = fmaf(bf16_hi(w.z), xr[c * 8 + 5], a1);
        a0 = fmaf(bf16_lo(w.w), xr[c * 8 + 6], a0); a1 = fmaf(bf16_hi(w.w), xr[c * 8 + 7], a1);
    }
    return warp_sum(a0 + a1);
}

enum { ME_STORE = 0, ME_RESID = 1, ME_SWIGLU = 2, ME_ARGMAX = 3 };

// consumer: process all chunks of a slice.  `xs` holds the (already normalised) activation vector.
template <int K, int EPI>
__device__ __forceinline__ void consume(const Slice& s, const Ring& ring, uint32_t& q, const float* xs, float* out,
                                        float& best_v, int& best_i) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RSTEP = (EPI == ME_SWIGLU) ? 2 : 1;
    float xr[K / 32];
    load_xr<K>(xs, xr, lane);
    int unit = 0;                                  // unit index within this CTA's slice
    for (int r = s.r0; r < s.r1; r += s.rpc, ++q) {
        const int rows = min(s.rpc, s.r1 - r);
        const uint32_t slot = q % NSLOT, par = (q / NSLOT) & 1;
        mbar_wait(&ring.full[slot], par);
        const uint4* base = reinterpret_cast<const uint4*>(ring.slots + (size_t)slot * SLOT_BYTES);
        const int units_here = rows / RSTEP;
        // units are dealt round-robin to warps across the whole slice
        int first = (warp - (unit % NCONS_WARPS) + NCONS_WARPS) % NCONS_WARPS;
        for (int u = first; u < units_here; u += NCONS_WARPS) {
            const int row = r + u * RSTEP;
            float v0 = row_dot<K>(base + (size_t)(u * RSTEP) * (K / 8), xr, lane);
            if (EPI == ME_SWIGLU) {
                float v1 = row_dot<K>(base + (size_t)(u * RSTEP + 1) * (K / 8), xr, lane);
                if (lane == 0) out[row >> 1] = silu(v0) * v1;
            } else if (EPI == ME_STORE) {
                if (lane == 0) out[row] = v0;
            } else if (EPI == ME_RESID) {
                if (lane == 0) out[row] = __ldcg(out + row) + v0;
            } else {
                if (v0 > best_v) { best_v = v0; best_i = row; }
            }
        }
        unit += units_here;
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[slot]);
    }
}

// RMSNorm of a global fp32 vector into shared memory (all consumer threads)
__device__ __forceinline__ void norm_to_smem(const float* __restrict__ x, const float* __restrict__ w, int n, float eps,
                                             float* xs, float* red) {
    const int tid = threadIdx.x;
    float s = 0.f;
    for (int i = tid; i < n; i += NCONS) { float v = __ldcg(x + i); xs[i] = v; s = fmaf(v, v, s); }
    s = warp_sum(s);
    cons_sync();
    if ((tid & 31) == 0) red[tid >> 5] = s;
    cons_sync();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NCONS_WARPS; ++i) tot += red[i];
    const float r = 1.0f / sqrtf(tot / n + eps);
    for (int i = tid; i < n; i += NCONS) xs[i] = (xs[i] * r) * w[i];
    cons_sync();
}

// per-head RMSNorm + RoPE of one 128-vector by one warp (lane holds d = lane, +32, +64, +96)
__device__ __forceinline__ void head_norm_rope(const float* __restrict__ src, const float* __restrict__ nw, float eps,
                                               const float* __restrict__ cs, const float* __restrict__ sn, float* dst, int lane) {
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __ldcg(src + lane + 32 * i);
    float ss = warp_sum(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    const float r = 1.0f / sqrtf(ss / 128.f + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (v[i] * r) * nw[lane + 32 * i];
    // pairs (d, d+64): (lane, lane+64) and (lane+32, lane+96)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int d = lane + 32 * i;
        const float c = cs[d], s = sn[d];
        const float a = v[i], b = v[i + 2];
        dst[d] = a * c - b * s;
        dst[d + 64] = b * c + a * s;
    }
}

#define MEGA_MARK()                                                                                    \
    do {                                                                                               \
        if (dbg_row && tid == 0 && dbg_i < DBG_SLOTS) dbg_row[dbg_i++] = clock64();                    \
    } while (0)

template <int H, int QD, int I>
__global__ void __launch_bounds__(NTHREADS, 1) decode_step_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Ring ring;
    ring.slots = smem;
    uint8_t* kv_smem = smem + (size_t)NSLOT * SLOT_BYTES;              // [K tile | V tile]
    float* xs = reinterpret_cast<float*>(kv_smem + 2 * KV_TILE_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xs + XS_FLOATS);
    ring.full = bars; ring.empty = bars + NSLOT;
    uint64_t* kv_full = bars + 2 * NSLOT; uint64_t* kv_empty = kv_full + 1;
    float* red = reinterpret_cast<float*>(bars + 2 * NSLOT + 2);      // [64]
    int* ired = reinterpret_cast<int*>(red + 64);                      // [64]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_producer = warp == NCONS_WARPS;

    if (__ldcg(p.done) != 0) return;            // sequence finished: nothing to do this step

    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) { mbar_init(&ring.full[i], 1); mbar_init(&ring.empty[i], NCONS_WARPS); }
        mbar_init(kv_full, 1); mbar_init(kv_empty, NCONS_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // attention work item of this CTA: (kv head att_g, keys [att_j0, att_j0 + KV_KEYS)); keys < pos are
    // already in the cache (n_old of them fall in this split), key `pos` is produced in this step.
    const int pos = __ldcg(p.pos);
    const bool att_cta = (int)blockIdx.x < p.nkv * p.nsplit;
    const int att_g = blockIdx.x / p.nsplit, att_j0 = (blockIdx.x % p.nsplit) * KV_KEYS;
    const int n_old = att_cta ? max(0, min(pos - att_j0, KV_KEYS)) : 0;
    uint32_t kvq = 0;
    uint32_t q = 0;
    if (is_producer) {
        if (lane == 0) {
            for (int l = 0; l < p.L; ++l) {
                const DecLayerW w = p.layers[l];
                produce(make_slice(w.wqkv, QD + 2 * p.KVD, H, 1), ring, q);
                if (n_old > 0) {   // K/V rows of earlier positions do not depend on this step: prefetch them too
                    mbar_wait(kv_empty, (kvq & 1) ^ 1);
                    const uint32_t bytes = (uint32_t)n_old * 128 * 4;
                    mbar_expect_tx(kv_full, 2 * bytes);
                    const size_t off = (size_t)l * p.cache_layer_stride + ((size_t)att_g * p.max_ctx + att_j0) * 128;
                    bulk_g2s(kv_smem, p.kcache + off, bytes, kv_full);
                    bulk_g2s(kv_smem + KV_TILE_BYTES, p.vcache + off, bytes, kv_full);
                    ++kvq;
                }
                produce(make_slice(w.wo, H, QD, 1), ring, q);
                produce(make_slice(w.wgu, 2 * I, H, 2), ring, q);
                produce(make_slice(w.wdown, H, I, 1), ring, q);
            }
            produce(make_slice(p.lm_head, p.V, H, 1), ring, q);
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    long long* dbg_row = nullptr; int dbg_i = 0;
    if (p.dbg && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) dbg_row = p.dbg + (blockIdx.x == 0 ? 0 : DBG_SLOTS);
    MEGA_MARK();
    const int HD = 128, half = 64;
    const float* cs = p.rope_cos + (size_t)pos * half;
    const float* sn = p.rope_sin + (size_t)pos * half;
    unsigned bar_target = 0;
    const unsigned G = gridDim.x;
    float best_v = -INFINITY; int best_i = 0x7fffffff;
    const int PSTRIDE = HD + 2;

    for (int l = 0; l < p.L; ++l) {
        const DecLayerW w = p.layers[l];
        // ---- phase 1: RMSNorm + [q|k|v] GEMV ----
        norm_to_smem(p.x, w.ln_in, H, p.eps, xs, red);
        consume<H, ME_STORE>(make_slice(w.wqkv, QD + 2 * p.KVD, H, 1), ring, q, xs, p.qkv, best_v, best_i);
        MEGA_MARK();
        bar_target += G; grid_sync(p.bar, bar_target);
        MEGA_MARK();
        // ---- phase 2: attention partials, work item = (kv head, 64-key split) ----
        {
            const int nk = pos + 1;                              // keys 0..pos
            const int nloc = att_cta ? max(0, min(nk - att_j0, KV_KEYS)) : 0;   // keys of this split incl. the new one
            if (nloc > 0) {
                const int g = att_g;
                float* qs = xs;                       // [group][128]
                float* kn = qs + p.group * HD;        // [128]
                float* vn = kn + HD;                  // [128]
                float* sc = vn + HD;                  // [group][KV_KEYS]
                float* ex = sc + p.group * KV_KEYS;   // [group][128]
                float* Ks = reinterpret_cast<float*>(kv_smem);
                float* Vs = reinterpret_cast<float*>(kv_smem + KV_TILE_BYTES);
                const bool has_new = (pos >= att_j0) && (pos < att_j0 + KV_KEYS);
                if (warp < p.group) head_norm_rope(p.qkv + (size_t)(g * p.group + warp) * HD, w.qnorm, p.eps, cs, sn, qs + warp * HD, lane);
                else if (warp == p.group && has_new) head_norm_rope(p.qkv + QD + (size_t)g * HD, w.knorm, p.eps, cs, sn, kn, lane);
                else if (warp == p.group + 1 && has_new) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) vn[lane + 32 * i] = __ldcg(p.qkv + QD + p.KVD + (size_t)g * HD + lane + 32 * i);
                }
                if (n_old > 0) mbar_wait(kv_full, kvq & 1);      // prefetched K/V tiles have landed
                cons_sync();
                if (has_new && tid < HD) {            // KV append (replaces Tensor::cat, layers.rs:311-317)
                    float* kc = p.kcache + (size_t)l * p.cache_layer_stride + ((size_t)g * p.max_ctx + pos) * HD;
                    float* vc = p.vcache + (size_t)l * p.cache_layer_stride + ((size_t)g * p.max_ctx + pos) * HD;
                    const float kx = kn[tid], vx = vn[tid];
                    kc[tid] = kx; vc[tid] = vx;
                    Ks[(pos - att_j0) * HD + tid] = kx; Vs[(pos - att_j0) * HD + tid] = vx;
                }
                cons_sync();
                const float div = sqrtf((float)HD);
                for (int j = warp; j < nloc; j += NCONS_WARPS) {
                    const float4 kv = *reinterpret_cast<const float4*>(Ks + j * HD + lane * 4);
                    for (int hq = 0; hq < p.group; ++hq) {
                        const float4 qv = *reinterpret_cast<const float4*>(qs + hq * HD + lane * 4);
                        float dot = warp_sum(kv.x * qv.x + kv.y * qv.y + kv.z * qv.z + kv.w * qv.w);
                        if (lane == 0) sc[hq * KV_KEYS + j] = dot / div;
                    }
                }
                cons_sync();
                if (warp < p.group) {                 // softmax partial of head `warp` over this split
                    float mx = -INFINITY;
                    for (int j = lane; j < nloc; j += 32) mx = fmaxf(mx, sc[warp * KV_KEYS + j]);
                    mx = warp_max(mx);
                    float sum = 0.f;
                    for (int j = lane; j < nloc; j += 32) {
                        float e = expf(sc[warp * KV_KEYS + j] - mx);
                        sc[warp * KV_KEYS + j] = e; sum += e;
                    }
                    sum = warp_sum(sum);
                    if (lane == 0) {
                        float* pp = p.attn_part + ((size_t)blockIdx.x * p.group + warp) * PSTRIDE;
                        pp[HD] = mx; pp[HD + 1] = sum;
                    }
                }
                cons_sync();
                {   // o[hq][d] = sum_j e[hq][j] * V[j][d]; thread = (d, key-parity partition)
                    const int d = tid & (HD - 1), part = tid >> 7;
                    float acc[8];
#pragma unroll
                    for (int hq = 0; hq < 8; ++hq) acc[hq] = 0.f;
                    for (int j = part; j < nloc; j += 2) {
                        const float vv = Vs[j * HD + d];
#pragma unroll
                        for (int hq = 0; hq < 8; ++hq) if (hq < p.group) acc[hq] = fmaf(sc[hq * KV_KEYS + j], vv, acc[hq]);
                    }
                    if (part == 1) for (int hq = 0; hq < p.group; ++hq) ex[hq * HD + d] = acc[hq];
                    cons_sync();
                    if (part == 0)
                        for (int hq = 0; hq < p.group; ++hq)
                            p.attn_part[((size_t)blockIdx.x * p.group + hq) * PSTRIDE + d] = acc[hq] + ex[hq * HD + d];
                }
                if (n_old > 0) {                      // hand the K/V staging buffer back to the producer
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(kv_empty);
                    ++kvq;
                }
            }
        }
        MEGA_MARK();
        bar_target += G; grid_sync(p.bar, bar_target);
        MEGA_MARK();
        // ---- phase 3: merge attention partials -> o_proj GEMV + residual ----
        {
            const int nact = min(p.nsplit, (pos + KV_KEYS) / KV_KEYS);     // splits holding at least one of keys 0..pos
            for (int o = tid; o < QD; o += NCONS) {
                const int h = o / HD, d = o - h * HD;
                const int g = h / p.group, hq = h - g * p.group;
                float M = -INFINITY;
                for (int s2 = 0; s2 < nact; ++s2) M = fmaxf(M, __ldcg(p.attn_part + ((size_t)(g * p.nsplit + s2) * p.group + hq) * PSTRIDE + HD));
                float Lsum = 0.f, O = 0.f;
                for (int s2 = 0; s2 < nact; ++s2) {
                    const float* pp = p.attn_part + ((size_t)(g * p.nsplit + s2) * p.group + hq) * PSTRIDE;
                    const float sc_ = expf(__ldcg(pp + HD) - M);
                    Lsum = fmaf(__ldcg(pp + HD + 1), sc_, Lsum);
                    O = fmaf(__ldcg(pp + d), sc_, O);
                }
                xs[o] = O / Lsum;
            }
        }
        cons_sync();
        consume<QD, ME_RESID>(make_slice(w.wo, H, QD, 1), ring, q, xs, p.x, best_v, best_i);
        MEGA_MARK();
        bar_target += G; grid_sync(p.bar, bar_target);
        MEGA_MARK();
        // ---- phase 4: RMSNorm + gate/up GEMV + SiLU*mul ----
        norm_to_smem(p.x, w.ln_post, H, p.eps, xs, red);
        consume<H, ME_SWIGLU>(make_slice(w.wgu, 2 * I, H, 2), ring, q, xs, p.act, best_v, best_i);
        MEGA_MARK();
        bar_target += G; grid_sync(p.bar, bar_target);
        MEGA_MARK();
        // ---- phase 5: down GEMV + residual ----
        for (int i = tid; i < I; i += NCONS) xs[i] = __ldcg(p.act + i);
        cons_sync();
        consume<I, ME_RESID>(make_slice(w.wdown, H, I, 1), ring, q, xs, p.x, best_v, best_i);
        MEGA_MARK();
        bar_target += G; grid_sync(p.bar, bar_target);
        MEGA_MARK();
    }
    // ---- final RMSNorm + tied lm_head GEMV + argmax ----
    norm_to_smem(p.x, p.final_norm, H, p.eps, xs, red);
    consume<H, ME_ARGMAX>(make_slice(p.lm_head, p.V, H, 1), ring, q, xs, nullptr, best_v, best_i);
    MEGA_MARK();
    // every lane of a warp saw the same values: lane 0 publishes the warp's best
    cons_sync();
    if (lane == 0) { red[warp] = best_v; ired[warp] = best_i; }
    cons_sync();
    __shared__ int is_last;
    if (tid == 0) {
        float v = -INFINITY; int idx = 0x7fffffff;
        for (int wq = 0; wq < NCONS_WARPS; ++wq)
            if (red[wq] > v || (red[wq] == v && ired[wq] < idx)) { v = red[wq]; idx = ired[wq]; }
        p.part_val[blockIdx.x] = v; p.part_idx[blockIdx.x] = idx;
        __threadfence();
        unsigned t = atomicAdd(p.bar + 1, 1u);
        is_last = (t == G - 1);
    }
    cons_sync();
    if (!is_last) return;
    // ---- greedy bookkeeping by the last CTA (inference.rs:161-170) ----
    __threadfence();
    {
        float v = -INFINITY; int idx = 0x7fffffff;
        for (int i = tid; i < (int)G; i += NCONS) {
            float pv = __ldcg(p.part_val + i); int pi = __ldcg(p.part_idx + i);
            if (pv > v || (pv == v && pi < idx)) { v = pv; idx = pi; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(0xffffffffu, v, o); int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        if (lane == 0) { red[warp] = v; ired[warp] = idx; }
        cons_sync();
        __shared__ int tok_s;
        if (tid == 0) {
            for (int wq = 1; wq < NCONS_WARPS; ++wq)
                if (red[wq] > v || (red[wq] == v && ired[wq] < idx)) { v = red[wq]; idx = ired[wq]; }
            int tok = idx;
            const int n = *p.n_out;
            if (tok == 151643 || tok == 151645 || n >= p.max_new) { *p.done = 1; *p.next_id = -1; tok = -1; }
            else { p.ids_out[n] = tok; *p.n_out = n + 1; *p.pos = pos + 1; *p.next_id = tok; }
            tok_s = tok;
            p.bar[0] = 0; p.bar[1] = 0;          // all CTAs are past every barrier: reset for the next launch
        }
        cons_sync();
        const int tok = tok_s;
        if (tok >= 0) {
            const bf16* e = p.embed + (size_t)tok * H;
            for (int i = tid; i < H; i += NCONS) p.x[i] = __bfloat162float(e[i]);
        }
    }
}

}  // namespace mega

// host side ---------------------------------------------------------------------------------------
struct MegaState {
    DecLayerW* d_layers = nullptr; unsigned* d_bar = nullptr; float* d_part = nullptr; long long* d_dbg = nullptr;
    int part_cap = 0; const Model* model = nullptr;
};
static MegaState g_mega;   // one model per process in practice; re-created when the model changes

static size_t mega_smem_bytes() {
    return (size_t)mega::NSLOT * mega::SLOT_BYTES + 2 * mega::KV_TILE_BYTES + mega::XS_FLOATS * 4 + (2 * mega::NSLOT + 2) * 8 +
           64 * 4 + 64 * 4 + 64;
}

template <int H, int QD, int I> static bool dims_match(const asrb_dims& c) {
    return c.hidden_size == H && c.num_attention_heads * c.head_dim == QD && c.intermediate_size == I;
}

bool decode_mega_supported(const Model& m, int B, int max_ctx) {
    const asrb_dims& c = m.d.c;
    if (B != 1 || c.head_dim != 128) return false;
    const int group = c.num_attention_heads / c.num_key_value_heads;
    if (group + 2 > mega::NCONS_WARPS) return false;
    if ((size_t)(group * 128 + 256 + group * mega::KV_KEYS + group * 128) > (size_t)mega::XS_FLOATS) return false;
    if (m.ctx->smem_optin < mega_smem_bytes()) return false;
    if (((max_ctx + mega::KV_KEYS - 1) / mega::KV_KEYS) * c.num_key_value_heads > m.ctx->sm_count) return false;   // one CTA per (kv head, 64-key split)
    return dims_match<1024, 2048, 3072>(c) || dims_match<256, 512, 512>(c);
}

void launch_decode_step_mega(const Model& m, const DecodeBufs& b, int B, float* kcache, float* vcache,
                             size_t cache_layer_stride, size_t cache_seq_stride, int max_ctx, cudaStream_t st,
                             int64_t* launches) {
    (void)cache_seq_stride;
    ASRB_REQUIRE(decode_mega_supported(m, B, max_ctx), ASRB_ERR_STATE, "fused decode step not supported for this model/batch");
    const asrb_dims& c = m.d.c;
    const int G = m.ctx->sm_count;
    const int group = c.num_attention_heads / c.num_key_value_heads;
    const int nsplit = (max_ctx + mega::KV_KEYS - 1) / mega::KV_KEYS;
    ASRB_REQUIRE(nsplit * c.num_key_value_heads <= G, ASRB_ERR_INVALID, "context too long for the fused decode step");
    if (g_mega.model != &m) {
        if (g_mega.d_layers) { cudaFree(g_mega.d_layers); cudaFree(g_mega.d_bar); g_mega.d_layers = nullptr; }
        ASRB_CUDA_CHECK(cudaMalloc(&g_mega.d_layers, m.dec.size() * sizeof(DecLayerW)));
        ASRB_CUDA_CHECK(cudaMemcpy(g_mega.d_layers, m.dec.data(), m.dec.size() * sizeof(DecLayerW), cudaMemcpyHostToDevice));
        ASRB_CUDA_CHECK(cudaMalloc(&g_mega.d_bar, 4 * sizeof(unsigned)));
        ASRB_CUDA_CHECK(cudaMemset(g_mega.d_bar, 0, 4 * sizeof(unsigned)));
        g_mega.model = &m;
    }
    const int need_part = G * group * (128 + 2);
    if (g_mega.part_cap < need_part) {
        if (g_mega.d_part) cudaFree(g_mega.d_part);
        ASRB_CUDA_CHECK(cudaMalloc(&g_mega.d_part, (size_t)need_part * sizeof(float)));
        g_mega.part_cap = need_part;
    }
    mega::Params p{};
    p.layers = g_mega.d_layers; p.lm_head = m.lm_head; p.embed = m.embed; p.final_norm = m.final_norm;
    p.rope_cos = m.rope_cos; p.rope_sin = m.rope_sin; p.eps = (float)c.rms_norm_eps;
    p.L = c.num_hidden_layers; p.H = c.hidden_size; p.QD = m.d.q_dim; p.KVD = m.d.kv_dim; p.I = c.intermediate_size;
    p.V = c.vocab_size; p.nq = c.num_attention_heads; p.nkv = c.num_key_value_heads; p.group = group;
    p.x = b.x; p.qkv = b.qkv; p.act = b.act; p.attn_part = g_mega.d_part;
    p.kcache = kcache; p.vcache = vcache; p.cache_layer_stride = cache_layer_stride; p.max_ctx = max_ctx; p.nsplit = nsplit;
    p.part_val = b.part_val; p.part_idx = b.part_idx;
    p.pos = b.pos; p.done = b.done; p.next_id = b.next_id; p.ids_out = b.ids_out; p.n_out = b.n_out; p.max_new = b.max_new;
    p.bar = g_mega.d_bar;
    if (getenv("ASRB_MEGA_DEBUG")) {
        if (!g_mega.d_dbg) { ASRB_CUDA_CHECK(cudaMalloc(&g_mega.d_dbg, 2 * mega::DBG_SLOTS * sizeof(long long)));
                             ASRB_CUDA_CHECK(cudaMemset(g_mega.d_dbg, 0, 2 * mega::DBG_SLOTS * sizeof(long long))); }
        p.dbg = g_mega.d_dbg;
    }
    const size_t smem = mega_smem_bytes();
    void* args[] = {(void*)&p};
    const void* fn = nullptr;
    if (dims_match<1024, 2048, 3072>(c)) fn = (const void*)mega::decode_step_kernel<1024, 2048, 3072>;
    else fn = (const void*)mega::decode_step_kernel<256, 512, 512>;
    ASRB_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ASRB_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(mega::NTHREADS), args, smem, st));
    if (launches) *launches += 1;
}

// debug: copy the clock64 timeline of the most recent fused step (CTA 0 then CTA G-1), returns slots per CTA
int decode_mega_debug_timeline(long long* out, int cap) {
    if (!g_mega.d_dbg || cap < 2 * mega::DBG_SLOTS) return 0;
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_mega.d_dbg, 2 * mega::DBG_SLOTS * sizeof(long long), cudaMemcpyDeviceToHost);
    return mega::DBG_SLOTS;
}

}  // namespace asrb
