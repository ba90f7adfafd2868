// decode_mega.cu -- one greedy decode iteration (inference.rs:160-200) as ONE persistent kernel.
//
// Why: at batch 1 a decoder forward is 1.19 GB of weights read once (HBM-bound, ~190 us at the
// measured 6.5 TB/s) but has ~141 dependent phases (28 layers x {qkv, attention, o_proj, gate/up,
// down} + lm_head).  As separate kernels each phase pays launch latency and a cold weight stream;
// with device-wide barriers each phase pays an atomic + spin + re-read (measured: 1.4 ms/step).
// Here one CTA per SM stays resident for the whole step:
//   * a producer warp streams this CTA's slice of EVERY weight matrix, in phase order, into a
//     shared-memory ring with cp.async.bulk (TMA bulk copy) + mbarrier transaction counts, and the
//     K/V rows of earlier positions into a staging tile.  None of these addresses depend on
//     activations, so the producer runs AHEAD across phase boundaries: HBM stays busy while
//     consumers wait for activations.
//   * 8 consumer warps do the fp32 GEMV from shared memory (activation vector held in registers,
//     bf16 -> fp32 up-cast is exact, fp32 FMA accumulate).
//   * activations are exchanged between CTAs WITHOUT barriers: every published fp32 value travels
//     in an 8-byte {value, tag} word (tag = launch epoch/layer/phase), published with one 64-bit
//     fire-and-forget `red.max` (performed at L2 immediately; the tag grows monotonically so max acts as
//     an exchange) and polled with 128-bit relaxed loads until the tags match -- data and "ready" flag
//     arrive in the same L2 round trip (the low-latency protocol of collective libraries, on-chip).
//   * attention (QK-RMSNorm + RoPE + KV append + softmax.V) is split over kv-heads x 64-key tiles;
//     the tile-0 CTA of each kv-head merges the partials and publishes the head outputs.
//   * the last CTA to finish the lm_head performs the greedy bookkeeping (argmax, EOS, append,
//     embedding of the next token), so there is no host sync and no extra launch per token.
// Reference semantics per phase: see decode.cu.  The kernel advances ONE sequence; a batch is B back-to-back launches
// (decode.cu remains the path for logits output, other model dimensions and contexts beyond 1152 keys).
#include "mega_common.cuh"

namespace asrb {

namespace mega {

struct Params {
    const DecLayerW* layers;     // device array [L]
    const bf16* lm_head;
    const bf16* embed;
    const float* final_norm;
    const float* rope_cos; const float* rope_sin;
    float eps;
    int L, H, QD, KVD, I, V, nq, nkv, group;
    // plain state (kernel-boundary visibility)
    float* x;                    // [H] embedding of the pending token (in) / of the next token (out)
    float* kcache; float* vcache; size_t cache_layer_stride; int max_ctx;
    int nsplit;
    float* part_val; int* part_idx;     // [gridDim.x]
    int* pos; int* done; int* next_id; int* ids_out; int* n_out; int max_new;
    unsigned* bar;               // [0] finish ticket, [1] launch epoch (starts at 1; 0 marks never-written words)
    // tagged exchange buffers ({value, tag} words)
    uint2* qkv_ll; uint2* part_ll; uint2* attn_ll; uint2* x_ll; uint2* act_ll;
    long long* dbg;              // optional timeline [2][DBG_SLOTS] of clock64 (CTA 0 and CTA G-1), else null
    int pf_ahead;                // L2 prefetch distance in ring chunks (0 = off)
};

// The shared-memory ring only buffers ~2.7 us of the weight stream (120 KB at this SM's 44 GB/s share of HBM), and a ring
// slot can only be refilled once its rows have been consumed: the refill of the slots a phase has just released must land
// before the phase after next needs them.  With the refill coming from DRAM (~2 us) that was hidden behind the slow
// all-gathers; once the gathers re-read only their missing words (ll_gather, mega_common.cuh) the GEMV turns started to wait
// ~600 cycles for their slots.  So the producer also runs an L2 prefetch cursor `pf_ahead` ring chunks ahead of the ring
// (ASRB_MEGA_PF overrides): the refill then comes from L2.  Measured with the selective re-poll (trip 34, us per step):
// distance 0: 452, 2: 432, 4: 432, 8: 436, 16: 512 (prefetched lines evicted before use: +31 % DRAM reads in round 1).
static constexpr int PF_AHEAD = 2;

template <int H, int QD, int I>
struct ChunkCursor {
    int l, ph, r; Slice s; bool done;
    const DecLayerW* ltab;
    __device__ void load(const Params& p) {
        if (l >= p.L) { if (l == p.L && ph == 0) s = make_slice(p.lm_head, p.V, H, 1); else { done = true; return; } }
        else {
            const DecLayerW w = ltab[l];
            s = ph == 0 ? make_slice(w.wqkv, QD + 2 * p.KVD, H, 1) : ph == 1 ? make_slice(w.wo, H, QD, 1)
              : ph == 2 ? make_slice(w.wgu, 2 * I, H, 2) : make_slice(w.wdown, H, I, 1);
        }
        r = s.r0;
    }
    __device__ void init(const Params& p, const DecLayerW* table) { ltab = table; l = 0; ph = 0; done = false; load(p); }
    __device__ bool next(const Params& p, const bf16*& src, uint32_t& bytes) {
        while (!done && r >= s.r1) {
            if (l >= p.L) { done = true; break; }
            if (++ph == 4) { ph = 0; ++l; }
            load(p);
        }
        if (done) return false;
        const int rows = min(s.rpc, s.r1 - r);
        src = s.W + (size_t)r * s.K; bytes = (uint32_t)rows * s.K * 2;
        r += s.rpc;
        return true;
    }
};

// producer: issue all chunks of a slice into the ring; every issued chunk advances the L2 prefetch cursor
template <class Cursor>
__device__ __forceinline__ void produce(const Slice& s, const Ring& ring, uint32_t& q, Cursor& pf, const Params& p) {
    for (int r = s.r0; r < s.r1; r += s.rpc, ++q) {
        int rows = min(s.rpc, s.r1 - r);
        uint32_t slot = q % ring.nslot, par = (q / ring.nslot) & 1;
        mbar_wait(&ring.empty[slot], par ^ 1);
        uint32_t bytes = (uint32_t)rows * s.K * 2;
        mbar_expect_tx(&ring.full[slot], bytes);
        bulk_g2s(ring.slots + (size_t)slot * SLOT_BYTES, s.W + (size_t)r * s.K, bytes, &ring.full[slot]);
        const bf16* psrc; uint32_t pbytes;
        if (p.pf_ahead > 0 && pf.next(p, psrc, pbytes)) l2_prefetch(psrc, pbytes);
    }
}

// GEMV row mapping: a row of K bf16 weights is contracted by one warp; lane `lane` holds the activations of elements
// (c * 32 + lane) * 8 .. +8 for every 256-element chunk c in registers.
template <int K>
__device__ __forceinline__ void load_xr(const float* xs, float (&xr)[K / 32], int lane) {
    const int sw = ((lane >> 2) & 1) * 4;          // xs_swz for this lane's two 16-byte groups: swapped when bit 3 of k is set
#pragma unroll
    for (int c = 0; c < K / 256; ++c) {
        const float4 a = *reinterpret_cast<const float4*>(xs + (c * 32 + lane) * 8 + sw);
        const float4 b = *reinterpret_cast<const float4*>(xs + (c * 32 + lane) * 8 + 4 - sw);
        xr[c * 8 + 0] = a.x; xr[c * 8 + 1] = a.y; xr[c * 8 + 2] = a.z; xr[c * 8 + 3] = a.w;
        xr[c * 8 + 4] = b.x; xr[c * 8 + 5] = b.y; xr[c * 8 + 6] = b.z; xr[c * 8 + 7] = b.w;
    }
}
// same, with the RMSNorm applied on the fly: xr = (x * r) * w  (rounding order of layers.rs:48-54)
template <int K>
__device__ __forceinline__ void load_xr_norm(const float* xs, const float* wn, float r, float (&xr)[K / 32], int lane) {
    const int sw = ((lane >> 2) & 1) * 4;          // both vectors are stored in the xs_swz layout
#pragma unroll
    for (int c = 0; c < K / 256; ++c) {
        const int o = (c * 32 + lane) * 8;
        const float4 a = *reinterpret_cast<const float4*>(xs + o + sw), b = *reinterpret_cast<const float4*>(xs + o + 4 - sw);
        const float4 wa = *reinterpret_cast<const float4*>(wn + o + sw), wb = *reinterpret_cast<const float4*>(wn + o + 4 - sw);
        xr[c * 8 + 0] = (a.x * r) * wa.x; xr[c * 8 + 1] = (a.y * r) * wa.y; xr[c * 8 + 2] = (a.z * r) * wa.z; xr[c * 8 + 3] = (a.w * r) * wa.w;
        xr[c * 8 + 4] = (b.x * r) * wb.x; xr[c * 8 + 5] = (b.y * r) * wb.y; xr[c * 8 + 6] = (b.z * r) * wb.z; xr[c * 8 + 7] = (b.w * r) * wb.w;
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
// two rows at once (independent FMA chains, one shared reduction): the result of row 0 ends up in lanes 0-15 and
// that of row 1 in lanes 16-31
template <int K>
__device__ __forceinline__ float row_dot2(const uint4* w0, const uint4* w1, const float (&xr)[K / 32], int lane) {
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
    for (int c = 0; c < K / 256; ++c) {
        const uint4 w = w0[c * 32 + lane], v = w1[c * 32 + lane];
        a0 = fmaf(bf16_lo(w.x), xr[c * 8 + 0], a0); a1 = fmaf(bf16_hi(w.x), xr[c * 8 + 1], a1);
        b0 = fmaf(bf16_lo(v.x), xr[c * 8 + 0], b0); b1 = fmaf(bf16_hi(v.x), xr[c * 8 + 1], b1);
        a0 = fmaf(bf16_lo(w.y), xr[c * 8 + 2], a0); a1 = fmaf(bf16_hi(w.y), xr[c * 8 + 3], a1);
        b0 = fmaf(bf16_lo(v.y), xr[c * 8 + 2], b0); b1 = fmaf(bf16_hi(v.y), xr[c * 8 + 3], b1);
        a0 = fmaf(bf16_lo(w.z), xr[c * 8 + 4], a0); a1 = fmaf(bf16_hi(w.z), xr[c * 8 + 5], a1);
        b0 = fmaf(bf16_lo(v.z), xr[c * 8 + 4], b0); b1 = fmaf(bf16_hi(v.z), xr[c * 8 + 5], b1);
        a0 = fmaf(bf16_lo(w.w), xr[c * 8 + 6], a0); a1 = fmaf(bf16_hi(w.w), xr[c * 8 + 7], a1);
        b0 = fmaf(bf16_lo(v.w), xr[c * 8 + 6], b0); b1 = fmaf(bf16_hi(v.w), xr[c * 8 + 7], b1);
    }
    const float ra = a0 + a1, rb = b0 + b1;
    const bool hi = lane & 16;
    float keep = (hi ? rb : ra) + __shfl_xor_sync(0xffffffffu, hi ? ra : rb, 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    return keep;
}

// same contraction with the activation vector read from shared memory (long K: keeps registers free)
template <int K>
__device__ __forceinline__ float row_dot_smem(const uint4* wrow, const float* xs, int lane) {
    float a0 = 0.f, a1 = 0.f;
    const int sw = ((lane >> 2) & 1) * 4;
#pragma unroll 4
    for (int c = 0; c < K / 256; ++c) {
        const uint4 w = wrow[c * 32 + lane];
        const float4 xa = *reinterpret_cast<const float4*>(xs + (c * 32 + lane) * 8 + sw);
        const float4 xb = *reinterpret_cast<const float4*>(xs + (c * 32 + lane) * 8 + 4 - sw);
        a0 = fmaf(bf16_lo(w.x), xa.x, a0); a1 = fmaf(bf16_hi(w.x), xa.y, a1);
        a0 = fmaf(bf16_lo(w.y), xa.z, a0); a1 = fmaf(bf16_hi(w.y), xa.w, a1);
        a0 = fmaf(bf16_lo(w.z), xb.x, a0); a1 = fmaf(bf16_hi(w.z), xb.y, a1);
        a0 = fmaf(bf16_lo(w.w), xb.z, a0); a1 = fmaf(bf16_hi(w.w), xb.w, a1);
    }
    return warp_sum(a0 + a1);
}


// four rows at once (K <= 1024, activations in registers): 8 independent FMA chains and ONE 6-shuffle transposed
// reduction for the four rows instead of two 5-shuffle reductions of row pairs.  The sum of row j ends up in all 8 lanes
// with (lane >> 3) == j.
template <int K>
__device__ __forceinline__ float row_dot4(const uint4* w0, const uint4* w1, const uint4* w2, const uint4* w3,
                                          const float (&xr)[K / 32], int lane) {
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f, c0 = 0.f, c1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int c = 0; c < K / 256; ++c) {
        const uint4 wa = w0[c * 32 + lane], wb = w1[c * 32 + lane], wc = w2[c * 32 + lane], wd = w3[c * 32 + lane];
#define ROW4_STEP(F, X0, X1)                                                                                          \
        a0 = fmaf(bf16_lo(wa.F), X0, a0); a1 = fmaf(bf16_hi(wa.F), X1, a1);                                           \
        b0 = fmaf(bf16_lo(wb.F), X0, b0); b1 = fmaf(bf16_hi(wb.F), X1, b1);                                           \
        c0 = fmaf(bf16_lo(wc.F), X0, c0); c1 = fmaf(bf16_hi(wc.F), X1, c1);                                           \
        d0 = fmaf(bf16_lo(wd.F), X0, d0); d1 = fmaf(bf16_hi(wd.F), X1, d1);
        ROW4_STEP(x, xr[c * 8 + 0], xr[c * 8 + 1])
        ROW4_STEP(y, xr[c * 8 + 2], xr[c * 8 + 3])
        ROW4_STEP(z, xr[c * 8 + 4], xr[c * 8 + 5])
        ROW4_STEP(w, xr[c * 8 + 6], xr[c * 8 + 7])
#undef ROW4_STEP
    }
    const float ra = a0 + a1, rb = b0 + b1, rc = c0 + c1, rd = d0 + d1;
    const bool h16 = lane & 16, h8 = lane & 8;
    float k0 = (h16 ? rc : ra) + __shfl_xor_sync(0xffffffffu, h16 ? ra : rc, 16);      // rows {0,1} stay low, {2,3} high
    float k1 = (h16 ? rd : rb) + __shfl_xor_sync(0xffffffffu, h16 ? rb : rd, 16);
    float keep = (h8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, h8 ? k0 : k1, 8);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    return keep;
}

// transposed warp reduction of 8 per-lane partials: returns, in every lane, the warp total of value index (lane >> 2)
__device__ __forceinline__ float warp_reduce8(const float (&a)[8], int lane) {
    const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
    float k[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (h16 ? a[4 + i] : a[i]) + __shfl_xor_sync(0xffffffffu, h16 ? a[i] : a[4 + i], 16);
    float m0 = (h8 ? k[2] : k[0]) + __shfl_xor_sync(0xffffffffu, h8 ? k[0] : k[2], 8);
    float m1 = (h8 ? k[3] : k[1]) + __shfl_xor_sync(0xffffffffu, h8 ? k[1] : k[3], 8);
    float v = (h4 ? m1 : m0) + __shfl_xor_sync(0xffffffffu, h4 ? m0 : m1, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

enum { ME_STORE = 0, ME_RESID = 1, ME_SWIGLU = 2, ME_ARGMAX = 3 };

// Residual GEMVs (o_proj, down_proj: a handful of rows per CTA, long K): all 8 warps split K of EVERY row instead of one
// warp per row.  A thread owns the 16-byte weight groups tid, tid + 256, ... of each row (its 8 activations per group come
// straight from xs), keeps one partial per row (8 independent FMA chains), the warp reduces its 8 partials with one
// transposed reduction (9 shuffles), the 8 warp partials of a row are summed in a fixed order by one thread, which applies
// the residual and publishes.  Rows are taken 8 at a time; a group spans two ring slots when a slot holds fewer than 8.
// (The one-warp-per-row form left 1 of 8 warps idle and paid the load -> unpack -> FMA -> 5-shuffle latency chain once per
// row with nothing to overlap it: 1.6 k / 1.9 k cycles per layer for 7 rows.)
template <int K>
__device__ __forceinline__ void consume_ksplit(const Slice& s, const Ring& ring, uint32_t& q, const float* xs, uint2* out,
                                               uint32_t tag, float* xres, float* part) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NU = K / 8;                                  // 16-byte groups per row
    constexpr int J = (NU + NCONS - 1) / NCONS;
    uint32_t grp = 0;
    for (int r = s.r0; r < s.r1;) {
        const int rowsA = min(s.rpc, s.r1 - r);
        const uint32_t slotA = q % ring.nslot;
        mbar_wait(&ring.full[slotA], (q / ring.nslot) & 1);
        const bool pair = s.rpc < 8 && r + rowsA < s.r1;
        const int rowsB = pair ? min(s.rpc, s.r1 - r - rowsA) : 0;
        const uint32_t slotB = (q + 1) % ring.nslot;
        if (pair) mbar_wait(&ring.full[slotB], ((q + 1) / ring.nslot) & 1);
        const uint4* baseA = reinterpret_cast<const uint4*>(ring.slots + (size_t)slotA * SLOT_BYTES);
        const uint4* baseB = reinterpret_cast<const uint4*>(ring.slots + (size_t)slotB * SLOT_BYTES);
        const int R = rowsA + rowsB;
        for (int g0 = 0; g0 < R; g0 += 8, ++grp) {
            const int ng = min(8, R - g0);
            const uint4* rp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ri = g0 + (i < ng ? i : 0);              // rows past the group re-read row 0 (result unused)
                rp[i] = ri < rowsA ? baseA + (size_t)ri * NU : baseB + (size_t)(ri - rowsA) * NU;
            }
            float acc[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int idx = tid + j * NCONS;
                if (idx < NU) {
                    const int sw = ((idx >> 2) & 1) * 4;            // xs_swz of the two 16-byte activation groups
                    const float4 xa = *reinterpret_cast<const float4*>(xs + idx * 8 + sw);
                    const float4 xb = *reinterpret_cast<const float4*>(xs + idx * 8 + 4 - sw);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint4 w = rp[i][idx];
                        float t = acc[i];
                        t = fmaf(bf16_lo(w.x), xa.x, t); t = fmaf(bf16_hi(w.x), xa.y, t);
                        t = fmaf(bf16_lo(w.y), xa.z, t); t = fmaf(bf16_hi(w.y), xa.w, t);
                        t = fmaf(bf16_lo(w.z), xb.x, t); t = fmaf(bf16_hi(w.z), xb.y, t);
                        t = fmaf(bf16_lo(w.w), xb.z, t); t = fmaf(bf16_hi(w.w), xb.w, t);
                        acc[i] = t;
                    }
                }
            }
            const float v = warp_reduce8(acc, lane);
            if (g0 + 8 >= R) {                                      // last group of these slots: every weight has been consumed
                __syncwarp();
                if (lane == 0) { mbar_arrive(&ring.empty[slotA]); if (pair) mbar_arrive(&ring.empty[slotB]); }
            }
            float* pb = part + (grp & 1) * 64;
            if ((lane & 3) == 0) pb[warp * 8 + (lane >> 2)] = v;
            cons_sync();                                            // (also: every read of xs by this group is done)
            if (tid < ng) {
                float t = pb[tid];
#pragma unroll
                for (int w8 = 1; w8 < NCONS_WARPS; ++w8) t += pb[w8 * 8 + tid];
                const int row = r + g0 + tid;
                const float nv = xres[row - s.r0] + t; xres[row - s.r0] = nv;
                ll_store(out + row, nv, tag);
            }
        }
        r += R; q += pair ? 2 : 1;
    }
}

// K <= 1024 GEMVs (qkv, gate/up, lm_head of the 0.6B dims): four rows per warp and turn, two ring slots (32 rows) per turn.
template <int K, int EPI>
__device__ __forceinline__ void consume_quad(const Slice& s, const Ring& ring, uint32_t& q, const float* xs, uint2* out,
                                             uint32_t tag, float& best_v, int& best_i,
                                             const float* norm_w, float norm_r, long long* fine) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int fi = 0;
#define CF() do { if (fine && threadIdx.x == 0 && fi < 24) fine[fi++] = clock64(); } while (0)
    CF();
    constexpr int NU = K / 8;
    float xr[K / 32];
    if (norm_w) load_xr_norm<K>(xs, norm_w, norm_r, xr, lane);
    else load_xr<K>(xs, xr, lane);
    cons_sync();                                   // every warp holds its copy: xs may be overwritten from here on
    CF();
    const int j = lane >> 3;                       // row of the quad whose total this lane holds after row_dot4
    for (int r = s.r0; r < s.r1;) {
        const int rowsA = min(s.rpc, s.r1 - r);
        const uint32_t slotA = q % ring.nslot;
        mbar_wait(&ring.full[slotA], (q / ring.nslot) & 1);
        const bool pair = s.rpc < 32 && r + rowsA < s.r1;
        const int rowsB = pair ? min(s.rpc, s.r1 - r - rowsA) : 0;
        const uint32_t slotB = (q + 1) % ring.nslot;
        if (pair) mbar_wait(&ring.full[slotB], ((q + 1) / ring.nslot) & 1);
        CF();
        const uint4* baseA = reinterpret_cast<const uint4*>(ring.slots + (size_t)slotA * SLOT_BYTES);
        const uint4* baseB = reinterpret_cast<const uint4*>(ring.slots + (size_t)slotB * SLOT_BYTES);
        const int R = rowsA + rowsB;
        for (int i0 = 4 * warp; i0 < R; i0 += 4 * NCONS_WARPS) {
            const int nv = min(4, R - i0);
            const uint4* rp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ri = i0 + (i < nv ? i : 0);
                rp[i] = ri < rowsA ? baseA + (size_t)ri * NU : baseB + (size_t)(ri - rowsA) * NU;
            }
            const float v = row_dot4<K>(rp[0], rp[1], rp[2], rp[3], xr, lane);
            const int row = r + i0 + j;
            if (EPI == ME_SWIGLU) {                // rows (gate, up, gate, up): lanes 0 / 16 hold a gate, lanes 8 / 24 its up row
                const float up = __shfl_xor_sync(0xffffffffu, v, 8);
                if ((lane & 15) == 0 && j < nv) ll_store(out + (row >> 1), silu(v) * up, tag);
            } else if (EPI == ME_STORE) {
                if ((lane & 7) == 0 && j < nv) ll_store(out + row, v, tag);
            } else {                               // ME_ARGMAX: rows arrive in increasing order per lane, strict > keeps the first maximum
                if ((lane & 7) == 0 && j < nv && v > best_v) { best_v = v; best_i = row; }
            }
        }
        CF();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&ring.empty[slotA]); if (pair) mbar_arrive(&ring.empty[slotB]); }
        r += R; q += pair ? 2 : 1;
    }
    CF();
#undef CF
}

// consumer: process all chunks of a slice.  `xs` holds the (already normalised) activation vector.
// Results are published as tagged words to `out` (ME_STORE / ME_SWIGLU), added to the CTA-local
// residual rows `xres` and published (ME_RESID), or folded into the running argmax (ME_ARGMAX).
template <int K, int EPI>
__device__ __forceinline__ void consume(const Slice& s, const Ring& ring, uint32_t& q, const float* xs, uint2* out,
                                        uint32_t tag, float* xres, float& best_v, int& best_i,
                                        const float* norm_w = nullptr, float norm_r = 1.f, long long* fine = nullptr,
                                        float* part = nullptr) {
    if constexpr (EPI == ME_RESID) { consume_ksplit<K>(s, ring, q, xs, out, tag, xres, part); return; }
    else if constexpr (K <= 1024) { consume_quad<K, EPI>(s, ring, q, xs, out, tag, best_v, best_i, norm_w, norm_r, fine); return; }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int fi = 0;
#define CF() do { if (fine && threadIdx.x == 0 && fi < 24) fine[fi++] = clock64(); } while (0)
    CF();
    constexpr int RSTEP = (EPI == ME_SWIGLU) ? 2 : 1;
    constexpr bool XREG = K <= 2048;               // long-K slices (down_proj, 7 rows per CTA) read x from smem instead
    constexpr bool DUAL = XREG && K <= 1024;       // two rows per warp and turn: a 16-row slot is one pass of the 8 warps
    constexpr int KX = XREG ? K : 256;             // dummy instantiation size when x stays in smem
    constexpr int UPW = (DUAL && EPI != ME_SWIGLU) ? 2 : 1;       // units (rows or pairs) a warp takes per turn
    const int sub = lane >> 4;
    float xr[XREG ? K / 32 : 1];
    if (XREG) {
        if (norm_w) load_xr_norm<KX>(xs, norm_w, norm_r, reinterpret_cast<float (&)[KX / 32]>(xr), lane);
        else load_xr<KX>(xs, reinterpret_cast<float (&)[KX / 32]>(xr), lane);
        cons_sync();                               // every warp holds its copy: xs may be overwritten from here on
    }
    CF();
    int unit = 0;                                  // unit index within this CTA's slice (kept a multiple of UPW per slot)
    for (int r = s.r0; r < s.r1; r += s.rpc, ++q) {
        const int rows = min(s.rpc, s.r1 - r);
        const uint32_t slot = q % ring.nslot, par = (q / ring.nslot) & 1;
        mbar_wait(&ring.full[slot], par);
        CF();
        const uint4* base = reinterpret_cast<const uint4*>(ring.slots + (size_t)slot * SLOT_BYTES);
        const int units_here = rows / RSTEP;
        // units are dealt round-robin to warps across the whole slice (UPW consecutive units per warp and turn)
        const int first = (warp - ((unit / UPW) % NCONS_WARPS) + NCONS_WARPS) % NCONS_WARPS;
        for (int ub = first * UPW; ub < units_here; ub += NCONS_WARPS * UPW) {
            if (EPI == ME_SWIGLU) {
                const int row = r + ub * 2;        // gate row; up row = row + 1
                float v0, v1;
                if (DUAL) {
                    const float v = row_dot2<KX>(base + (size_t)(ub * 2) * (K / 8), base + (size_t)(ub * 2 + 1) * (K / 8),
                                                 reinterpret_cast<const float (&)[KX / 32]>(xr), lane);
                    v0 = v; v1 = __shfl_xor_sync(0xffffffffu, v, 16);      // valid on the lower half (lane 0 publishes)
                } else if (XREG) {
                    v0 = row_dot<KX>(base + (size_t)(ub * 2) * (K / 8), reinterpret_cast<const float (&)[KX / 32]>(xr), lane);
                    v1 = row_dot<KX>(base + (size_t)(ub * 2 + 1) * (K / 8), reinterpret_cast<const float (&)[KX / 32]>(xr), lane);
                } else {
                    v0 = row_dot_smem<K>(base + (size_t)(ub * 2) * (K / 8), xs, lane);
                    v1 = row_dot_smem<K>(base + (size_t)(ub * 2 + 1) * (K / 8), xs, lane);
                }
                if (lane == 0) ll_store(out + (row >> 1), silu(v0) * v1, tag);
            } else {
                bool act; int row; float v0;
                if (DUAL) {
                    const bool two = ub + 1 < units_here;                 // the slot's last turn may hold a single row
                    v0 = row_dot2<KX>(base + (size_t)ub * (K / 8), base + (size_t)(two ? ub + 1 : ub) * (K / 8),
                                      reinterpret_cast<const float (&)[KX / 32]>(xr), lane);
                    act = ((lane & 15) == 0) && (sub == 0 || two);
                    row = r + ub + sub;
                } else {
                    if (XREG) v0 = row_dot<KX>(base + (size_t)ub * (K / 8), reinterpret_cast<const float (&)[KX / 32]>(xr), lane);
                    else v0 = row_dot_smem<K>(base + (size_t)ub * (K / 8), xs, lane);
                    act = lane == 0; row = r + ub;
                }
                if (EPI == ME_STORE) {
                    if (act) ll_store(out + row, v0, tag);
                } else if (EPI == ME_RESID) {
                    if (act) { const float nv = xres[row - s.r0] + v0; xres[row - s.r0] = nv; ll_store(out + row, nv, tag); }
                } else {
                    if (act && v0 > best_v) { best_v = v0; best_i = row; }
                }
            }
        }
        unit += (units_here + UPW - 1) / UPW * UPW;
        CF();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[slot]);
    }
    CF();
#undef CF
    if (!XREG) cons_sync();                        // xs was read in place: nobody may overwrite it before this point
}

// RMSNorm scale of the vector sitting in xs from the per-thread partial sums of squares; the scaling itself is
// fused into the register load of the GEMV (load_xr_norm).  Ends with a barrier: xs is complete for every warp.
__device__ __forceinline__ float norm_scale(float ss, int n, float eps, float* red) {
    const int tid = threadIdx.x;
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    cons_sync();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NCONS_WARPS; ++i) tot += red[i];
    return 1.0f / sqrtf(tot / n + eps);
}

// per-head RMSNorm + RoPE of one 128-vector by one warp (lane holds d = lane, +32, +64, +96);
// input = tagged words
__device__ __forceinline__ void head_norm_rope(const uint2* __restrict__ src, uint32_t tag, const float* __restrict__ nw,
                                               float eps, const float* __restrict__ cs, const float* __restrict__ sn,
                                               float* dst, int lane) {
    float v[4];
    ll_poll4(src + lane, 32, tag, v);
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

#define MEGA_FINE(k)                                                                                    \
    do {                                                                                               \
        if (dbg_row && tid == 0 && l == 5) dbg_row[400 + (k)] = clock64();                              \
    } while (0)
// every CTA: wall-clock (globaltimer, ns) of three events of layer 5 -> row 0, slots [512 + 3 * cta ...)
#define MEGA_GT(k)                                                                                     \
    do {                                                                                               \
        if (p.dbg && tid == 0 && l == 5) {                                                             \
            unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));               \
            p.dbg[512 + 3 * blockIdx.x + (k)] = (long long)t_;                                         \
        }                                                                                              \
    } while (0)
#define MEGA_MARK()                                                                                    \
    do {                                                                                               \
        if (dbg_row && tid == 0 && dbg_i < DBG_SLOTS) dbg_row[dbg_i++] = clock64();                    \
    } while (0)

template <int H, int QD, int I, int NS>
__global__ void __launch_bounds__(NTHREADS, 1) decode_step_kernel(const Params p) {
    constexpr int XS_FLOATS = (I > XS_MIN ? I : XS_MIN) + 64;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int PARAM_FLOATS = 2 * H + 2 * HD;      // per-layer small vectors: ln_in[H], ln_post[H], q_norm[128], k_norm[128]
    Ring ring;
    ring.slots = smem;
    ring.nslot = NS;
    uint8_t* kv_smem = smem + (size_t)NS * SLOT_BYTES;              // [K tile | V tile]
    float* xs = reinterpret_cast<float*>(kv_smem + 2 * KV_TILE_BYTES);
    float* xres = xs + XS_FLOATS;                                      // [XRES_MAX] residual rows owned by this CTA
    float* pbuf = xres + XRES_MAX;                                     // [2][PARAM_FLOATS] per-layer small vectors (double buffer)
    float* ropes = pbuf + 2 * PARAM_FLOATS;                            // [128] cos | sin of this step's position
    DecLayerW* ltab = reinterpret_cast<DecLayerW*>(ropes + 128);       // [MAX_LAYERS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ltab + MAX_LAYERS);
    ring.full = bars; ring.empty = bars + NSLOT_MAX;
    uint64_t* kv_full = bars + 2 * NSLOT_MAX; uint64_t* kv_empty = kv_full + 1;
    uint64_t* p_full = kv_empty + 1; uint64_t* p_empty = p_full + 2;   // [2] each
    float* red = reinterpret_cast<float*>(bars + 2 * NSLOT_MAX + 6);      // [64]
    int* ired = reinterpret_cast<int*>(red + 64);                      // [64]
    float* part = reinterpret_cast<float*>(ired + 64);                 // [2][8 warps][8 rows] K-split partials (consume_ksplit)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_producer = warp == NCONS_WARPS;

    if (__ldcg(p.done) != 0) return;            // sequence finished: nothing to do this step

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&ring.full[i], 1); mbar_init(&ring.empty[i], NCONS_WARPS); }
        mbar_init(kv_full, 1); mbar_init(kv_empty, NCONS_WARPS);
        for (int i = 0; i < 2; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], NCONS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Small read-only tables go to shared memory: with 1.2 GB streaming through L2 every step they are never
    // L2-resident, and a dependent DRAM round trip (~1-2 us) per use is what the phases cannot afford.
    const int pos = __ldcg(p.pos);
    {
        const uint2* src = reinterpret_cast<const uint2*>(p.layers);
        uint2* dst = reinterpret_cast<uint2*>(ltab);
        for (int i = tid; i < p.L * (int)(sizeof(DecLayerW) / 8); i += NTHREADS) dst[i] = src[i];
        if (tid < 64) ropes[tid] = p.rope_cos[(size_t)pos * 64 + tid];
        else if (tid < 128) ropes[tid] = p.rope_sin[(size_t)pos * 64 + tid - 64];
    }
    __syncthreads();

    // attention work item of this CTA: (kv head att_g, keys [att_j0, att_j0 + KV_KEYS)); keys < pos are
    // already in the cache (n_old of them fall in this split), key `pos` is produced in this step.
    const bool att_cta = (int)blockIdx.x < p.nkv * p.nsplit;
    const int att_g = blockIdx.x / p.nsplit, att_sp = blockIdx.x % p.nsplit, att_j0 = att_sp * KV_KEYS;
    const int n_old = att_cta ? max(0, min(pos - att_j0, KV_KEYS)) : 0;
    uint32_t kvq = 0;
    uint32_t q = 0;
    if (is_producer) {
        if (lane == 0) {
            ChunkCursor<H, QD, I> pf;
            pf.init(p, ltab);
            {   // start the HBM stream immediately: the first PF_AHEAD chunks go to L2 now
                const bf16* psrc; uint32_t pbytes;
                for (int i = 0; i < p.pf_ahead; ++i) if (pf.next(p, psrc, pbytes)) l2_prefetch(psrc, pbytes);
            }
            const size_t kv_row = ((size_t)att_g * p.max_ctx + att_j0) * HD;
            if (n_old > 0) {   // K/V tiles of the first two layers
                l2_prefetch(p.kcache + kv_row, (uint32_t)n_old * HD * 4);
                l2_prefetch(p.vcache + kv_row, (uint32_t)n_old * HD * 4);
            }
            for (int l = 0; l <= p.L; ++l) {
                {   // small per-layer vectors -> pbuf[l & 1] (layer L = final norm only)
                    float* pb = pbuf + (l & 1) * PARAM_FLOATS;
                    mbar_wait(&p_empty[l & 1], ((l >> 1) & 1) ^ 1);
                    if (l < p.L) {
                        const DecLayerW w = ltab[l];
                        mbar_expect_tx(&p_full[l & 1], (uint32_t)(2 * H + 2 * HD) * 4);
                        bulk_g2s(pb, w.ln_in, H * 4, &p_full[l & 1]);
                        bulk_g2s(pb + H, w.ln_post, H * 4, &p_full[l & 1]);
                        bulk_g2s(pb + 2 * H, w.qnorm, HD * 4, &p_full[l & 1]);
                        bulk_g2s(pb + 2 * H + HD, w.knorm, HD * 4, &p_full[l & 1]);
                    } else {
                        mbar_expect_tx(&p_full[l & 1], (uint32_t)H * 4);
                        bulk_g2s(pb, p.final_norm, H * 4, &p_full[l & 1]);
                        break;
                    }
                }
                const DecLayerW w = ltab[l];
                produce(make_slice(w.wqkv, QD + 2 * p.KVD, H, 1), ring, q, pf, p);
                if (n_old > 0) {   // K/V rows of earlier positions do not depend on this step: prefetch them too
                    const uint32_t bytes = (uint32_t)n_old * HD * 4;
                    if (l + 1 < p.L) {
                        l2_prefetch(p.kcache + (size_t)(l + 1) * p.cache_layer_stride + kv_row, bytes);
                        l2_prefetch(p.vcache + (size_t)(l + 1) * p.cache_layer_stride + kv_row, bytes);
                    }
                    mbar_wait(kv_empty, (kvq & 1) ^ 1);
                    mbar_expect_tx(kv_full, 2 * bytes);
                    const size_t off = (size_t)l * p.cache_layer_stride + kv_row;
                    bulk_g2s(kv_smem, p.kcache + off, bytes, kv_full);
                    bulk_g2s(kv_smem + KV_TILE_BYTES, p.vcache + off, bytes, kv_full);
                    ++kvq;
                }
                produce(make_slice(w.wo, H, QD, 1), ring, q, pf, p);
                produce(make_slice(w.wgu, 2 * I, H, 2), ring, q, pf, p);
                produce(make_slice(w.wdown, H, I, 1), ring, q, pf, p);
            }
            produce(make_slice(p.lm_head, p.V, H, 1), ring, q, pf, p);
            {   // warm L2 with the head of the NEXT step's stream (same addresses every step)
                ChunkCursor<H, QD, I> nx;
                nx.init(p, ltab);
                const bf16* psrc; uint32_t pbytes;
                for (int i = 0; i < p.pf_ahead; ++i) if (nx.next(p, psrc, pbytes)) l2_prefetch(psrc, pbytes);
            }
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    long long* dbg_row = nullptr; int dbg_i = 0;
    if (p.dbg && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) dbg_row = p.dbg + (blockIdx.x == 0 ? 0 : DBG_SLOTS);
    MEGA_MARK();
    const int half = HD / 2;
    const float* cs = ropes;
    const float* sn = ropes + half;
    const unsigned G = gridDim.x;
    float best_v = -INFINITY; int best_i = 0x7fffffff;
    // tag = launch epoch (unique per executed step, survives new utterances that revisit the same positions)
    const unsigned epoch = __ldcg(p.bar + 1);
    const uint32_t tag_base = (epoch & 0xffffffu) << 8;
    // residual rows owned by this CTA (same row partition for o_proj and down_proj)
    const Slice xsl = make_slice(nullptr, H, QD, 1);
    // slice geometry does not depend on the layer: computed once, only the weight pointer changes
    Slice sl_qkv = make_slice(nullptr, QD + 2 * p.KVD, H, 1), sl_o = make_slice(nullptr, H, QD, 1),
          sl_gu = make_slice(nullptr, 2 * I, H, 2), sl_dn = make_slice(nullptr, H, I, 1);
    for (int i = tid; i < xsl.r1 - xsl.r0; i += NCONS) xres[i] = __ldcg(p.x + xsl.r0 + i);

    for (int l = 0; l < p.L; ++l) {
        const DecLayerW w = ltab[l];
        const uint32_t tl = tag_base | ((uint32_t)l << 3);
        const float* pb = pbuf + (l & 1) * PARAM_FLOATS;               // ln_in | ln_post | q_norm | k_norm of this layer
        mbar_wait(&p_full[l & 1], (l >> 1) & 1);
        // ---- phase 1: RMSNorm + [q|k|v] GEMV ----
        float nr;
        {
            float ss = 0.f;
            if (l == 0) { for (int i = tid; i < H; i += NCONS) { const float v = __ldcg(p.x + i); xs[xs_swz(i)] = v; ss = fmaf(v, v, ss); } }
            else { MEGA_FINE(0); MEGA_FINE(1); ss = ll_gather(p.x_ll, H, (tag_base | ((uint32_t)(l - 1) << 3)) | PH_XD, xs); MEGA_FINE(2); }
            nr = norm_scale(ss, H, p.eps, red);
            MEGA_FINE(3);
        }
        consume<H, ME_STORE>(sl_qkv, ring, q, xs, p.qkv_ll, tl | PH_QKV, xres, best_v, best_i, pb, nr);
        MEGA_FINE(4);
        MEGA_GT(0);
        MEGA_MARK();
        // ---- phase 2: attention partials, work item = (kv head, 64-key split) ----
        {
            // Splits only hold keys that were cached before this step (n_old of them, prefetched by the producer, so a
            // split depends on nothing but q); the key/value of the current token is folded in as one more partial by the
            // merging CTA (split 0 of the kv head), which also appends it to the cache.
            const int nloc = n_old;
            const int nact = min(p.nsplit, (pos + KV_KEYS - 1) / KV_KEYS);      // splits holding at least one cached key
            const bool merger = att_cta && att_sp == 0;                         // pos >= 1: split 0 always has cached keys
            if (nloc > 0) {
                const int g = att_g;
                float* qs = xs;                       // [group][128]
                float* kn = qs + p.group * HD;        // [128]
                float* vn = kn + HD;                  // [128]
                float* sc = vn + HD;                  // [group][KV_KEYS]
                float* ml = sc + p.group * KV_KEYS;   // [group][2] (max, sum)   (generic path) / score of the new key per head
                float* osum = ml + 8;                 // [warps][2][128] per-warp partial outputs (group == 2 path)
                float* wml = osum + NCONS_WARPS * 2 * HD;   // [warps][2][2] per-warp (max, sum)
                float* snew = wml + NCONS_WARPS * 4;  // [group] score of the current token's key per head (merging CTA)
                float* Ks = reinterpret_cast<float*>(kv_smem);
                float* Vs = reinterpret_cast<float*>(kv_smem + KV_TILE_BYTES);
                MEGA_FINE(24);
                cons_sync();                          // xs (phase-1 activations) no longer needed by any warp; q/k/v words are
                                                      // polled directly below (few readers per word)
                if (warp < p.group) head_norm_rope(p.qkv_ll + (size_t)(g * p.group + warp) * HD, tl | PH_QKV, pb + 2 * H, p.eps, cs, sn, qs + warp * HD, lane);
                else if (warp == p.group && merger) head_norm_rope(p.qkv_ll + QD + (size_t)g * HD, tl | PH_QKV, pb + 2 * H + HD, p.eps, cs, sn, kn, lane);
                else if (warp == p.group + 1 && merger) {
                    float vv[4];
                    ll_poll4(p.qkv_ll + QD + p.KVD + (size_t)g * HD + lane, 32, tl | PH_QKV, vv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) vn[lane + 32 * i] = vv[i];
                }
                MEGA_FINE(25);
                mbar_wait(kv_full, kvq & 1);          // prefetched K/V tiles have landed
                cons_sync();
                MEGA_FINE(26);
                if (merger) {
                    if (tid < HD) {                   // KV append (replaces Tensor::cat, layers.rs:311-317)
                        float* kc = p.kcache + (size_t)l * p.cache_layer_stride + ((size_t)g * p.max_ctx + pos) * HD;
                        float* vc = p.vcache + (size_t)l * p.cache_layer_stride + ((size_t)g * p.max_ctx + pos) * HD;
                        kc[tid] = kn[tid]; vc[tid] = vn[tid];
                    }
                    if (warp >= NCONS_WARPS - p.group) {   // score of the new key for head hq (last `group` warps)
                        const int hq = warp - (NCONS_WARPS - p.group);
                        const float4 a = *reinterpret_cast<const float4*>(qs + hq * HD + lane * 4);
                        const float4 b = *reinterpret_cast<const float4*>(kn + lane * 4);
                        const float sn_ = warp_sum(fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))));
                        if (lane == 0) snew[hq] = sn_ / sqrtf((float)HD);
                    }
                }
                MEGA_FINE(27);
                if (p.group == 2) {
                    // 2 query heads per kv head: warp w owns keys w, w+8, ... of the split and computes a complete local softmax
                    // partial (max, sum, unnormalised output) for them; the 8 warp partials are merged through shared memory
                    // exactly like the per-split partials are merged later.  One barrier, no score array.
                    //  scores: every lane multiplies its 4 dims of the K row (one conflict-free LDS.128) with both q vectors held
                    //  in registers; the 16 partial sums per lane (8 keys x 2 heads) are reduced across the warp with a
                    //  transposing butterfly (16 shuffles instead of 80); lane 2 * (2 * kk + h) (and its odd twin) ends up
                    //  with the score of key kk, head h
                    const float4 q0 = *reinterpret_cast<const float4*>(qs + lane * 4);
                    const float4 q1 = *reinterpret_cast<const float4*>(qs + HD + lane * 4);
                    float pv[16];
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const int j = warp + 8 * kk;
                        // branch-free: rows past the split's last key are read (stale data) and masked below
                        const float4 kv = *reinterpret_cast<const float4*>(Ks + j * HD + lane * 4);
                        pv[2 * kk] = fmaf(kv.x, q0.x, fmaf(kv.y, q0.y, fmaf(kv.z, q0.z, kv.w * q0.w)));
                        pv[2 * kk + 1] = fmaf(kv.x, q1.x, fmaf(kv.y, q1.y, fmaf(kv.z, q1.z, kv.w * q1.w)));
                    }
#pragma unroll
                    for (int o = 16, n = 16; n > 1; o >>= 1, n >>= 1) {
                        const bool up = lane & o;
#pragma unroll
                        for (int i = 0; i < n / 2; ++i) {
                            const float send = up ? pv[i] : pv[i + n / 2];
                            const float keep = up ? pv[i + n / 2] : pv[i];
                            pv[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    pv[0] += __shfl_xor_sync(0xffffffffu, pv[0], 1);
                    const bool mine = warp + 8 * (lane >> 2) < nloc;               // this lane's key exists
                    const float sv = mine ? pv[0] / sqrtf((float)HD) : -INFINITY;
                    float mw = sv;                                                   // max over this warp's keys, per head (lane bit 1)
                    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, 4));
                    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, 8));
                    mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, 16));
                    const float ev = mine ? expf(sv - mw) : 0.f;
                    float lw = ev;
                    lw += __shfl_xor_sync(0xffffffffu, lw, 4);
                    lw += __shfl_xor_sync(0xffffffffu, lw, 8);
                    lw += __shfl_xor_sync(0xffffffffu, lw, 16);
                    // o_w[h][d] = sum_kk e[kk][h] * V[j][d]: 4 dims per lane (one LDS.128 of V per key), both heads
                    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const int j = warp + 8 * kk;
                        const float4 vv = *reinterpret_cast<const float4*>(Vs + j * HD + lane * 4);
                        const float e0 = __shfl_sync(0xffffffffu, ev, 4 * kk), e1 = __shfl_sync(0xffffffffu, ev, 4 * kk + 2);
                        if (j < nloc) {       // (a stale V row may hold non-finite garbage: 0 * inf must not reach the sum)
                            o0.x = fmaf(e0, vv.x, o0.x); o0.y = fmaf(e0, vv.y, o0.y); o0.z = fmaf(e0, vv.z, o0.z); o0.w = fmaf(e0, vv.w, o0.w);
                            o1.x = fmaf(e1, vv.x, o1.x); o1.y = fmaf(e1, vv.y, o1.y); o1.z = fmaf(e1, vv.z, o1.z); o1.w = fmaf(e1, vv.w, o1.w);
                        }
                    }
                    *reinterpret_cast<float4*>(osum + (warp * 2 + 0) * HD + lane * 4) = o0;
                    *reinterpret_cast<float4*>(osum + (warp * 2 + 1) * HD + lane * 4) = o1;
                    if (lane == 0 || lane == 2) { wml[(warp * 2 + (lane >> 1)) * 2] = mw; wml[(warp * 2 + (lane >> 1)) * 2 + 1] = lw; }
                    cons_sync();
                    MEGA_FINE(28);
                    {
                        const int hq = tid / HD, d = tid - hq * HD;      // NCONS == 2 * HD
                        float M = -INFINITY;
#pragma unroll
                        for (int w8 = 0; w8 < NCONS_WARPS; ++w8) M = fmaxf(M, wml[(w8 * 2 + hq) * 2]);
                        float acc = 0.f, Ls = 0.f;
#pragma unroll
                        for (int w8 = 0; w8 < NCONS_WARPS; ++w8) {
                            const float f = expf(wml[(w8 * 2 + hq) * 2] - M);           // exp(-inf) = 0: warps without keys
                            acc = fmaf(f, osum[(w8 * 2 + hq) * HD + d], acc);
                            Ls = fmaf(f, wml[(w8 * 2 + hq) * 2 + 1], Ls);
                        }
                        uint2* rec = p.part_ll + ((size_t)blockIdx.x * 2 + hq) * PSTRIDE;
                        ll_store(rec + d, acc, tl | PH_PART);
                        if (d < 2) ll_store(rec + HD + d, d == 0 ? M : Ls, tl | PH_PART);
                    }
                    MEGA_FINE(29);
                } else {
                    // generic group size: one thread per (head, key); the d loop is rotated by the key index so that the
                    // 32 lanes of a warp hit 32 different banks of the row-major K tile
                    for (int idx = tid; idx < p.group * KV_KEYS; idx += NCONS) {
                        const int hq = idx / KV_KEYS, j = idx - hq * KV_KEYS;
                        if (j < nloc) {
                            const float* kr = Ks + j * HD; const float* qr = qs + hq * HD;
                            float a0 = 0.f, a1 = 0.f;
#pragma unroll 8
                            for (int dd = 0; dd < HD; dd += 2) {
                                const int d0 = (dd + j) & (HD - 1), d1 = (dd + 1 + j) & (HD - 1);
                                a0 = fmaf(kr[d0], qr[d0], a0); a1 = fmaf(kr[d1], qr[d1], a1);
                            }
                            sc[hq * KV_KEYS + j] = (a0 + a1) / sqrtf((float)HD);
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
                        if (lane == 0) { ml[warp * 2] = mx; ml[warp * 2 + 1] = sum; }
                    }
                    cons_sync();
                    for (int idx = tid; idx < p.group * HD; idx += NCONS) {
                        const int hq = idx / HD, d = idx - hq * HD;
                        float acc = 0.f;
                        for (int j = 0; j < nloc; ++j) acc = fmaf(sc[hq * KV_KEYS + j], Vs[j * HD + d], acc);
                        uint2* rec = p.part_ll + ((size_t)blockIdx.x * p.group + hq) * PSTRIDE;
                        ll_store(rec + d, acc, tl | PH_PART);
                        if (d < 2) ll_store(rec + HD + d, ml[hq * 2 + d], tl | PH_PART);
                    }
                }
                {                                     // hand the K/V staging buffer back to the producer
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(kv_empty);
                    ++kvq;
                }
                MEGA_FINE(30);
                MEGA_GT(1);
                // split 0 of every kv head merges the partials of all active splits and publishes the head outputs
                if (merger) {
                    for (int idx = tid; idx < p.group * HD; idx += NCONS) {
                        const int hq = idx / HD, d = idx - hq * HD;          // hq is uniform per warp (HD = 4 warps)
                        // lane s of every warp fetches (max, sum) of split s (nact <= MAX_SPLITS <= 32 lanes); all loads of a
                        // round are issued before any tag is examined (one round trip when ready)
                        const uint32_t tg = tl | PH_PART;
                        constexpr int RB = 9;                                // partial outputs fetched per round (registers)
                        uint2 mv, lv, ov[RB];
                        bool ok;
                        auto load_round = [&](int u0) {
#pragma unroll
                            for (int u = 0; u < RB; ++u) {
                                if (u0 + u < nact) {
                                    const uint2* rec = p.part_ll + ((size_t)(g * p.nsplit + u0 + u) * p.group + hq) * PSTRIDE;
                                    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(ov[u].x), "=r"(ov[u].y) : "l"(rec + d) : "memory");
                                }
                            }
                        };
                        auto round_ok = [&](int u0) {
                            bool k = true;
#pragma unroll
                            for (int u = 0; u < RB; ++u) if (u0 + u < nact) k = k && (ov[u].y == tg);
                            return k;
                        };
                        do {        // first round: (max, sum) of every split + the first RB partial outputs, one round trip
                            mv.y = tg; lv.y = tg; mv.x = 0u; lv.x = 0u;
                            if (lane < nact) {
                                const uint2* rec = p.part_ll + ((size_t)(g * p.nsplit + lane) * p.group + hq) * PSTRIDE;
                                asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(mv.x), "=r"(mv.y) : "l"(rec + HD) : "memory");
                                asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(lv.x), "=r"(lv.y) : "l"(rec + HD + 1) : "memory");
                            }
                            load_round(0);
                            ok = __all_sync(0xffffffffu, (mv.y == tg) && (lv.y == tg) && round_ok(0));
                        } while (!ok);
                        MEGA_FINE(35);
                        // softmax merge, one partial per lane: lanes < nact hold a split, lane nact the current token's key
                        // (max = its score, sum = 1, output = its value row)
                        const float m_l = lane < nact ? __uint_as_float(mv.x) : (lane == nact ? snew[hq] : -INFINITY);
                        const float l_l = lane < nact ? __uint_as_float(lv.x) : (lane == nact ? 1.f : 0.f);
                        const float M = warp_max(m_l);
                        const float f = expf(m_l - M);                       // exp(-inf) = 0 on idle lanes
                        const float Lsum = warp_sum(f * l_l);
                        float O = __shfl_sync(0xffffffffu, f, nact) * vn[d];
                        for (int u0 = 0; u0 < nact; u0 += RB) {              // contexts beyond 9 splits: one more round trip each
                            if (u0 > 0) { do { load_round(u0); ok = __all_sync(0xffffffffu, round_ok(u0)); } while (!ok); }
#pragma unroll
                            for (int u = 0; u < RB; ++u)
                                if (u0 + u < nact) O = fmaf(__shfl_sync(0xffffffffu, f, u0 + u), __uint_as_float(ov[u].x), O);
                        }
                        ll_store(p.attn_ll + (size_t)(g * p.group + hq) * HD + d, O / Lsum, tl | PH_ATTN);
                    }
                }
                MEGA_FINE(31);
                cons_sync();                          // attention scratch (aliases xs) is free again
            }
        }
        MEGA_MARK();
        // ---- phase 3: o_proj GEMV + residual ----
        MEGA_FINE(32);
        ll_gather(p.attn_ll, QD, tl | PH_ATTN, xs);
        MEGA_FINE(33);
        MEGA_GT(2);
        cons_sync();
        consume<QD, ME_RESID>(sl_o, ring, q, xs, p.x_ll, tl | PH_XO, xres, best_v, best_i, nullptr, 1.f, nullptr, part);
        MEGA_FINE(34);
        MEGA_MARK();
        // ---- phase 4: RMSNorm + gate/up GEMV + SiLU*mul ----
        cons_sync();
        {
            MEGA_FINE(8); MEGA_FINE(9);
            const float ss = ll_gather(p.x_ll, H, tl | PH_XO, xs); MEGA_FINE(10);
            nr = norm_scale(ss, H, p.eps, red); MEGA_FINE(11);
        }
        consume<H, ME_SWIGLU>(sl_gu, ring, q, xs, p.act_ll, tl | PH_ACT, xres, best_v, best_i, pb + H, nr, (dbg_row && l == 5) ? dbg_row + 440 : nullptr);
        MEGA_FINE(12);
        MEGA_FINE(13);
        MEGA_MARK();
        // ---- phase 5: down GEMV + residual ----
        MEGA_FINE(16); MEGA_FINE(17);
        ll_gather(p.act_ll, I, tl | PH_ACT, xs); MEGA_FINE(18);
        cons_sync();
        consume<I, ME_RESID>(sl_dn, ring, q, xs, p.x_ll, tl | PH_XD, xres, best_v, best_i, nullptr, 1.f, nullptr, part);
        MEGA_FINE(19);
        MEGA_FINE(20);
        MEGA_MARK();
        cons_sync();
        if (lane == 0) mbar_arrive(&p_empty[l & 1]);                   // this layer's parameter buffer may be refilled
    }
    // ---- final RMSNorm + tied lm_head GEMV + argmax ----
    float nrf;
    {
        const float ss = ll_gather(p.x_ll, H, (tag_base | ((uint32_t)(p.L - 1) << 3)) | PH_XD, xs);
        mbar_wait(&p_full[p.L & 1], (p.L >> 1) & 1);
        nrf = norm_scale(ss, H, p.eps, red);
    }
    consume<H, ME_ARGMAX>(make_slice(p.lm_head, p.V, H, 1), ring, q, xs, nullptr, 0u, xres, best_v, best_i,
                          pbuf + (p.L & 1) * PARAM_FLOATS, nrf);
    MEGA_MARK();
    // candidates live in lanes 0, 8, 16, 24 of every warp (the four rows of a turn; lanes 0 / 16 in the two-row form):
    // merge them, lane 0 publishes the warp's best
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best_v, o); const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    cons_sync();
    if (lane == 0) { red[warp] = best_v; ired[warp] = best_i; }
    cons_sync();
    int& is_last = ired[63];
    if (tid == 0) {
        float v = -INFINITY; int idx = 0x7fffffff;
        for (int wq = 0; wq < NCONS_WARPS; ++wq)
            if (red[wq] > v || (red[wq] == v && ired[wq] < idx)) { v = red[wq]; idx = ired[wq]; }
        p.part_val[blockIdx.x] = v; p.part_idx[blockIdx.x] = idx;
        __threadfence();
        unsigned t = atomicAdd(p.bar, 1u);
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
        int& tok_s = ired[62];
        if (tid == 0) {
            for (int wq = 1; wq < NCONS_WARPS; ++wq)
                if (red[wq] > v || (red[wq] == v && ired[wq] < idx)) { v = red[wq]; idx = ired[wq]; }
            int tok = idx;
            const int n = *p.n_out;
            if (tok == 151643 || tok == 151645 || n >= p.max_new) { *p.done = 1; *p.next_id = -1; tok = -1; }
            else { p.ids_out[n] = tok; *p.n_out = n + 1; *p.pos = pos + 1; *p.next_id = tok; }
            tok_s = tok;
            p.bar[0] = 0;                        // every CTA has taken its ticket: reset for the next launch
            p.bar[1] = p.bar[1] + 1;             // new epoch: words published by this step can never match again
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
static long long* g_last_dbg = nullptr;   // debug only (ASRB_MEGA_DEBUG): timeline buffer of the last launch

static int mega_xs_floats(int I) { return std::max(I, mega::XS_MIN) + 64; }
static size_t mega_smem_bytes(int H, int I, int nslot) {
    return (size_t)nslot * mega::SLOT_BYTES + 2 * mega::KV_TILE_BYTES +
           (mega_xs_floats(I) + mega::XRES_MAX + 2 * (2 * H + 2 * mega::HD) + 128) * 4 + mega::MAX_LAYERS * sizeof(DecLayerW) +
           (2 * mega::NSLOT_MAX + 6) * 8 + 64 * 4 + 64 * 4 + 128 * 4 + 64;
}
// instantiations: (hidden, q_dim, intermediate) -> ring depth
static int mega_nslot(const asrb_dims& c) { return c.hidden_size > 1024 ? 3 : 4; }

template <int H, int QD, int I> static bool dims_match(const asrb_dims& c) {
    return c.hidden_size == H && c.num_attention_heads * c.head_dim == QD && c.intermediate_size == I;
}

// `ctx` = number of keys the step may attend to (position + 1 upper bound), NOT the cache capacity
bool decode_mega_supported(const Model& m, int B, int ctx) {
    const int max_ctx = ctx;
    const asrb_dims& c = m.d.c;
    if (B < 1 || c.head_dim != 128) return false;      // batch > 1: one fused launch per sequence, back to back
    const int group = c.num_attention_heads / c.num_key_value_heads;
    if (group + 2 > mega::NCONS_WARPS) return false;
    if ((size_t)(group * 128 + 256 + group * mega::KV_KEYS + 8 + mega::NCONS_WARPS * 2 * 128 + mega::NCONS_WARPS * 4 + 8) > (size_t)mega_xs_floats(c.intermediate_size)) return false;
    if (m.ctx->smem_optin < mega_smem_bytes(c.hidden_size, c.intermediate_size, mega_nslot(c))) return false;
    if ((c.hidden_size + m.ctx->sm_count - 1) / m.ctx->sm_count + 1 > mega::XRES_MAX) return false;
    if (c.num_hidden_layers > 32) return false;                  // 5-bit layer field
    if ((max_ctx + mega::KV_KEYS - 1) / mega::KV_KEYS > mega::MAX_SPLITS) return false;                           // merge loop bound (SB)
    if (((max_ctx + mega::KV_KEYS - 1) / mega::KV_KEYS) * c.num_key_value_heads > m.ctx->sm_count) return false;   // one CTA per (kv head, 64-key split)
    return dims_match<1024, 2048, 3072>(c) || dims_match<2048, 2048, 6144>(c) || dims_match<256, 512, 512>(c);
}

// floats of session scratch the fused step needs: tagged exchange buffers (2 floats per value)
size_t decode_mega_part_floats(const Model& m) {
    const asrb_dims& c = m.d.c;
    const int group = c.num_attention_heads / c.num_key_value_heads;
    const size_t words = (size_t)m.d.qkv_dim + (size_t)m.ctx->sm_count * group * mega::PSTRIDE + m.d.q_dim + c.hidden_size +
                         c.intermediate_size + 64;
    return 2 * words + 64;
}

void launch_decode_step_mega(const Model& m, const DecodeBufs& b, int B, float* kcache, float* vcache,
                             size_t cache_layer_stride, size_t cache_seq_stride, int max_ctx, int ctx_now, const MegaBufs& mb,
                             cudaStream_t st, int64_t* launches) {
    ASRB_REQUIRE(decode_mega_supported(m, B, ctx_now), ASRB_ERR_STATE, "fused decode step not supported for this model/batch/context");
    ASRB_REQUIRE(m.d_dec_layers && mb.bar && mb.part, ASRB_ERR_STATE, "fused decode step buffers missing");
    const asrb_dims& c = m.d.c;
    const int G = m.ctx->sm_count;
    const int group = c.num_attention_heads / c.num_key_value_heads;
    // split count is fixed per session (buffer layout); splits beyond the current context are simply empty
    const int nsplit = std::min(mega::MAX_SPLITS, std::min(G / c.num_key_value_heads, (max_ctx + mega::KV_KEYS - 1) / mega::KV_KEYS));
    const size_t smem = mega_smem_bytes(c.hidden_size, c.intermediate_size, mega_nslot(c));
    const void* fn = nullptr;
    if (dims_match<1024, 2048, 3072>(c)) fn = (const void*)mega::decode_step_kernel<1024, 2048, 3072, 4>;          // Qwen3-ASR-0.6B
    else if (dims_match<2048, 2048, 6144>(c)) fn = (const void*)mega::decode_step_kernel<2048, 2048, 6144, 3>;     // Qwen3-ASR-1.7B
    else fn = (const void*)mega::decode_step_kernel<256, 512, 512, 4>;                                             // test config
    ASRB_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // The kernel handles one sequence.  A batch runs as B launches on the stream (weights are re-streamed per sequence:
    // 2.0 k tokens/s at any batch size, still ~1.8x the per-phase path at batch 8); a sequence that has finished
    // returns at once.  Exchange buffers are shared: launches are serialised by the stream and tagged by epoch.
    for (int sb = 0; sb < B; ++sb) {
        mega::Params p{};
        p.layers = m.d_dec_layers; p.lm_head = m.lm_head; p.embed = m.embed; p.final_norm = m.final_norm_sw;
        p.rope_cos = m.rope_cos; p.rope_sin = m.rope_sin; p.eps = (float)c.rms_norm_eps;
        p.L = c.num_hidden_layers; p.H = c.hidden_size; p.QD = m.d.q_dim; p.KVD = m.d.kv_dim; p.I = c.intermediate_size;
        p.V = c.vocab_size; p.nq = c.num_attention_heads; p.nkv = c.num_key_value_heads; p.group = group;
        p.x = b.x + (size_t)sb * c.hidden_size;
        p.kcache = kcache + (size_t)sb * cache_seq_stride; p.vcache = vcache + (size_t)sb * cache_seq_stride;
        p.cache_layer_stride = cache_layer_stride; p.max_ctx = max_ctx; p.nsplit = nsplit;
        p.part_val = b.part_val; p.part_idx = b.part_idx;
        p.pos = b.pos + sb; p.done = b.done + sb; p.next_id = b.next_id + sb;
        p.ids_out = b.ids_out + (size_t)sb * b.max_new; p.n_out = b.n_out + sb; p.max_new = b.max_new;
        p.bar = mb.bar;
        uint2* w = reinterpret_cast<uint2*>(mb.part);            // 16-byte aligned sub-buffers (even word counts)
        p.qkv_ll = w; w += m.d.qkv_dim;
        p.part_ll = w; w += (size_t)G * group * mega::PSTRIDE;
        p.attn_ll = w; w += m.d.q_dim;
        p.x_ll = w; w += c.hidden_size;
        p.act_ll = w; w += c.intermediate_size;
        p.dbg = mb.dbg;
        g_last_dbg = mb.dbg;
        static const int pf_env = [] { const char* e = getenv("ASRB_MEGA_PF"); return e ? atoi(e) : mega::PF_AHEAD; }();
        p.pf_ahead = pf_env;
        // tags must stay monotonic for red.max publication: long before the 24-bit epoch wraps, wipe the exchange buffers
        if (mb.steps_issued && ++*mb.steps_issued >= 0xFFFF00u) {
            ASRB_CUDA_CHECK(cudaMemsetAsync(mb.part, 0, mb.part_bytes, st));
            const unsigned one = 1;
            ASRB_CUDA_CHECK(cudaMemcpyAsync(mb.bar + 1, &one, sizeof(one), cudaMemcpyHostToDevice, st));
            *mb.steps_issued = 1;
        }
        void* args[] = {(void*)&p};
        ASRB_CUDA_CHECK(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(mega::NTHREADS), args, smem, st));
        if (launches) *launches += 1;
    }
}

// debug: copy the clock64 timeline of the most recent fused step (CTA 0 then CTA G-1), returns slots per CTA
int decode_mega_debug_timeline(long long* out, int cap) {
    if (!g_last_dbg || cap < 2 * mega::DBG_SLOTS) return 0;
    cudaDeviceSynchronize();
    cudaMemcpy(out, g_last_dbg, 2 * mega::DBG_SLOTS * sizeof(long long), cudaMemcpyDeviceToHost);
    return mega::DBG_SLOTS;
}
int decode_mega_dbg_slots() { return 4 * mega::DBG_SLOTS; }   // [0, 2048): two CTA timelines; [2048, 4096): per-CTA wall-clock table (decode_batch.cu)

}  // namespace asrb
