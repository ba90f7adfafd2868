// decode_batch.cu -- one greedy decode iteration (inference.rs:160-200) for NB INDEPENDENT sequences as ONE
// persistent kernel: every weight byte is streamed from HBM once per step and contracted against the activation
// vectors of all NB sequences (the reference runs the same loop body once per utterance, src/inference.rs:89).
//
// Same machinery as the single-sequence step (decode_mega.cu): one CTA per SM, a producer warp that streams this
// CTA's row slice of every weight matrix through a shared-memory ring with cp.async.bulk + mbarrier, 8 consumer
// warps, {value, tag} words published with fire-and-forget red.max and polled by the consumers (no grid barriers).
// What changes with a batch:
//   * activations live in shared memory as xs[NB][H] (RMSNorm applied once when the vector is gathered); a warp
//     "unit" contracts 4 weight rows against 8 sequences with packed fp32 FMAs (FFMA2: weights are up-cast once per
//     256-element chunk and reused for 8 sequences, every activation LDS.128 is reused for 4 rows); the 32 sums of
//     a unit are reduced across the warp with one transposing butterfly (31 shuffles).
//   * o_proj / down_proj (7 rows per CTA, K = 2048 / 3072) keep their rows resident in the ring and walk K in
//     chunks of H (the capacity of xs); their units split K across warps and combine through shared memory.
//   * attention work items are (sequence, kv head, KVK-key split of the CACHED keys), dealt round-robin to the
//     CTAs; K and V tiles travel through a two-slot shared-memory stage that the producer refills while the
//     consumers compute; one CTA per (sequence, kv head) merges the partials, folds in the current token's own
//     key/value and appends it to the cache (replaces Tensor::cat, layers.rs:311-317).
//   * greedy bookkeeping (argmax, EOS, append, next embedding; inference.rs:161-170) for every sequence by the
//     last CTA to finish the lm_head.
// Arithmetic per (row, sequence) is the fp32 FMA chain of decode_mega.cu (same lane -> element mapping, same
// reduction tree for the K = H phases), so a batch reproduces the single-sequence results.
#include <algorithm>
#include "mega_common.cuh"

namespace asrb {

namespace megab {
using namespace mega;

static constexpr int MAXSPLIT = 32;       // partial records per (sequence, kv head): lanes of the merging warps
static constexpr int MAXROWS = 8;         // residual rows owned by one CTA (H / gridDim.x rounded up)

struct Params {
    const DecLayerW* layers;     // device array [L]; ln_in / ln_post in the xs_swz layout
    const bf16* lm_head;
    const bf16* embed;
    const float* final_norm;     // xs_swz layout
    const float* rope_cos; const float* rope_sin;
    float eps;
    int L, H, QD, KVD, I, V, nkv;
    int nb;                      // active sequences of this launch (<= NB)
    float* x;                    // [nb][H] embedding of the pending tokens (in) / of the next tokens (out)
    float* kcache; float* vcache; size_t cache_layer_stride, cache_seq_stride; int max_ctx;
    float* part_val; int* part_idx; int n_part;     // [nb][n_part] argmax partials (first gridDim.x used)
    int* pos; int* done; int* next_id; int* ids_out; int* n_out; int max_new;
    unsigned* bar;               // [0] finish ticket, [1] launch epoch
    uint2* qkv_ll; uint2* part_ll; uint2* attn_ll; uint2* x_ll; uint2* act_ll;
};

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// packed fp32 FMA (FFMA2): d = a * b + c on both halves
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

// 32 per-lane partial sums -> lane L ends with the warp total of value L (tree: xor 16, 8, 4, 2, 1 -- the tree of
// warp_sum / row_dot2 of the single-sequence kernel)
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16, n = 32; n > 1; o >>= 1, n >>= 1) {
        const bool up = lane & o;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}

// One unit: 4 weight rows x 8 sequences over NSUB 256-element sub-chunks.
//   wrow[r]  : row r of the unit in shared memory (uint4 index 0 = first element of the range contracted)
//   xs8      : xs row of the unit's first sequence (float index 0 = same element), XSTR floats between sequences
// acc[r][s] = (a0, a1): even / odd element chains exactly as row_dot() of the single-sequence kernel.
template <int XSTR>
__device__ __forceinline__ void unit_fma(const uint4* const (&wrow)[4], const float* xs8, int nsub, int lane,
                                         float2 (&acc)[4][8]) {
    const int sw = ((lane >> 2) & 1) * 4;
#pragma unroll 1
    for (int c = 0; c < nsub; ++c) {
        float2 wp[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint4 w = wrow[r][c * 32 + lane];
            wp[r][0] = make_float2(bf16_lo(w.x), bf16_hi(w.x)); wp[r][1] = make_float2(bf16_lo(w.y), bf16_hi(w.y));
            wp[r][2] = make_float2(bf16_lo(w.z), bf16_hi(w.z)); wp[r][3] = make_float2(bf16_lo(w.w), bf16_hi(w.w));
        }
        const float* xc = xs8 + (c * 32 + lane) * 8;
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            const float4 xa = *reinterpret_cast<const float4*>(xc + s * XSTR + sw);
            const float4 xb = *reinterpret_cast<const float4*>(xc + s * XSTR + 4 - sw);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r][s] = ffma2(wp[r][0], make_float2(xa.x, xa.y), acc[r][s]);
                acc[r][s] = ffma2(wp[r][1], make_float2(xa.z, xa.w), acc[r][s]);
                acc[r][s] = ffma2(wp[r][2], make_float2(xb.x, xb.y), acc[r][s]);
                acc[r][s] = ffma2(wp[r][3], make_float2(xb.z, xb.w), acc[r][s]);
            }
        }
    }
}
// lane L <- total of (row L >> 3, sequence L & 7)
__device__ __forceinline__ float unit_reduce(float2 (&acc)[4][8], int lane) {
    float v[32];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 8; ++s) v[r * 8 + s] = acc[r][s].x + acc[r][s].y;
    return transpose_reduce32(v, lane);
}

// per-head RMSNorm + RoPE of one 128-vector by one warp (lane holds d = lane, +32, +64, +96); input = tagged words
__device__ __forceinline__ void head_norm_rope_b(const uint2* __restrict__ src, uint32_t tag, const float* __restrict__ nw,
                                                 float eps, const float* __restrict__ cs, const float* __restrict__ sn,
                                                 float* dst, int lane) {
    float v[4];
    ll_poll4(src + lane, 32, tag, v);
    const float ss = warp_sum(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    const float r = 1.0f / sqrtf(ss / 128.f + eps);
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (v[i] * r) * __ldg(nw + lane + 32 * i);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int d = lane + 32 * i;
        const float c = cs[d], s = sn[d];
        const float a = v[i], b = v[i + 2];
        dst[d] = a * c - b * s;
        dst[d + 64] = b * c + a * s;
    }
}

enum { BE_STORE = 0, BE_SWIGLU = 1, BE_ARGMAX = 2 };

// weight stream of one CTA in phase order: (layer, phase) slices, SLOT_BYTES chunks
template <int H, int QD, int I>
struct WCursor {
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
    __device__ void init(const Params& p, const DecLayerW* table) { ltab = table; l = 0; ph = 0; done = false; load(p); skip(p); }
    __device__ void skip(const Params& p) {
        while (!done && r >= s.r1) {
            if (l >= p.L) { done = true; break; }
            if (++ph == 4) { ph = 0; ++l; }
            load(p);
        }
    }
    // current chunk (valid while !done); advance() moves on
    __device__ void cur(const bf16*& src, uint32_t& bytes) const {
        const int rows = min(s.rpc, s.r1 - r);
        src = s.W + (size_t)r * s.K; bytes = (uint32_t)rows * s.K * 2;
    }
    __device__ void advance(const Params& p) { r += s.rpc; skip(p); }
};

template <int H, int QD, int I, int NB, int NS, int KVK>
__global__ void __launch_bounds__(NTHREADS, 1) decode_batch_kernel(const Params p) {
    static_assert(NB % 8 == 0 && NB <= 16, "NB must be 8 or 16");
    static_assert(H % 256 == 0 && QD % H == 0 && I % H == 0, "chunking needs QD, I multiples of H, H multiple of 256");
    constexpr int NSG = NB / 8;                 // sequence groups of 8
    constexpr int XSTR = H;                     // floats between the xs rows of consecutive sequences
    constexpr int KPW = KVK / NCONS_WARPS;      // keys per warp in an attention tile
    constexpr int NV = 2 * KPW;                 // scores per lane before the butterfly (keys x 2 heads)
    constexpr int KV_TILE = KVK * HD * 4;
    constexpr int GROUP = 2;                    // q heads per kv head (checked on the host)
    extern __shared__ __align__(128) uint8_t smem[];
    Ring ring;
    ring.slots = smem; ring.nslot = NS;
    uint8_t* kv_smem = smem + (size_t)NS * SLOT_BYTES;                 // [K tile | V tile]
    float* xs = reinterpret_cast<float*>(kv_smem + 2 * KV_TILE);       // [NB][XSTR]
    float* xres = xs + NB * XSTR;                                       // [NB][MAXROWS]
    float* ropes = xres + NB * MAXROWS;                                 // [NB][128]  cos | sin of each sequence's position
    float* redk = ropes + NB * 128;                                     // [8 warps][32] K-split partials / scratch
    float* ssred = redk + NCONS_WARPS * 32;                             // [8 warps][NB] sums of squares
    float* rs = ssred + NCONS_WARPS * NB;                               // [NB] RMSNorm scales
    float* bestv = rs + NB;                                             // [8 warps][NB]
    int* besti = reinterpret_cast<int*>(bestv + NCONS_WARPS * NB);      // [8 warps][NB]
    int* seqi = besti + NCONS_WARPS * NB;                               // pos[NB] | nact[NB] | off[NB + 1] | misc[4]
    DecLayerW* ltab = reinterpret_cast<DecLayerW*>(seqi + 3 * NB + 8);  // [MAX_LAYERS]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ltab + MAX_LAYERS);
    ring.full = bars; ring.empty = bars + NSLOT_MAX;
    uint64_t* kv_full = bars + 2 * NSLOT_MAX; uint64_t* kv_empty = kv_full + 2;   // [2] each: K stage, V stage
    int* pos_s = seqi; int* nact_s = seqi + NB; int* off_s = seqi + 2 * NB; int* misc = seqi + 3 * NB + 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool is_producer = warp == NCONS_WARPS;
    const int nb = p.nb;
    const unsigned G = gridDim.x;

    {   // every sequence finished: nothing to do this step
        bool all = true;
        for (int b = 0; b < nb; ++b) all = all && (__ldcg(p.done + b) != 0);
        if (all) return;
    }
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&ring.full[i], 1); mbar_init(&ring.empty[i], NCONS_WARPS); }
        for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], NCONS_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        int o = 0;
        for (int b = 0; b < NB; ++b) {
            const int ps = b < nb ? __ldcg(p.pos + b) : 0;
            pos_s[b] = ps; nact_s[b] = (ps + KVK - 1) / KVK; off_s[b] = o; o += nact_s[b];
        }
        off_s[NB] = o;
    }
    {
        const uint2* src = reinterpret_cast<const uint2*>(p.layers);
        uint2* dst = reinterpret_cast<uint2*>(ltab);
        for (int i = tid; i < p.L * (int)(sizeof(DecLayerW) / 8); i += NTHREADS) dst[i] = src[i];
        for (int i = tid; i < nb * 128; i += NTHREADS) {
            const int b = i >> 7, d = i & 127;
            const int ps = __ldcg(p.pos + b);
            ropes[i] = d < 64 ? p.rope_cos[(size_t)ps * 64 + d] : p.rope_sin[(size_t)ps * 64 + d - 64];
        }
    }
    __syncthreads();
    const int T = off_s[NB] * p.nkv;            // attention work items of this step: (sequence, kv head, split)
    // item t -> (b, g, sp): sequence-major, then kv head, then split
    auto item_decode = [&](int t, int& b, int& g, int& sp) {
        b = 0;
#pragma unroll
        for (int i = 1; i < NB; ++i) if (t >= off_s[i] * p.nkv) b = i;
        const int r = t - off_s[b] * p.nkv, na = nact_s[b];
        g = r / na; sp = r - g * na;
    };

    if (is_producer) {
        if (lane == 0) {
            WCursor<H, QD, I> wc;
            wc.init(p, ltab);
            uint32_t q = 0, kq = 0;
            int kl = 0, kt = (int)blockIdx.x;     // K/V stream: (layer, item); item uses: 2 * i (K), 2 * i + 1 (V)
            bool kv_done = (kt >= T);
            int kb = 0, kg = 0, ksp = 0;
            if (!kv_done) item_decode(kt, kb, kg, ksp);
            while (!wc.done || !kv_done) {
                bool prog = false;
                if (!wc.done) {
                    const uint32_t slot = q % NS, par = (q / NS) & 1;
                    if (mbar_test(&ring.empty[slot], par ^ 1)) {
                        const bf16* src; uint32_t bytes;
                        wc.cur(src, bytes);
                        mbar_expect_tx(&ring.full[slot], bytes);
                        bulk_g2s(ring.slots + (size_t)slot * SLOT_BYTES, src, bytes, &ring.full[slot]);
                        wc.advance(p); ++q; prog = true;
                    }
                }
                if (!kv_done) {
                    const uint32_t st = kq & 1, par = (kq >> 1) & 1;
                    if (mbar_test(&kv_empty[st], par ^ 1)) {
                        const int nloc = min(KVK, pos_s[kb] - ksp * KVK);
                        const size_t off = (size_t)kl * p.cache_layer_stride + (size_t)kb * p.cache_seq_stride +
                                           ((size_t)kg * p.max_ctx + (size_t)ksp * KVK) * HD;
                        const uint32_t bytes = (uint32_t)nloc * HD * 4;
                        mbar_expect_tx(&kv_full[st], bytes);
                        bulk_g2s(kv_smem + (size_t)st * KV_TILE, (st == 0 ? p.kcache : p.vcache) + off, bytes, &kv_full[st]);
                        ++kq; prog = true;
                        if (st == 1) {          // V issued: next item
                            kt += (int)G;
                            if (kt >= T) { kt = (int)blockIdx.x; if (++kl >= p.L) kv_done = true; }
                            if (!kv_done) item_decode(kt, kb, kg, ksp);
                        }
                    }
                }
                if (!prog) __nanosleep(20);
            }
        }
        return;
    }

    // ------------------------------ consumers ------------------------------
    uint32_t q = 0, kq = 0;
    const unsigned epoch = __ldcg(p.bar + 1);
    const uint32_t tag_base = (epoch & 0xffffffu) << 8;
    const Slice xsl = make_slice(nullptr, H, QD, 1);                  // residual rows owned by this CTA
    const int xrows = xsl.r1 - xsl.r0;
    Slice sl_qkv = make_slice(nullptr, QD + 2 * p.KVD, H, 1), sl_o = make_slice(nullptr, H, QD, 1),
          sl_gu = make_slice(nullptr, 2 * I, H, 2), sl_dn = make_slice(nullptr, H, I, 1);
    for (int i = tid; i < nb * xrows; i += NCONS) {
        const int b = i / xrows, r = i - b * xrows;
        xres[b * MAXROWS + r] = __ldcg(p.x + (size_t)b * H + xsl.r0 + r);
    }
    float best_v[NSG]; int best_i[NSG];
#pragma unroll
    for (int g = 0; g < NSG; ++g) { best_v[g] = -INFINITY; best_i[g] = 0x7fffffff; }

    // ---- gather of one H-long chunk of every active sequence into xs (raw values), optional RMSNorm ----
    // src: tagged words, sequence b at src + b * src_stride; NORM: xs <- (x * r_b) * w  (rounding order of layers.rs:48-54)
    constexpr int PP = (H / 2 + NCONS - 1) / NCONS;     // word pairs per thread and sequence
    constexpr int SB = 8 / PP > 0 ? 8 / PP : 1;         // sequences per polling round (8 loads in flight per thread)
    auto gather = [&](const uint2* src, size_t src_stride, uint32_t tag, const float* normw) {
        float2 wv[PP];
        if (normw) {
#pragma unroll
            for (int i = 0; i < PP; ++i) {
                const int j = tid + i * NCONS;
                wv[i] = (j < H / 2) ? __ldg(reinterpret_cast<const float2*>(normw + xs_swz(2 * j))) : make_float2(0.f, 0.f);
            }
        }
        float ss[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) ss[b] = 0.f;
#pragma unroll
        for (int b0 = 0; b0 < NB; b0 += SB) {
            if (b0 < nb) {
                uint4 v[SB][PP];
                bool ok;
                do {      // loads are unconditional (idle slots re-poll the last active sequence / word pair 0): registers only
                    ok = true;
#pragma unroll
                    for (int sb = 0; sb < SB; ++sb)
#pragma unroll
                        for (int i = 0; i < PP; ++i) {
                            const int j = tid + i * NCONS;
                            const int bb = min(b0 + sb, nb - 1), jj = (PP * NCONS > H / 2 && j >= H / 2) ? 0 : j;
                            v[sb][i] = ll_load2(src + (size_t)bb * src_stride + 2 * jj);
                        }
#pragma unroll
                    for (int sb = 0; sb < SB; ++sb)
#pragma unroll
                        for (int i = 0; i < PP; ++i) ok = ok && (v[sb][i].y == tag) && (v[sb][i].w == tag);
                } while (!ok);
#pragma unroll
                for (int sb = 0; sb < SB; ++sb)
#pragma unroll
                    for (int i = 0; i < PP; ++i) {
                        const int j = tid + i * NCONS;
                        if (b0 + sb < nb && j < H / 2) {
                            const float a = __uint_as_float(v[sb][i].x), c = __uint_as_float(v[sb][i].z);
                            *reinterpret_cast<float2*>(xs + (b0 + sb) * XSTR + xs_swz(2 * j)) = make_float2(a, c);
                            ss[b0 + sb] = fmaf(a, a, ss[b0 + sb]); ss[b0 + sb] = fmaf(c, c, ss[b0 + sb]);
                        }
                    }
            }
        }
        if (normw) {
#pragma unroll
            for (int b = 0; b < NB; ++b) { const float t = warp_sum(ss[b]); if (lane == 0) ssred[warp * NB + b] = t; }
            cons_sync();
            if (tid < nb) {
                float tot = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < NCONS_WARPS; ++w8) tot += ssred[w8 * NB + tid];
                rs[tid] = 1.0f / sqrtf(tot / H + p.eps);
            }
            cons_sync();
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                if (b < nb) {
                    const float r = rs[b];
#pragma unroll
                    for (int i = 0; i < PP; ++i) {
                        const int j = tid + i * NCONS;
                        if (j < H / 2) {
                            float2* px = reinterpret_cast<float2*>(xs + b * XSTR + xs_swz(2 * j));
                            const float2 xv = *px;
                            *px = make_float2((xv.x * r) * wv[i].x, (xv.y * r) * wv[i].y);
                        }
                    }
                }
            }
        }
        cons_sync();
    };
    // layer 0: the pending tokens' embeddings come from plain memory (written by the previous step / prefill)
    auto load_x0 = [&](const float* normw) {
        float ss[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            ss[b] = 0.f;
            if (b < nb)
                for (int i = tid; i < H; i += NCONS) { const float v = __ldcg(p.x + (size_t)b * H + i); xs[b * XSTR + xs_swz(i)] = v; ss[b] = fmaf(v, v, ss[b]); }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) { const float t = warp_sum(ss[b]); if (lane == 0) ssred[warp * NB + b] = t; }
        cons_sync();
        if (tid < nb) {
            float tot = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < NCONS_WARPS; ++w8) tot += ssred[w8 * NB + tid];
            rs[tid] = 1.0f / sqrtf(tot / H + p.eps);
        }
        cons_sync();
        for (int b = 0; b < nb; ++b) {
            const float r = rs[b];
            for (int i = tid; i < H; i += NCONS) { const int e = xs_swz(i); xs[b * XSTR + e] = (xs[b * XSTR + e] * r) * __ldg(normw + e); }
        }
        cons_sync();
    };

    // ---- K = H phases: rows stream through the ring, unit = (4 rows of a slot, 8 sequences) ----
    auto rows_phase = [&](const Slice& s, int epi, uint2* out, size_t out_stride, uint32_t tag) {
        int ubase = 0;
        for (int r = s.r0; r < s.r1; r += s.rpc, ++q) {
            const int rows = min(s.rpc, s.r1 - r);
            const uint32_t slot = q % NS, par = (q / NS) & 1;
            mbar_wait(&ring.full[slot], par);
            const uint4* base = reinterpret_cast<const uint4*>(ring.slots + (size_t)slot * SLOT_BYTES);
            const int units = ((rows + 3) >> 2) * NSG;
            for (int j = (warp - (ubase & 7) + 8) & 7; j < units; j += NCONS_WARPS) {
                const int rg = j / NSG, sg = j - rg * NSG;
                const uint4* wrow[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) wrow[i] = base + (size_t)(rg * 4 + i) * (H / 8);
                float2 acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc[i][k] = make_float2(0.f, 0.f);
                unit_fma<XSTR>(wrow, xs + sg * 8 * XSTR, H / 256, lane, acc);
                const float v = unit_reduce(acc, lane);
                const int rr = rg * 4 + (lane >> 3), sq = sg * 8 + (lane & 7), row = r + rr;
                const bool valid = rr < rows && sq < nb;
                if (epi == BE_STORE) {
                    if (valid) ll_store(out + (size_t)sq * out_stride + row, v, tag);
                } else if (epi == BE_SWIGLU) {
                    const float up = __shfl_xor_sync(0xffffffffu, v, 8);      // rows 2j (gate) and 2j + 1 (up) sit 8 lanes apart
                    if (valid && !(rr & 1)) ll_store(out + (size_t)sq * out_stride + (row >> 1), silu(v) * up, tag);
                } else {
#pragma unroll
                    for (int g = 0; g < NSG; ++g)
                        if (valid && sg == g && (v > best_v[g] || (v == best_v[g] && row < best_i[g]))) { best_v[g] = v; best_i[g] = row; }
                }
            }
            ubase += units;
            __syncwarp();
            if (lane == 0) mbar_arrive(&ring.empty[slot]);
        }
    };

    // ---- K = NCH * H phases with <= 8 resident rows (o_proj, down_proj): x walks through xs in H-long chunks,
    //      units = (4 rows, 8 sequences, 1 / KP of every chunk); partials combine through shared memory;
    //      result: residual add + publication (layers.rs:454,460) ----
    auto resident_phase = [&](const Slice& s, int NCH, const uint2* src, size_t src_stride, uint32_t src_tag, uint32_t tag) {
        const int rows = s.r1 - s.r0;
        const int nslots = (rows + s.rpc - 1) / s.rpc;
        const int nrg = (rows + 3) >> 2;
        int KP = 8 / (nrg * NSG); if (KP > H / 256) KP = H / 256; if (KP > 4) KP = 4;
        const int u = warp / KP, kp = warp - u * KP;
        const bool active = u < nrg * NSG;
        const int rg = u / NSG, sg = u - rg * NSG;
        const uint8_t* rowp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int ri = min(rg * 4 + i, rows - 1);                     // rows past the slice alias the last row (never published)
            const uint32_t slot = (q + ri / s.rpc) % NS;
            rowp[i] = ring.slots + (size_t)slot * SLOT_BYTES + (size_t)(ri % s.rpc) * s.K * 2;
        }
        float2 acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = make_float2(0.f, 0.f);
        const int nsub = (H / 256) / KP;
        for (int ch = 0; ch < NCH; ++ch) {
            gather(src + (size_t)ch * H, src_stride, src_tag, nullptr);
            if (ch == 0)
                for (int i = 0; i < nslots; ++i) mbar_wait(&ring.full[(q + i) % NS], ((q + i) / NS) & 1);
            if (active) {
                const uint4* wrow[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) wrow[i] = reinterpret_cast<const uint4*>(rowp[i] + ((size_t)ch * H + (size_t)kp * nsub * 256) * 2);
                unit_fma<XSTR>(wrow, xs + sg * 8 * XSTR + kp * nsub * 256, nsub, lane, acc);
            }
            cons_sync();                                                  // xs may be overwritten by the next chunk
        }
        __syncwarp();
        if (lane == 0) for (int i = 0; i < nslots; ++i) mbar_arrive(&ring.empty[(q + i) % NS]);
        q += nslots;
        const float v = unit_reduce(acc, lane);
        redk[warp * 32 + lane] = v;
        cons_sync();
        if (active && kp == 0) {
            float tot = v;
            for (int k = 1; k < KP; ++k) tot += redk[(warp + k) * 32 + lane];
            const int rr = rg * 4 + (lane >> 3), sq = sg * 8 + (lane & 7);
            if (rr < rows && sq < nb) {
                const float nv = xres[sq * MAXROWS + rr] + tot;
                xres[sq * MAXROWS + rr] = nv;
                ll_store(p.x_ll + (size_t)sq * H + s.r0 + rr, nv, tag);
            }
        }
    };

    for (int l = 0; l < p.L; ++l) {
        const DecLayerW w = ltab[l];
        const uint32_t tl = tag_base | ((uint32_t)l << 3);
        // ---- phase 1: RMSNorm + [q|k|v] GEMV ----
        if (l == 0) load_x0(w.ln_in);
        else gather(p.x_ll, H, (tag_base | ((uint32_t)(l - 1) << 3)) | PH_XD, w.ln_in);
        sl_qkv.W = w.wqkv;
        rows_phase(sl_qkv, BE_STORE, p.qkv_ll, (size_t)(QD + 2 * p.KVD), tl | PH_QKV);
        cons_sync();                                  // xs is free: attention scratch aliases it
        // ---- phase 2: attention partials of this CTA's work items ----
        {
            float* qs = xs;                           // [2][128]
            float* kn = qs + GROUP * HD;              // [128]
            float* vn = kn + HD;                      // [128]
            float* osum = vn + HD;                    // [warps][2][128]
            float* wml = osum + NCONS_WARPS * GROUP * HD;   // [warps][2][2]
            float* snew = wml + NCONS_WARPS * 4;      // [2]
            float* Ks = reinterpret_cast<float*>(kv_smem);
            float* Vs = reinterpret_cast<float*>(kv_smem + KV_TILE);
            for (int t = (int)blockIdx.x; t < T; t += (int)G) {
                int b, g, sp;
                item_decode(t, b, g, sp);
                const int nloc = min(KVK, pos_s[b] - sp * KVK);
                const float* cs = ropes + b * 128; const float* sn = cs + 64;
                if (warp < GROUP)
                    head_norm_rope_b(p.qkv_ll + (size_t)b * (QD + 2 * p.KVD) + (size_t)(g * GROUP + warp) * HD, tl | PH_QKV,
                                     w.qnorm, p.eps, cs, sn, qs + warp * HD, lane);
                cons_sync();
                const float4 q0 = *reinterpret_cast<const float4*>(qs + lane * 4);
                const float4 q1 = *reinterpret_cast<const float4*>(qs + HD + lane * 4);
                mbar_wait(&kv_full[0], (kq >> 1) & 1);
                float pv[NV];
#pragma unroll
                for (int kk = 0; kk < KPW; ++kk) {
                    const int j = warp + 8 * kk;      // rows past the split's last key hold stale data and are masked below
                    const float4 kv = *reinterpret_cast<const float4*>(Ks + j * HD + lane * 4);
                    pv[2 * kk] = fmaf(kv.x, q0.x, fmaf(kv.y, q0.y, fmaf(kv.z, q0.z, kv.w * q0.w)));
                    pv[2 * kk + 1] = fmaf(kv.x, q1.x, fmaf(kv.y, q1.y, fmaf(kv.z, q1.z, kv.w * q1.w)));
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&kv_empty[0]);     // K stage may be refilled
                // transposing butterfly: lane L ends with score index L >> SH (index = 2 * key + head)
                constexpr int SH = (NV == 16) ? 1 : (NV == 8 ? 2 : 3);
#pragma unroll
                for (int o = 16, n = NV; n > 1; o >>= 1, n >>= 1) {
                    const bool up = lane & o;
#pragma unroll
                    for (int i = 0; i < n / 2; ++i) {
                        const float send = up ? pv[i] : pv[i + n / 2];
                        const float keep = up ? pv[i + n / 2] : pv[i];
                        pv[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                    }
                }
#pragma unroll
                for (int o = (1 << SH) >> 1; o > 0; o >>= 1) pv[0] += __shfl_xor_sync(0xffffffffu, pv[0], o);
                const int sidx = lane >> SH;                                   // 2 * kk + head
                const bool mine = warp + 8 * (sidx >> 1) < nloc;
                const float sv = mine ? pv[0] / sqrtf((float)HD) : -INFINITY;
                float mw = sv;                                                 // max over this warp's keys, per head
#pragma unroll
                for (int o = 2 << SH; o < 32; o <<= 1) mw = fmaxf(mw, __shfl_xor_sync(0xffffffffu, mw, o));
                const float ev = mine ? expf(sv - mw) : 0.f;
                float lw = ev;
#pragma unroll
                for (int o = 2 << SH; o < 32; o <<= 1) lw += __shfl_xor_sync(0xffffffffu, lw, o);
                mbar_wait(&kv_full[1], (kq >> 1) & 1);
                float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
#pragma unroll
                for (int kk = 0; kk < KPW; ++kk) {
                    const int j = warp + 8 * kk;
                    const float4 vv = *reinterpret_cast<const float4*>(Vs + j * HD + lane * 4);
                    const float e0 = __shfl_sync(0xffffffffu, ev, (2 * kk) << SH), e1 = __shfl_sync(0xffffffffu, ev, (2 * kk + 1) << SH);
                    if (j < nloc) {       // (a stale V row may hold non-finite garbage: 0 * inf must not reach the sum)
                        o0.x = fmaf(e0, vv.x, o0.x); o0.y = fmaf(e0, vv.y, o0.y); o0.z = fmaf(e0, vv.z, o0.z); o0.w = fmaf(e0, vv.w, o0.w);
                        o1.x = fmaf(e1, vv.x, o1.x); o1.y = fmaf(e1, vv.y, o1.y); o1.z = fmaf(e1, vv.z, o1.z); o1.w = fmaf(e1, vv.w, o1.w);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&kv_empty[1]);     // V stage may be refilled
                kq += 2;
                *reinterpret_cast<float4*>(osum + (warp * 2 + 0) * HD + lane * 4) = o0;
                *reinterpret_cast<float4*>(osum + (warp * 2 + 1) * HD + lane * 4) = o1;
                if (lane == 0 || lane == (1 << SH)) { const int h = lane >> SH; wml[(warp * 2 + h) * 2] = mw; wml[(warp * 2 + h) * 2 + 1] = lw; }
                cons_sync();
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
                    uint2* rec = p.part_ll + ((((size_t)b * p.nkv + g) * MAXSPLIT + sp) * GROUP + hq) * PSTRIDE;
                    ll_store(rec + d, acc, tl | PH_PART);
                    if (d < 2) ll_store(rec + HD + d, d == 0 ? M : Ls, tl | PH_PART);
                }
                cons_sync();                          // scratch may be overwritten by the next item
            }
            // ---- merge of one (sequence, kv head): all its splits + the current token's key / value ----
            const int mid = (int)G - 1 - (int)blockIdx.x;
            if (mid < nb * p.nkv) {
                const int b = mid / p.nkv, g = mid - b * p.nkv;
                const int pos = pos_s[b], nact = nact_s[b];
                const float* cs = ropes + b * 128; const float* sn = cs + 64;
                const uint2* qkvb = p.qkv_ll + (size_t)b * (QD + 2 * p.KVD);
                if (warp < GROUP) head_norm_rope_b(qkvb + (size_t)(g * GROUP + warp) * HD, tl | PH_QKV, w.qnorm, p.eps, cs, sn, qs + warp * HD, lane);
                else if (warp == GROUP) head_norm_rope_b(qkvb + QD + (size_t)g * HD, tl | PH_QKV, w.knorm, p.eps, cs, sn, kn, lane);
                else if (warp == GROUP + 1) {
                    float vv[4];
                    ll_poll4(qkvb + QD + p.KVD + (size_t)g * HD + lane, 32, tl | PH_QKV, vv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) vn[lane + 32 * i] = vv[i];
                }
                cons_sync();
                if (tid < HD) {                       // KV append (replaces Tensor::cat, layers.rs:311-317)
                    const size_t off = (size_t)l * p.cache_layer_stride + (size_t)b * p.cache_seq_stride + ((size_t)g * p.max_ctx + pos) * HD;
                    p.kcache[off + tid] = kn[tid]; p.vcache[off + tid] = vn[tid];
                }
                if (warp >= NCONS_WARPS - GROUP) {    // score of the new key for head hq
                    const int hq = warp - (NCONS_WARPS - GROUP);
                    const float4 a = *reinterpret_cast<const float4*>(qs + hq * HD + lane * 4);
                    const float4 c4 = *reinterpret_cast<const float4*>(kn + lane * 4);
                    const float s_ = warp_sum(fmaf(a.x, c4.x, fmaf(a.y, c4.y, fmaf(a.z, c4.z, a.w * c4.w))));
                    if (lane == 0) snew[hq] = s_ / sqrtf((float)HD);
                }
                cons_sync();
                {
                    const int hq = tid / HD, d = tid - hq * HD;          // hq is uniform per warp (HD = 4 warps)
                    const uint32_t tg = tl | PH_PART;
                    constexpr int RB = 8;                                // partial outputs fetched per round (registers)
                    uint2 mv, lv, ov[RB];
                    bool ok;
                    const uint2* recb = p.part_ll + (((size_t)b * p.nkv + g) * MAXSPLIT * GROUP + hq) * PSTRIDE;
                    auto load_round = [&](int u0) {
#pragma unroll
                        for (int u = 0; u < RB; ++u)
                            if (u0 + u < nact) {
                                const uint2* rec = recb + (size_t)(u0 + u) * GROUP * PSTRIDE;
                                asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(ov[u].x), "=r"(ov[u].y) : "l"(rec + d) : "memory");
                            }
                    };
                    auto round_ok = [&](int u0) {
                        bool k = true;
#pragma unroll
                        for (int u = 0; u < RB; ++u) if (u0 + u < nact) k = k && (ov[u].y == tg);
                        return k;
                    };
                    do {        // first round: (max, sum) of every split (one per lane) + the first RB partial outputs
                        mv.y = tg; lv.y = tg; mv.x = 0u; lv.x = 0u;
                        if (lane < nact) {
                            const uint2* rec = recb + (size_t)lane * GROUP * PSTRIDE;
                            asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(mv.x), "=r"(mv.y) : "l"(rec + HD) : "memory");
                            asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(lv.x), "=r"(lv.y) : "l"(rec + HD + 1) : "memory");
                        }
                        load_round(0);
                        ok = __all_sync(0xffffffffu, (mv.y == tg) && (lv.y == tg) && round_ok(0));
                    } while (!ok);
                    // softmax merge, one partial per lane: lanes < nact hold a split; the current token's key is one more
                    // partial (max = its score, sum = 1, output = its value row) handled outside the lane array
                    const float m_l = lane < nact ? __uint_as_float(mv.x) : -INFINITY;
                    const float l_l = lane < nact ? __uint_as_float(lv.x) : 0.f;
                    const float sn_ = snew[hq];
                    const float M = fmaxf(warp_max(m_l), sn_);
                    const float f = expf(m_l - M);                       // exp(-inf) = 0 on idle lanes
                    const float fn = expf(sn_ - M);
                    const float Lsum = warp_sum(f * l_l) + fn;
                    float O = fn * vn[d];
                    for (int u0 = 0; u0 < nact; u0 += RB) {
                        if (u0 > 0) { do { load_round(u0); ok = __all_sync(0xffffffffu, round_ok(u0)); } while (!ok); }
#pragma unroll
                        for (int u = 0; u < RB; ++u)
                            if (u0 + u < nact) O = fmaf(__shfl_sync(0xffffffffu, f, u0 + u), __uint_as_float(ov[u].x), O);
                    }
                    ll_store(p.attn_ll + (size_t)b * QD + (size_t)(g * GROUP + hq) * HD + d, O / Lsum, tl | PH_ATTN);
                }
                cons_sync();                          // attention scratch (aliases xs) is free again
            }
        }
        // ---- phase 3: o_proj GEMV + residual ----
        sl_o.W = w.wo;
        resident_phase(sl_o, QD / H, p.attn_ll, QD, tl | PH_ATTN, tl | PH_XO);
        // ---- phase 4: RMSNorm + gate/up GEMV + SiLU*mul ----
        gather(p.x_ll, H, tl | PH_XO, w.ln_post);
        sl_gu.W = w.wgu;
        rows_phase(sl_gu, BE_SWIGLU, p.act_ll, (size_t)I, tl | PH_ACT);
        cons_sync();
        // ---- phase 5: down GEMV + residual ----
        sl_dn.W = w.wdown;
        resident_phase(sl_dn, I / H, p.act_ll, I, tl | PH_ACT, tl | PH_XD);
    }
    // ---- final RMSNorm + tied lm_head GEMV + argmax ----
    gather(p.x_ll, H, (tag_base | ((uint32_t)(p.L - 1) << 3)) | PH_XD, p.final_norm);
    rows_phase(make_slice(p.lm_head, p.V, H, 1), BE_ARGMAX, nullptr, 0, 0u);
    // candidates of sequence sq live in the 4 lanes sharing (lane & 7): merge over the row bits (lane bits 3, 4)
#pragma unroll
    for (int g = 0; g < NSG; ++g) {
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best_v[g], o); const int oi = __shfl_xor_sync(0xffffffffu, best_i[g], o);
            if (ov > best_v[g] || (ov == best_v[g] && oi < best_i[g])) { best_v[g] = ov; best_i[g] = oi; }
        }
        if (lane < 8) { bestv[warp * NB + g * 8 + lane] = best_v[g]; besti[warp * NB + g * 8 + lane] = best_i[g]; }
    }
    cons_sync();
    int& is_last = misc[0];
    if (tid < nb) {
        float v = -INFINITY; int idx = 0x7fffffff;
        for (int wq = 0; wq < NCONS_WARPS; ++wq) {
            const float cv = bestv[wq * NB + tid]; const int ci = besti[wq * NB + tid];
            if (cv > v || (cv == v && ci < idx)) { v = cv; idx = ci; }
        }
        p.part_val[(size_t)tid * p.n_part + blockIdx.x] = v; p.part_idx[(size_t)tid * p.n_part + blockIdx.x] = idx;
        __threadfence();
    }
    cons_sync();
    if (tid == 0) {
        __threadfence();
        const unsigned t = atomicAdd(p.bar, 1u);
        is_last = (t == G - 1);
    }
    cons_sync();
    if (!is_last) return;
    // ---- greedy bookkeeping by the last CTA (inference.rs:161-170), one warp per sequence ----
    __threadfence();
    for (int b = warp; b < nb; b += NCONS_WARPS) {
        if (__ldcg(p.done + b) != 0) { if (lane == 0) p.next_id[b] = -1; continue; }
        float v = -INFINITY; int idx = 0x7fffffff;
        for (int i = lane; i < (int)G; i += 32) {
            const float pv = __ldcg(p.part_val + (size_t)b * p.n_part + i); const int pi = __ldcg(p.part_idx + (size_t)b * p.n_part + i);
            if (pv > v || (pv == v && pi < idx)) { v = pv; idx = pi; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o); const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        int tok = idx;
        const int n = p.n_out[b];
        if (tok == 151643 || tok == 151645 || n >= p.max_new) {
            if (lane == 0) { p.done[b] = 1; p.next_id[b] = -1; }
            tok = -1;
        } else if (lane == 0) {
            p.ids_out[(size_t)b * p.max_new + n] = tok; p.n_out[b] = n + 1; p.pos[b] = pos_s[b] + 1; p.next_id[b] = tok;
        }
        if (tok >= 0) {
            const bf16* e = p.embed + (size_t)tok * H;
            for (int i = lane; i < H; i += 32) p.x[(size_t)b * H + i] = __bfloat162float(e[i]);
        }
    }
    cons_sync();
    if (tid == 0) {
        p.bar[0] = 0;                        // every CTA has taken its ticket: reset for the next launch
        p.bar[1] = p.bar[1] + 1;             // new epoch: words published by this step can never match again
    }
}

}  // namespace megab

// host side ---------------------------------------------------------------------------------------
template <int H, int QD, int I> static bool bdims_match(const asrb_dims& c) {
    return c.hidden_size == H && c.num_attention_heads * c.head_dim == QD && c.intermediate_size == I;
}
struct BatchCfg { int NB, NS, KVK; };
static BatchCfg batch_cfg(int B) { return B <= 8 ? BatchCfg{8, 3, 64} : BatchCfg{16, 3, 32}; }

static size_t batch_smem_bytes(int H, const BatchCfg& k) {
    return (size_t)k.NS * mega::SLOT_BYTES + 2 * (size_t)k.KVK * 128 * 4 +
           ((size_t)k.NB * H + k.NB * megab::MAXROWS + k.NB * 128 + mega::NCONS_WARPS * 32 + mega::NCONS_WARPS * k.NB + k.NB +
            2 * mega::NCONS_WARPS * k.NB + 3 * k.NB + 8) * 4 +
           mega::MAX_LAYERS * sizeof(DecLayerW) + (2 * mega::NSLOT_MAX + 4) * 8 + 128;
}

// `ctx` = upper bound of (position + 1) over the batch for this step
bool decode_batch_supported(const Model& m, int B, int ctx) {
    const asrb_dims& c = m.d.c;
    if (B < 2 || c.head_dim != 128) return false;
    if (c.num_attention_heads != 2 * c.num_key_value_heads) return false;
    if (!(bdims_match<1024, 2048, 3072>(c) || bdims_match<256, 512, 512>(c))) return false;
    if (c.num_hidden_layers > 32) return false;
    const int G = m.ctx->sm_count;
    if ((c.hidden_size + G - 1) / G > megab::MAXROWS) return false;
    const BatchCfg k = batch_cfg(std::min(B, 16));
    if (k.NB * c.num_key_value_heads > G) return false;                    // one merging CTA per (sequence, kv head)
    if ((ctx + k.KVK - 1) / k.KVK > megab::MAXSPLIT) return false;
    if (m.ctx->smem_optin < batch_smem_bytes(c.hidden_size, k)) return false;
    return true;
}

// floats of session scratch the batched step needs (tagged exchange buffers, 2 floats per word)
size_t decode_batch_part_floats(const Model& m) {
    const asrb_dims& c = m.d.c;
    const size_t NBm = 16;
    const size_t words = NBm * ((size_t)m.d.qkv_dim + m.d.q_dim + c.hidden_size + c.intermediate_size) +
                         NBm * c.num_key_value_heads * megab::MAXSPLIT * 2 * mega::PSTRIDE + 64;
    return 2 * words + 64;
}

void launch_decode_step_batch(const Model& m, const DecodeBufs& b, int B, float* kcache, float* vcache,
                              size_t cache_layer_stride, size_t cache_seq_stride, int max_ctx, int ctx_now, const MegaBufs& mb,
                              cudaStream_t st, int64_t* launches) {
    ASRB_REQUIRE(decode_batch_supported(m, B, ctx_now), ASRB_ERR_STATE, "batched fused decode step not supported for this model/batch/context");
    ASRB_REQUIRE(m.d_dec_layers && mb.bar && mb.part, ASRB_ERR_STATE, "fused decode step buffers missing");
    const asrb_dims& c = m.d.c;
    const int G = m.ctx->sm_count;
    for (int b0 = 0; b0 < B; b0 += 16) {             // passes of up to 16 sequences (weights are streamed once per pass)
        const int nb = std::min(16, B - b0);
        const BatchCfg k = batch_cfg(nb);
        const size_t smem = batch_smem_bytes(c.hidden_size, k);
        const void* fn = nullptr;
        if (bdims_match<1024, 2048, 3072>(c))
            fn = k.NB == 8 ? (const void*)megab::decode_batch_kernel<1024, 2048, 3072, 8, 3, 64>
                           : (const void*)megab::decode_batch_kernel<1024, 2048, 3072, 16, 3, 32>;
        else
            fn = k.NB == 8 ? (const void*)megab::decode_batch_kernel<256, 512, 512, 8, 3, 64>
                           : (const void*)megab::decode_batch_kernel<256, 512, 512, 16, 3, 32>;
        ASRB_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        megab::Params p{};
        p.layers = m.d_dec_layers; p.lm_head = m.lm_head; p.embed = m.embed; p.final_norm = m.final_norm_sw;
        p.rope_cos = m.rope_cos; p.rope_sin = m.rope_sin; p.eps = (float)c.rms_norm_eps;
        p.L = c.num_hidden_layers; p.H = c.hidden_size; p.QD = m.d.q_dim; p.KVD = m.d.kv_dim; p.I = c.intermediate_size;
        p.V = c.vocab_size; p.nkv = c.num_key_value_heads; p.nb = nb;
        p.x = b.x + (size_t)b0 * c.hidden_size;
        p.kcache = kcache + (size_t)b0 * cache_seq_stride; p.vcache = vcache + (size_t)b0 * cache_seq_stride;
        p.cache_layer_stride = cache_layer_stride; p.cache_seq_stride = cache_seq_stride; p.max_ctx = max_ctx;
        p.part_val = b.part_val + (size_t)b0 * b.n_part; p.part_idx = b.part_idx + (size_t)b0 * b.n_part; p.n_part = b.n_part;
        p.pos = b.pos + b0; p.done = b.done + b0; p.next_id = b.next_id + b0;
        p.ids_out = b.ids_out + (size_t)b0 * b.max_new; p.n_out = b.n_out + b0; p.max_new = b.max_new;
        p.bar = mb.bar;
        uint2* w = reinterpret_cast<uint2*>(mb.part);            // 16-byte aligned sub-buffers (even word counts)
        p.qkv_ll = w; w += (size_t)16 * m.d.qkv_dim;
        p.attn_ll = w; w += (size_t)16 * m.d.q_dim;
        p.x_ll = w; w += (size_t)16 * c.hidden_size;
        p.act_ll = w; w += (size_t)16 * c.intermediate_size;
        p.part_ll = w;
        if (mb.steps_issued && ++*mb.steps_issued >= 0xFFFF00u) {   // tags must stay monotonic: wipe long before the epoch wraps
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

}  // namespace asrb
