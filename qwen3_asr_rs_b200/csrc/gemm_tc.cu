// gemm_tc.cu -- tcgen05 tensor-core GEMM for the encoder / prefill contractions.
//
//   D[m][n] = sum_p sum_k A_p(m,k) * W[n][k]        p = bf16 split planes of the fp32 activation
//
// Replaces every Linear::forward (layers.rs:74-80) with M > 1 and conv2d2/conv2d3 (audio_encoder.rs:
// 128-129) as implicit GEMM.  B200-native structure: one 128x128 output tile per CTA, operands
// staged by TMA (cp.async.bulk.tensor, 128B swizzle) into a 3-stage shared-memory ring, a single
// elected thread issues tcgen05.mma (kind::f16, M=128 N=128 K=16) accumulating in TMEM (fp32),
// tcgen05.commit releases ring slots / signals the epilogue, eight epilogue warps read the
// accumulator with tcgen05.ld and apply the fused epilogues of epilogue.cuh.
//
// Precision: weights are exact bf16; each fp32 activation is stored as 3 bf16 planes
// (hi + mid + lo == x to 1 ulp, common.cuh).  bf16 x bf16 products are exact in fp32 and the
// accumulator is fp32, so the result matches an fp32 GEMM to accumulation-order noise -- that is
// what keeps greedy token ids identical to the fp32 oracle.  planes = 1 gives the plain bf16 GEMM.
//
// The conv A-operand is never materialised (no im2col): the activation lives in a parity-split
// channels-last layout so that each of the 9 filter taps is a plain (unit-stride) TMA box of a
// 5-D tensor map; out-of-bounds coordinates (the conv padding) are zero-filled by TMA.
#include <cuda.h>
#include <algorithm>
#include <cstring>
#include <map>
#include <tuple>
#include "internal.h"
#include "epilogue.cuh"

namespace asrb {
namespace tc {

static constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3;
static constexpr int TILE_A_BYTES = BM * BK * 2;      // 16 KB per plane
static constexpr int TILE_B_BYTES = BN * BK * 2;      // 16 KB
static constexpr int STAGE_BYTES = 3 * TILE_A_BYTES + TILE_B_BYTES;   // 64 KB
static constexpr int NTHREADS = 320;                  // warp0 TMA, warp1 MMA, warps 2-9 epilogue
static constexpr int EPI_WARPS = 8, EC = 64;          // epilogue warps, accumulator columns per epilogue warp
static constexpr int NACC = 4;                         // 128-column fp32 TMEM accumulators in rotation
static constexpr int TMEM_COLS = NACC * 128;           // all 512 columns: the MMA warp may run a whole short-K tile ahead of the epilogue
static constexpr int SLAB_LD = 20;                     // floats per staged epilogue row (16 + 4: 16-byte aligned, conflict-free)
static constexpr int CH = 4;                          // k-blocks (of 64) accumulated inside the tensor core per chunk

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar)) : "memory");
}
// K-major, 128B-swizzled operand tile: rows of 64 bf16 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, K-major both, N=128, M=128
__device__ __forceinline__ uint32_t make_idesc() {
    uint32_t d = 0;
    d |= 1u << 4;                  // c_format = F32
    d |= 1u << 7;                  // a_format = BF16
    d |= 1u << 10;                 // b_format = BF16
    d |= (uint32_t)(BN >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct ConvGeom { int OH, OW, box_h, tiles_per_chunk, kblk_per_tap, a_box_bytes; };

// A_MODE 0: plain [planes][M][K];  1: conv taps over the parity layout
//
// Accumulation: tcgen05 adds products into the fp32 TMEM accumulator with round-toward-zero, which
// biases long-K sums (measured on the 0.6B model: logits 1.3e-4 rel vs 1.3e-5 for an fp32 FMA GEMM).
// So the tensor core only accumulates CH k-blocks (K = 256) at a time into one of two TMEM
// accumulators; the epilogue warps drain each finished chunk with tcgen05.ld and add it into fp32
// registers (round-to-nearest) while the MMA warp fills the other accumulator.
// ASRB_GEMM_DEBUG timeline: CTAs 0 and 80 record clock64 stamps [cta][role: 0 TMA, 1 MMA, 2 epilogue][item < 8][4]
#define DBG_ON (E.dbg != nullptr && (blockIdx.x == 0 || blockIdx.x == 80))
#define DBG_VAL(role, it, k, v) do { if (DBG_ON && (it) < 8) E.dbg[(((blockIdx.x ? 1 : 0) * 3 + (role)) * 8 + (it)) * 4 + (k)] = (v); } while (0)
#define DBG_STAMP(role, it, k) DBG_VAL(role, it, k, clock64())

template <int A_MODE, int EPI_MODE>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               int M, int N, int K, int nplanes, int tiles_m, int tiles_n, int splits, int nacc, ConvGeom cg, GemmEpi E) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* acc_full = empty + STAGES;      // [NACC]
    uint64_t* acc_empty = acc_full + NACC;    // [NACC]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + NACC);
    float* epi_slab = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256);   // EPI_WARPS x [32][SLAB_LD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Persistent CTAs: work item = (m tile, n tile, k split), dealt round-robin (item = blockIdx.x + i * gridDim.x; n
    // fastest so that concurrently running CTAs share the A tile in L2).  The three roles keep GLOBAL stage / chunk
    // counters across items, so the TMA producer and the MMA issuer run ahead into the next tile while the epilogue
    // warps finish the previous one (its final stores overlap the next tile's loads and MMAs through the second TMEM
    // accumulator); barriers and TMEM are set up once per CTA.
    // split-K: `splits` items share one output tile, each contracts num_kb k-blocks starting at kb0 and writes its
    // partial tile to row block z of a [splits][M][N] fp32 workspace (plain epilogue, see launch_gemm_tc)
    const int num_kb = K / BK / splits;
    const int n_items = tiles_m * tiles_n * splits;

    if (threadIdx.x == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < NACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    }
    if (warp == 1) {   // TMEM allocation is warp-collective
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // item -> tile coordinates
    auto item_coords = [&](int item, int& n0, int& m0, int& chunk, int& oh0, int& z) {
        const int nt = item % tiles_n; int rest = item / tiles_n;
        const int mt = rest % tiles_m; z = rest / tiles_m;
        n0 = nt * BN; m0 = mt * BM; chunk = 0; oh0 = 0;
        if (A_MODE == 1) { chunk = mt / cg.tiles_per_chunk; oh0 = (mt % cg.tiles_per_chunk) * cg.box_h; }
    };

    if (warp == 0 && lane == 0) {
        // ================= TMA producer =================
        // TMA always delivers (and counts) the full box, zero-filling out-of-bounds elements; the conv box
        // has OW*box_h (<= 128) rows, the remaining rows of the UMMA tile are never read back.
        const uint32_t a_bytes = (A_MODE == 0) ? (uint32_t)TILE_A_BYTES : (uint32_t)cg.a_box_bytes;
        const uint32_t stage_tx = (uint32_t)nplanes * a_bytes + (uint32_t)TILE_B_BYTES;
        uint32_t kg = 0;                                   // global k-block counter (ring position)
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int n0, m0, chunk, oh0, z;
            item_coords(item, n0, m0, chunk, oh0, z);
            const int kb0 = z * num_kb;
            DBG_STAMP(0, it, 0);
            for (int kb = 0; kb < num_kb; ++kb, ++kg) {
                const int s = kg % STAGES; const uint32_t par = (kg / STAGES) & 1;
                mbar_wait(&empty[s], par ^ 1);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_expect_tx(&full[s], stage_tx);
                if (A_MODE == 0) {
                    for (int p = 0; p < nplanes; ++p) tma_load_3d(st + p * TILE_A_BYTES, &mapA, (kb0 + kb) * BK, m0, p, &full[s]);
                } else {
                    const int tap = kb / cg.kblk_per_tap, cb = kb % cg.kblk_per_tap;
                    const int kh = tap / 3, kw = tap % 3;
                    const int ph = (kh == 1) ? 0 : 1, pw = (kw == 1) ? 0 : 1;
                    const int h = oh0 + (kh == 0 ? -1 : 0), w = (kw == 0 ? -1 : 0);
                    for (int p = 0; p < nplanes; ++p)
                        tma_load_5d(st + p * TILE_A_BYTES, &mapA, cb * BK, w, h, (chunk * 2 + ph) * 2 + pw, p, &full[s]);
                }
                tma_load_2d(st + 3 * TILE_A_BYTES, &mapB, (kb0 + kb) * BK, n0, &full[s]);
            }
            DBG_STAMP(0, it, 1);
        }
    } else if (warp == 1 && lane == 0) {
        // ================= MMA issuer =================
        const uint32_t idesc = make_idesc();
        uint32_t kg = 0, cgl = 0;                          // global k-block / chunk counters
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            long long w_acc = 0, w_full = 0;
            DBG_STAMP(1, it, 0);
            for (int kb = 0; kb < num_kb; ++kb, ++kg) {
                const int s = kg % STAGES; const uint32_t par = (kg / STAGES) & 1;
                const int cb = cgl % nacc;                                  // accumulator buffer of the current chunk
                const bool chunk_first = (kb % CH) == 0;
                long long tw0 = DBG_ON ? clock64() : 0;
                if (chunk_first) { mbar_wait(&acc_empty[cb], ((cgl / nacc) & 1) ^ 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
                long long tw1 = DBG_ON ? clock64() : 0;
                mbar_wait(&full[s], par);
                if (DBG_ON) { w_acc += tw1 - tw0; w_full += clock64() - tw1; }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
                const uint64_t bdesc = make_smem_desc(sa + 3 * TILE_A_BYTES);
                const uint32_t tacc = tmem_base + (uint32_t)(cb * BN);
                for (int p = 0; p < nplanes; ++p) {
                    const uint64_t adesc = make_smem_desc(sa + p * TILE_A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)     // +32 B per K=16 step inside the 128 B swizzle atom
                        umma(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, !(chunk_first && p == 0 && k == 0));
                }
                umma_commit(&empty[s]);                   // slot reusable once these MMAs retire
                if ((kb % CH) == CH - 1 || kb == num_kb - 1) { umma_commit(&acc_full[cb]); ++cgl; }
            }
            DBG_STAMP(1, it, 1); DBG_VAL(1, it, 2, w_acc); DBG_VAL(1, it, 3, w_full);
        }
    } else if (warp >= 2) {
        // ================= epilogue =================
        // 8 warps: warp w reads TMEM lane quadrant w & 3 (its 32 accumulator rows) and owns EC = 64 of the tile's 128
        // columns.  (With 4 warps x 128 columns the GELU / split3 epilogues ran 29 k cycles per tile on one warp per
        // scheduler, longer than a K = 896 main loop; two warps per scheduler halve that and overlap each other's latency.)
        const int quad = warp & 3;
        const int hb = ((warp - 2) >> 2) * EC;        // first tile column of this warp
        const int num_chunks = (num_kb + CH - 1) / CH;
        uint32_t cgl = 0;
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
            int n0, m0, chunk, oh0, z;
            item_coords(item, n0, m0, chunk, oh0, z);
            long long w_wait = 0;
            float accum[EC];
#pragma unroll
            for (int j = 0; j < EC; ++j) accum[j] = 0.f;
            for (int c = 0; c < num_chunks; ++c, ++cgl) {
                const int cb = cgl % nacc;
                long long tw0 = DBG_ON ? clock64() : 0;
                mbar_wait(&acc_full[cb], (cgl / nacc) & 1);
                if (DBG_ON) { w_wait += clock64() - tw0; if (c == 0 && threadIdx.x == 64) DBG_STAMP(2, it, 0); }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c0 = 0; c0 < EC; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cb * BN + hb + c0), v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) accum[c0 + j] += __uint_as_float(v[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&acc_empty[cb])) : "memory");
            }
            if (threadIdx.x == 64) { DBG_STAMP(2, it, 1); DBG_VAL(2, it, 3, w_wait); }
            // Store through a per-warp shared-memory transpose.  A thread owns one accumulator ROW (its TMEM lane), so
            // storing from the registers directly makes every store instruction touch 32 different rows.  Instead each
            // warp stages 16 columns of its 32 rows ([32][20] floats, conflict-free for 128-bit accesses), then lane l
            // re-reads columns 4*(l&3).. of rows 8*i + (l>>2): a warp store instruction covers 8 rows x 64 contiguous bytes.
            float* slab = epi_slab + (warp - 2) * (32 * SLAB_LD);
            const int lr = lane >> 2, lc = (lane & 3) * 4;
#pragma unroll
            for (int c0 = 0; c0 < EC; c0 += 16) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(slab + lane * SLAB_LD + j) = make_float4(accum[c0 + j], accum[c0 + j + 1], accum[c0 + j + 2], accum[c0 + j + 3]);
                __syncwarp();
                const int n = n0 + hb + c0 + lc;
                const bool ncol = n < N;
                const float4 bias4 = ncol ? epi_bias4<EPI_MODE>(E, n) : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 v[4]; long long mr[4]; EpiIn in[4];            // 4 rows per lane: every global load first, then the stores
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = quad * 32 + 8 * i + lr;            // row inside the tile
                    v[i] = *reinterpret_cast<const float4*>(slab + (8 * i + lr) * SLAB_LD + lc);
                    long long m = -1;
                    if (A_MODE == 0) { if (m0 + rr < M) m = m0 + rr + (long long)z * M; }
                    else {
                        const int oh = oh0 + rr / cg.OW, ow = rr % cg.OW;
                        if (rr < cg.box_h * cg.OW && oh < cg.OH) m = ((long long)chunk * cg.OH + oh) * cg.OW + ow;
                    }
                    mr[i] = (m >= 0 && ncol) ? m : -1;
                    if (mr[i] >= 0) in[i] = epi_fetch4<EPI_MODE>(E, N, (int)mr[i], n);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (mr[i] >= 0) epi_store4<EPI_MODE>(E, N, (int)mr[i], n, v[i], bias4, in[i]);
                __syncwarp();
            }
            if (threadIdx.x == 64) DBG_STAMP(2, it, 2);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// split-K second pass: sum the partial tiles in a fixed order, then the GEMM's own epilogue (bias / GELU / residual / split3)
template <int MODE>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int S, int M, int N, GemmEpi E) {
    const int n8 = N / 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * n8) return;
    const int m = idx / n8, n = (idx - m * n8) * 8;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
    for (int sp = 0; sp < S; ++sp) {
        const float* src = ws + ((size_t)sp * M + m) * N + n;
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    epi_store8<MODE>(E, N, m, n, v);
}

// ---- host: tensor maps ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        ASRB_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
        ASRB_REQUIRE(p != nullptr && qr == cudaDriverEntryPointSuccess, ASRB_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
        fn = (EncodeTiledFn)p;
    }
    return fn;
}
static CUtensorMap make_map(const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    CUtensorMap m;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw Error(ASRB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    return m;
}

// Tensor maps depend only on (base, shape, strides, box): encode once per distinct operand instead of on every launch
// (a clip issues ~400 GEMMs over a few dozen distinct operands).  Thread-local: sessions on different host threads.
struct MapKey {
    const void* base; int rank; cuuint64_t d[5]; cuuint64_t s[4]; cuuint32_t b[5];
    bool operator<(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) < 0; }
};
static const CUtensorMap& cached_map(const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, const cuuint32_t* box) {
    static thread_local std::map<MapKey, CUtensorMap> cache;
    MapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.rank = rank;
    for (int i = 0; i < rank; ++i) { k.d[i] = dims[i]; k.b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) k.s[i] = strides_bytes[i];
    auto it = cache.find(k);
    if (it == cache.end()) {
        if (cache.size() > 4096) cache.clear();
        it = cache.emplace(k, make_map(base, rank, dims, strides_bytes, box)).first;
    }
    return it->second;
}

static int sm_count() {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
}
static size_t smem_bytes() { return (size_t)STAGES * STAGE_BYTES + 1024 + 256 + EPI_WARPS * 32 * SLAB_LD * 4; }   // stages + alignment slack + barriers + epilogue slabs

}  // namespace tc

// ASRB_GEMM_DEBUG="M,N,K": the next matching launch records the per-role timeline and prints it (debug builds of a trip)
static void gemm_debug_dump(long long* d_dbg, int M, int N, int K, cudaStream_t st) {
    long long h[2 * 3 * 8 * 4];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[gemm_tc timeline] M=%d N=%d K=%d (cycles relative to the CTA's first stamp)\n", M, N, K);
    for (int c = 0; c < 2; ++c) {
        const long long t0 = h[((c * 3 + 0) * 8 + 0) * 4 + 0];
        for (int it = 0; it < 8; ++it) {
            const long long* p = &h[((c * 3 + 0) * 8 + it) * 4]; const long long* q = &h[((c * 3 + 1) * 8 + it) * 4];
            const long long* e = &h[((c * 3 + 2) * 8 + it) * 4];
            if (!p[0]) break;
            fprintf(stderr, "  cta%d item%d  tma %lld..%lld | mma %lld..%lld wait_acc %lld wait_full %lld | epi first %lld drained %lld stored %lld wait %lld\n",
                    c ? 80 : 0, it, p[0] - t0, p[1] - t0, q[0] - t0, q[1] - t0, q[2], q[3], e[0] - t0, e[1] - t0, e[2] - t0, e[3]);
        }
    }
}

bool launch_gemm_tc(const GemmA& A, const bf16* W, int N, const GemmEpi& Ein, cudaStream_t st) {
    using namespace tc;
    GemmEpi E = Ein;
    static long long* d_dbg = nullptr; static int dbgM = -1, dbgN = 0, dbgK = 0, dbg_left = 0;
    if (dbgM == -1) {
        dbgM = 0;
        if (const char* e = getenv("ASRB_GEMM_DEBUG")) {
            if (sscanf(e, "%d,%d,%d", &dbgM, &dbgN, &dbgK) == 3) { cudaMalloc(&d_dbg, 2 * 3 * 8 * 4 * 8); dbg_left = 2; } else dbgM = 0;
        }
    }
    const bool dbg_this = dbg_left > 0 && A.M == dbgM && N == dbgN && A.K == dbgK;
    if (dbg_this) { cudaMemsetAsync(d_dbg, 0, 2 * 3 * 8 * 4 * 8, st); E.dbg = d_dbg; --dbg_left; }
    struct Dump { bool on; long long* d; int M, N, K; cudaStream_t st; ~Dump() { if (on) gemm_debug_dump(d, M, N, K, st); } } dump{dbg_this, d_dbg, A.M, N, A.K, st};
    // ASRB_GEMM_TIME=1: CUDA-event time of every launch (synchronises per launch; measurement runs only)
    static const bool time_all = getenv("ASRB_GEMM_TIME") != nullptr;
    struct Timer {
        bool on; cudaEvent_t e0, e1; int M, N, K, mode; cudaStream_t st;
        Timer(bool o, int M_, int N_, int K_, int mode_, cudaStream_t s) : on(o), M(M_), N(N_), K(K_), mode(mode_), st(s) {
            if (on) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
        }
        ~Timer() {
            if (!on) return;
            cudaEventRecord(e1, st); cudaEventSynchronize(e1);
            float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
            fprintf(stderr, "[gemm_tc time] M=%d N=%d K=%d epi=%d %.1f us\n", M, N, K, mode, ms * 1000.f);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
    } timer(time_all || dbg_this, A.M, N, A.K, E.mode, st);
    if (A.K % BK != 0 || A.M <= 0 || N <= 0) return false;
    if (A.nplanes < 1 || A.nplanes > 3) return false;
    if ((reinterpret_cast<uintptr_t>(A.a) & 15) || (reinterpret_cast<uintptr_t>(W) & 15)) return false;
    // B: [N][K] row-major bf16
    cuuint64_t bd[2] = {(cuuint64_t)A.K, (cuuint64_t)N};
    cuuint64_t bs[1] = {(cuuint64_t)A.K * 2};
    cuuint32_t bb[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
    const CUtensorMap mapB = cached_map(W, 2, bd, bs, bb);
    ConvGeom cg{};
    const size_t smem = smem_bytes();
    static const int nacc = [] { const char* e = getenv("ASRB_GEMM_NACC"); int v = e ? atoi(e) : NACC; return (v == 2 || v == 4) ? v : NACC; }();
    if (N % 8 != 0) return false;
    if (A.mode == A_PLAIN) {
        if (A.plane_stride % 8 != 0 && A.nplanes > 1) return false;
        cuuint64_t ad[3] = {(cuuint64_t)A.K, (cuuint64_t)A.M, (cuuint64_t)3};
        cuuint64_t as[2] = {(cuuint64_t)A.lda * 2, (cuuint64_t)A.plane_stride * 2};
        cuuint32_t ab[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
        const CUtensorMap mapA = cached_map(A.a, 3, ad, as, ab);
        const int tiles_n = (N + BN - 1) / BN, tiles_m = (A.M + BM - 1) / BM;
        const int sms = sm_count();
        // Split-K for plain GEMMs that cannot fill the GPU (e.g. prefill o_proj / down_proj: 32 tiles, K = 2048 / 3072;
        // encoder fc2: 28 tiles, K = 3584): such a CTA is bound by its own TMA load rate (64 KB of operands per k-block),
        // so 2-4 CTAs per tile finish 2-4x sooner; a second pass sums the partial tiles (fixed order) and applies the
        // epilogue.  Deterministic; costs one extra fp32 round trip of the tile through L2.
        const int tiles = tiles_n * tiles_m, kblocks = A.K / BK;
        int splits = 1;
        if ((E.mode == EPI_PLAIN || E.mode == EPI_CONVOUT) && E.splitk_ws && tiles <= 64 && N % 8 == 0 && (size_t)4 * A.M * N <= SPLITK_WS_FLOATS) {
            for (int sp = 4; sp >= 2; --sp)
                if (kblocks % sp == 0 && kblocks / sp >= 6 && tiles * sp <= 160) { splits = sp; break; }
        }
        if (splits > 1) {
            GemmEpi P;                               // partial tiles: plain fp32 rows [split][M][N]
            P.out_f32 = E.splitk_ws; P.ldo = N;
            // (the attribute is per device: set on every launch, a process may hold contexts on several GPUs)
            ASRB_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<0, EPI_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            gemm_tc_kernel<0, EPI_PLAIN><<<std::min(tiles * splits, sms), NTHREADS, smem, st>>>(mapA, mapB, A.M, N, A.K, A.nplanes, tiles_m, tiles_n, splits, nacc, cg, P);
            const int work = A.M * (N / 8);
            if (E.mode == EPI_PLAIN) splitk_reduce_kernel<EPI_PLAIN><<<(work + 255) / 256, 256, 0, st>>>(E.splitk_ws, splits, A.M, N, E);
            else splitk_reduce_kernel<EPI_CONVOUT><<<(work + 255) / 256, 256, 0, st>>>(E.splitk_ws, splits, A.M, N, E);
            ASRB_CUDA_CHECK(cudaGetLastError());
            if (E.extra_launches) *E.extra_launches += 1;
            return true;
        }
#define ASRB_TC_LAUNCH(AM, EM)                                                                                         \
    {                                                                                                                  \
        ASRB_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<AM, EM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        gemm_tc_kernel<AM, EM><<<std::min(tiles_m * tiles_n, sms), NTHREADS, smem, st>>>(mapA, mapB, A.M, N, A.K, A.nplanes, tiles_m, tiles_n, 1, nacc, cg, E); \
    }
        if (E.mode == EPI_PLAIN) ASRB_TC_LAUNCH(0, EPI_PLAIN)
        else if (E.mode == EPI_SWIGLU) ASRB_TC_LAUNCH(0, EPI_SWIGLU)
        else if (E.mode == EPI_CONVOUT) ASRB_TC_LAUNCH(0, EPI_CONVOUT)
        else return false;
    } else {
        if (A.cpad % BK != 0 || A.OW > BM) return false;
        const int per = A.OH * A.OW;
        const int chunks = A.M / per;
        cg.OH = A.OH; cg.OW = A.OW; cg.box_h = std::min(A.OH, BM / A.OW);
        cg.tiles_per_chunk = (A.OH + cg.box_h - 1) / cg.box_h; cg.kblk_per_tap = A.cpad / BK;
        cg.a_box_bytes = A.OW * cg.box_h * BK * 2;
        // [plane][chunk*4 + ph*2 + pw][Hh][Wh][cpad]
        cuuint64_t ad[5] = {(cuuint64_t)A.cpad, (cuuint64_t)A.Wh, (cuuint64_t)A.Hh, (cuuint64_t)chunks * 4, 3};
        cuuint64_t as[4] = {(cuuint64_t)A.cpad * 2, (cuuint64_t)A.Wh * A.cpad * 2, (cuuint64_t)A.Hh * A.Wh * A.cpad * 2,
                            (cuuint64_t)A.plane_stride * 2};
        cuuint32_t ab[5] = {(cuuint32_t)BK, (cuuint32_t)A.OW, (cuuint32_t)cg.box_h, 1, 1};
        const CUtensorMap mapA = cached_map(A.a, 5, ad, as, ab);
        const int tiles_n = (N + BN - 1) / BN, tiles_m = chunks * cg.tiles_per_chunk;
        const int sms = sm_count();
        if (E.mode == EPI_CONV_PARITY) ASRB_TC_LAUNCH(1, EPI_CONV_PARITY)
        else if (E.mode == EPI_CONV_FEAT) ASRB_TC_LAUNCH(1, EPI_CONV_FEAT)
        else return false;
#undef ASRB_TC_LAUNCH
    }
    ASRB_CUDA_CHECK(cudaGetLastError());
    return true;
}

}  // namespace asrb
