// epilogue.cuh -- GEMM epilogues shared by the SIMT and tcgen05 GEMMs.
// Fuses what the reference does as separate ATen calls after each matmul:
//   bias add (layers.rs:76-79), exact-erf GELU (audio_encoder.rs:127-129, layers.rs:193),
//   residual add (layers.rs:235,241,454,460), SiLU(gate)*up (layers.rs:396-399),
//   NCHW->[C,t,c*f] permute (audio_encoder.rs:132-133), positional add + valid-token gather
//   (audio_encoder.rs:137-149).
#pragma once
#include "internal.h"

namespace asrb {

__device__ __forceinline__ void epi_write(const GemmEpi& e, size_t f32_idx, size_t s3_idx, float v) {
    if (e.out_f32) e.out_f32[f32_idx] = v;
    if (e.out_s3) store_split3(e.out_s3, e.s3_plane_stride, s3_idx, v);
}

// one accumulator pair (m, n) and (m, n+1), n even; has1 = column n+1 exists
__device__ __forceinline__ void epi_store2(const GemmEpi& e, int N, int m, int n, float v0, float v1, bool has1) {
    switch (e.mode) {
        case EPI_PLAIN: {
            if (e.bias) { v0 += e.bias[n]; if (has1) v1 += e.bias[n + 1]; }
            if (e.act == 1) { v0 = gelu_erf(v0); v1 = gelu_erf(v1); }
            if (e.residual) {
                v0 += e.residual[(size_t)m * e.ldr + n];
                if (has1) v1 += e.residual[(size_t)m * e.ldr + n + 1];
            }
            epi_write(e, (size_t)m * e.ldo + n, (size_t)m * e.lds + n, v0);
            if (has1) epi_write(e, (size_t)m * e.ldo + n + 1, (size_t)m * e.lds + n + 1, v1);
            break;
        }
        case EPI_SWIGLU: {   // rows interleaved: even = gate_j, odd = up_j
            float o = silu(v0) * v1;
            int j = n >> 1;
            epi_write(e, (size_t)m * e.ldo + j, (size_t)m * e.lds + j, o);
            break;
        }
        case EPI_CONV_PARITY: {   // conv2 -> parity-split channels-last input of conv3
            int per = e.OH * e.OW;
            int chunk = m / per, r = m - chunk * per;
            int oh = r / e.OW, ow = r - oh * e.OW;
            size_t base = ((((size_t)chunk * 2 + (oh & 1)) * 2 + (ow & 1)) * e.Hh2 + (oh >> 1)) * e.Wh2 + (ow >> 1);
            base = base * e.cpad + n;
            v0 = gelu_erf(v0 + e.bias[n]);
            store_split3(e.out_s3, e.s3_plane_stride, base, v0);
            if (has1) {
                v1 = gelu_erf(v1 + e.bias[n + 1]);
                store_split3(e.out_s3, e.s3_plane_stride, base + 1, v1);
            }
            break;
        }
        case EPI_CONV_FEAT: {     // conv3 -> [chunk*OW + ow][oh*N + c]: permute(0,3,1,2).reshape with the feature index
                                  // transposed (c*OH + oh -> oh*N + c; conv_out.weight is permuted to match, model.cu)
            int per = e.OH * e.OW;
            int chunk = m / per, r = m - chunk * per;
            int oh = r / e.OW, ow = r - oh * e.OW;
            size_t row = (size_t)chunk * e.OW + ow;
            v0 = gelu_erf(v0 + e.bias[n]);
            store_split3(e.out_s3, e.s3_plane_stride, row * e.lds + (size_t)oh * N + n, v0);
            if (has1) {
                v1 = gelu_erf(v1 + e.bias[n + 1]);
                store_split3(e.out_s3, e.s3_plane_stride, row * e.lds + (size_t)oh * N + n + 1, v1);
            }
            break;
        }
        case EPI_CONVOUT: {       // + pos[t] then keep only valid tokens of each chunk
            int tok = e.row_map[m];
            if (tok < 0) break;
            int t = m % e.pos_period;
            if (e.bias) { v0 += e.bias[n]; if (has1) v1 += e.bias[n + 1]; }
            v0 += e.pos[(size_t)t * N + n];
            e.out_f32[(size_t)tok * e.ldo + n] = v0;
            if (has1) {
                v1 += e.pos[(size_t)t * N + n + 1];
                e.out_f32[(size_t)tok * e.ldo + n + 1] = v1;
            }
            break;
        }
    }
}

// A-operand fetch used by the SIMT GEMM (the tcgen05 GEMM expresses the same addressing as TMA
// tensor-map coordinates).  Returns hi+mid+lo of the split3 planes.
__device__ __forceinline__ float load_a(const GemmA& A, int m, int k) {
    if (A.mode == A_PLAIN) return load_split3(A.a, A.plane_stride, (size_t)m * A.lda + k, A.nplanes);
    int per = A.OH * A.OW;
    int chunk = m / per, r = m - chunk * per;
    int oh = r / A.OW, ow = r - oh * A.OW;
    int tap = k / A.cpad, cin = k - tap * A.cpad;
    int kh = tap / 3, kw = tap - kh * 3;
    int ph = (kh == 1) ? 0 : 1, pw = (kw == 1) ? 0 : 1;
    int hh = oh + (kh == 0 ? -1 : 0), wh = ow + (kw == 0 ? -1 : 0);
    if (hh < 0 || wh < 0) return 0.f;                             // padding = 1
    size_t idx = ((((size_t)chunk * 2 + ph) * 2 + pw) * A.Hh + hh) * A.Wh + wh;
    return load_split3(A.a, A.plane_stride, idx * A.cpad + cin, A.nplanes);
}

}  // namespace asrb

// ---------------------------------------------------------------------------------------------
// Vectorised epilogue for the tensor-core GEMM: one thread owns 8 consecutive accumulator columns
// n..n+7 (n % 8 == 0) of row m.  Same arithmetic as epi_store2; stores are 128-bit where the layout
// allows (fp32 rows: 2 x float4, split3 planes: one uint4 of 8 bf16 per plane).
// ---------------------------------------------------------------------------------------------
namespace asrb {

__device__ __forceinline__ void store_split3_x8(bf16* base, size_t plane_stride, size_t idx, const float* v) {
    uint32_t hi[4], mid[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        Split3 a = split3(v[2 * i]), b = split3(v[2 * i + 1]);
        hi[i] = (uint32_t)__bfloat16_as_ushort(a.hi) | ((uint32_t)__bfloat16_as_ushort(b.hi) << 16);
        mid[i] = (uint32_t)__bfloat16_as_ushort(a.mid) | ((uint32_t)__bfloat16_as_ushort(b.mid) << 16);
        lo[i] = (uint32_t)__bfloat16_as_ushort(a.lo) | ((uint32_t)__bfloat16_as_ushort(b.lo) << 16);
    }
    *reinterpret_cast<uint4*>(base + idx) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + plane_stride + idx) = make_uint4(mid[0], mid[1], mid[2], mid[3]);
    *reinterpret_cast<uint4*>(base + 2 * plane_stride + idx) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// 4 consecutive bf16 per plane = one 8-byte store per plane
__device__ __forceinline__ void store_split3_x4(bf16* base, size_t plane_stride, size_t idx, const float* v) {
    const Split3 a = split3(v[0]), b = split3(v[1]), c = split3(v[2]), d = split3(v[3]);
    auto pk = [](bf16 x, bf16 y) { return (uint32_t)__bfloat16_as_ushort(x) | ((uint32_t)__bfloat16_as_ushort(y) << 16); };
    *reinterpret_cast<uint2*>(base + idx) = make_uint2(pk(a.hi, b.hi), pk(c.hi, d.hi));
    *reinterpret_cast<uint2*>(base + plane_stride + idx) = make_uint2(pk(a.mid, b.mid), pk(c.mid, d.mid));
    *reinterpret_cast<uint2*>(base + 2 * plane_stride + idx) = make_uint2(pk(a.lo, b.lo), pk(c.lo, d.lo));
}

// Epilogue unit of the tcgen05 GEMM after its shared-memory transpose: one lane owns 4 consecutive columns n..n+3
// (n % 4 == 0) of row m, 8 lanes cover 32 consecutive columns of a row, so a warp store instruction writes whole 128-byte
// lines (fp32 rows) / whole 32-byte sectors (bf16 planes).  Same arithmetic as epi_store2 / epi_store8.
// The global READS of the epilogue (bias, residual, positional row, row map) are split from the stores: the compiler must
// keep a load behind any earlier store that may alias, so a load-inside-the-store-loop epilogue pays one L2 round trip
// per iteration (measured: 18 k cycles per 128x128 tile with a bias, 50 k with a residual).  The caller issues the
// epi_fetch4 calls of a group of rows first, then the epi_store4 calls.
struct EpiIn { float4 add; int tok; };
template <int MODE>
__device__ __forceinline__ float4 epi_bias4(const GemmEpi& e, int n) {
    const bool has = (MODE == EPI_CONV_PARITY || MODE == EPI_CONV_FEAT) ? true : (MODE == EPI_SWIGLU ? false : e.bias != nullptr);
    return has ? __ldg(reinterpret_cast<const float4*>(e.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int MODE>
__device__ __forceinline__ EpiIn epi_fetch4(const GemmEpi& e, int N, int m, int n) {
    EpiIn in; in.add = make_float4(0.f, 0.f, 0.f, 0.f); in.tok = 0;
    if (MODE == EPI_PLAIN) {
        // plain load: the residual may be the output buffer (x += proj(...)); this thread reads exactly the elements it
        // later overwrites
        if (e.residual) in.add = *reinterpret_cast<const float4*>(e.residual + (size_t)m * e.ldr + n);
    } else if (MODE == EPI_CONVOUT) {
        in.tok = __ldg(e.row_map + m);
        in.add = __ldg(reinterpret_cast<const float4*>(e.pos + (size_t)(m % e.pos_period) * N + n));
    }
    return in;
}
template <int MODE>
__device__ __forceinline__ void epi_store4(const GemmEpi& e, int N, int m, int n, float4 a, float4 b, const EpiIn& in) {
    float v[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};          // + bias (zero when there is none)
    if (MODE == EPI_PLAIN) {
        if (e.act == 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
        }
        if (e.residual) { v[0] += in.add.x; v[1] += in.add.y; v[2] += in.add.z; v[3] += in.add.w; }
        if (e.out_f32) *reinterpret_cast<float4*>(e.out_f32 + (size_t)m * e.ldo + n) = make_float4(v[0], v[1], v[2], v[3]);
        if (e.out_s3) store_split3_x4(e.out_s3, e.s3_plane_stride, (size_t)m * e.lds + n, v);
    } else if (MODE == EPI_CONV_PARITY || MODE == EPI_CONV_FEAT) {
        const int per = e.OH * e.OW;
        const int chunk = m / per, r = m - chunk * per;
        const int oh = r / e.OW, ow = r - oh * e.OW;
        size_t idx;
        if (MODE == EPI_CONV_PARITY) {
            idx = ((((size_t)chunk * 2 + (oh & 1)) * 2 + (ow & 1)) * e.Hh2 + (oh >> 1)) * e.Wh2 + (ow >> 1);
            idx = idx * e.cpad + n;
        } else {
            idx = ((size_t)chunk * e.OW + ow) * e.lds + (size_t)oh * N + n;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = gelu_erf(v[i]);
        store_split3_x4(e.out_s3, e.s3_plane_stride, idx, v);
    } else if (MODE == EPI_CONVOUT) {
        if (in.tok < 0) return;
        *reinterpret_cast<float4*>(e.out_f32 + (size_t)in.tok * e.ldo + n) =
            make_float4(v[0] + in.add.x, v[1] + in.add.y, v[2] + in.add.z, v[3] + in.add.w);
    } else {   // EPI_SWIGLU: columns (gate_j, up_j) x 2 -> 2 outputs
        const float o0 = silu(v[0]) * v[1], o1 = silu(v[2]) * v[3];
        const size_t j = (size_t)(n >> 1);
        if (e.out_f32) *reinterpret_cast<float2*>(e.out_f32 + (size_t)m * e.ldo + j) = make_float2(o0, o1);
        if (e.out_s3) {
            const Split3 a0 = split3(o0), a1 = split3(o1);
            auto pk = [](bf16 x, bf16 y) { return (uint32_t)__bfloat16_as_ushort(x) | ((uint32_t)__bfloat16_as_ushort(y) << 16); };
            const size_t idx = (size_t)m * e.lds + j;
            *reinterpret_cast<uint32_t*>(e.out_s3 + idx) = pk(a0.hi, a1.hi);
            *reinterpret_cast<uint32_t*>(e.out_s3 + e.s3_plane_stride + idx) = pk(a0.mid, a1.mid);
            *reinterpret_cast<uint32_t*>(e.out_s3 + 2 * e.s3_plane_stride + idx) = pk(a0.lo, a1.lo);
        }
    }
}

template <int MODE>
__device__ __forceinline__ void epi_store8(const GemmEpi& e, int N, int m, int n, const float* acc) {
    if (MODE == EPI_PLAIN) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[i];
        if (e.bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(e.bias + n), b1 = *reinterpret_cast<const float4*>(e.bias + n + 4);
            v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
        }
        if (e.act == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = gelu_erf(v[i]);
        }
        if (e.residual) {
            const float* r = e.residual + (size_t)m * e.ldr + n;
            const float4 r0 = *reinterpret_cast<const float4*>(r), r1 = *reinterpret_cast<const float4*>(r + 4);
            v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
        }
        if (e.out_f32) {
            float* o = e.out_f32 + (size_t)m * e.ldo + n;
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
        }
        if (e.out_s3) store_split3_x8(e.out_s3, e.s3_plane_stride, (size_t)m * e.lds + n, v);
    } else if (MODE == EPI_CONV_PARITY) {
        const int per = e.OH * e.OW;
        const int chunk = m / per, r = m - chunk * per;
        const int oh = r / e.OW, ow = r - oh * e.OW;
        size_t base = ((((size_t)chunk * 2 + (oh & 1)) * 2 + (ow & 1)) * e.Hh2 + (oh >> 1)) * e.Wh2 + (ow >> 1);
        base = base * e.cpad + n;
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(e.bias + n), b1 = *reinterpret_cast<const float4*>(e.bias + n + 4);
        v[0] = gelu_erf(acc[0] + b0.x); v[1] = gelu_erf(acc[1] + b0.y); v[2] = gelu_erf(acc[2] + b0.z); v[3] = gelu_erf(acc[3] + b0.w);
        v[4] = gelu_erf(acc[4] + b1.x); v[5] = gelu_erf(acc[5] + b1.y); v[6] = gelu_erf(acc[6] + b1.z); v[7] = gelu_erf(acc[7] + b1.w);
        store_split3_x8(e.out_s3, e.s3_plane_stride, base, v);
    } else if (MODE == EPI_CONVOUT) {
        const int tok = e.row_map[m];
        if (tok < 0) return;
        const int t = m % e.pos_period;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = acc[i] + (e.bias ? e.bias[n + i] : 0.f) + e.pos[(size_t)t * N + n + i];
        float* o = e.out_f32 + (size_t)tok * e.ldo + n;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if (MODE == EPI_SWIGLU) {   // 8 interleaved columns (gate_j, up_j) x 4 -> 4 outputs, one 8-byte store per plane
        float o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = silu(acc[2 * i]) * acc[2 * i + 1];
        const size_t idx = (size_t)m * e.lds + (n >> 1);
        if (e.out_f32) {
#pragma unroll
            for (int i = 0; i < 4; ++i) e.out_f32[(size_t)m * e.ldo + (n >> 1) + i] = o[i];
        }
        if (e.out_s3) {
            uint32_t hi[2], mid[2], lo[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                Split3 a = split3(o[2 * i]), b = split3(o[2 * i + 1]);
                hi[i] = (uint32_t)__bfloat16_as_ushort(a.hi) | ((uint32_t)__bfloat16_as_ushort(b.hi) << 16);
                mid[i] = (uint32_t)__bfloat16_as_ushort(a.mid) | ((uint32_t)__bfloat16_as_ushort(b.mid) << 16);
                lo[i] = (uint32_t)__bfloat16_as_ushort(a.lo) | ((uint32_t)__bfloat16_as_ushort(b.lo) << 16);
            }
            *reinterpret_cast<uint2*>(e.out_s3 + idx) = make_uint2(hi[0], hi[1]);
            *reinterpret_cast<uint2*>(e.out_s3 + e.s3_plane_stride + idx) = make_uint2(mid[0], mid[1]);
            *reinterpret_cast<uint2*>(e.out_s3 + 2 * e.s3_plane_stride + idx) = make_uint2(lo[0], lo[1]);
        }
    } else {   // EPI_CONV_FEAT: 8 consecutive channels of (chunk, ow, oh) are contiguous in the transposed feature layout
        const int per = e.OH * e.OW;
        const int chunk = m / per, r = m - chunk * per;
        const int oh = r / e.OW, ow = r - oh * e.OW;
        const size_t row = (size_t)chunk * e.OW + ow;
        float v[8];
        const float4 b0 = *reinterpret_cast<const float4*>(e.bias + n), b1 = *reinterpret_cast<const float4*>(e.bias + n + 4);
        v[0] = gelu_erf(acc[0] + b0.x); v[1] = gelu_erf(acc[1] + b0.y); v[2] = gelu_erf(acc[2] + b0.z); v[3] = gelu_erf(acc[3] + b0.w);
        v[4] = gelu_erf(acc[4] + b1.x); v[5] = gelu_erf(acc[5] + b1.y); v[6] = gelu_erf(acc[6] + b1.z); v[7] = gelu_erf(acc[7] + b1.w);
        store_split3_x8(e.out_s3, e.s3_plane_stride, row * e.lds + (size_t)oh * N + n, v);
    }
}

}  // namespace asrb
