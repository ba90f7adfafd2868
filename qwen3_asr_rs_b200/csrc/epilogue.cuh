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
        case EPI_CONV_FEAT: {     // conv3 -> [chunk*OW + ow][c*OH + oh]  (permute(0,3,1,2).reshape)
            int per = e.OH * e.OW;
            int chunk = m / per, r = m - chunk * per;
            int oh = r / e.OW, ow = r - oh * e.OW;
            size_t row = (size_t)chunk * e.OW + ow;
            v0 = gelu_erf(v0 + e.bias[n]);
            store_split3(e.out_s3, e.s3_plane_stride, row * e.lds + (size_t)n * e.OH + oh, v0);
            if (has1) {
                v1 = gelu_erf(v1 + e.bias[n + 1]);
                store_split3(e.out_s3, e.s3_plane_stride, row * e.lds + (size_t)(n + 1) * e.OH + oh, v1);
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
