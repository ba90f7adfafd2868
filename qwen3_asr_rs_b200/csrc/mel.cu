// mel.cu -- 16 kHz f32 samples -> 128-bin Whisper log-mel, fused.
//
// Replaces WhisperFeatureExtractor::extract (/root/reference/src/mel.rs:49-96): zero-pad to a
// multiple of hop (:51-53), reflect pad n_fft/2 (:63-65), periodic-Hann STFT 400/160 without
// centering (:68-76), |.|^2 (:80), drop the last frame (:83-84), filterbank matmul (:87),
// log10(clamp 1e-10) (:90), global max - 8 clamp (:91-92), (x+4)/4 (:93).
//
// One CTA = 16 frames.  The 400-point real DFT uses the even / odd symmetry of the twiddles about n = 200:
//   re[k] = x[0] + (-1)^k x[200] + sum_{n=1..199} (x[n] + x[400-n]) cos(2 pi k n / 400)
//   im[k] =                         sum_{n=1..199} (x[n] - x[400-n]) sin(2 pi k n / 400)
// i.e. half the FMAs of the direct form.  The pair sums / differences of the 16 windowed frames are staged in shared
// memory as [n][16 frames] (every thread reads them with broadcast LDS.128), the {cos, sin} table (400 entries, 3.2 KB)
// lives in shared memory too and thread k walks it with r += k (mod 400) -- the [400][208] table the first version
// streamed from L2 cost 125 MB of L2 traffic per 30 s clip, 60x the kernel's HBM bytes.  Power spectrum kept in shared
// memory (the [201, F+1] STFT magnitude is never written to HBM), triangular filters applied over their non-zero
// support only, log10 and the per-utterance running max (atomicMax on an order-preserving int key).  A second
// elementwise kernel applies the max-8 clamp and affine.
// Bound: fp32 FMA (241 MFMA per 30 s clip); HBM traffic 1.92 MB in + 1.54 MB out per clip.
#include "internal.h"

namespace asrb {

static constexpr int NFFT = 400, HOP = 160, NBIN = 201, KP = 208, FT = 16, MEL_THREADS = 224;

__global__ void __launch_bounds__(MEL_THREADS)
mel_power_kernel(const float* __restrict__ samples, const int64_t* __restrict__ soff,
                 const int64_t* __restrict__ n_true, const int64_t* __restrict__ n_pad,
                 const int64_t* __restrict__ foff, const float* __restrict__ hann,
                 const float2* __restrict__ tw,
                 const float* __restrict__ fb, const int* __restrict__ krange, int n_mels,
                 float* __restrict__ mel_out, int* __restrict__ maxkey) {
    constexpr int NH = NFFT / 2;                           // 200
    __shared__ __align__(16) float buf[(2 * NH + 1) * FT];
    float (*xe)[FT] = reinterpret_cast<float (*)[FT]>(buf);                   // xe[n] = x[n] + x[400-n] (n = 1..199); xe[0] = x[0]; xe[200] = x[200]
    float (*xo)[FT] = reinterpret_cast<float (*)[FT]>(buf + (NH + 1) * FT);   // xo[n] = x[n] - x[400-n]
    __shared__ float2 tws[NFFT];
    __shared__ float red[32];
    float (*pw)[KP] = reinterpret_cast<float (*)[KP]>(buf);                   // power spectrum [FT][KP] reuses the buffer after the DFT
    static_assert(FT * KP <= (2 * NH + 1) * FT, "pw must fit in the pair buffer");
    const int b = blockIdx.y;
    const int64_t npad = n_pad[b], ntrue = n_true[b];
    const int F = (int)(npad / HOP);
    const int f0 = blockIdx.x * FT;
    if (f0 >= F) return;
    const float* x = samples + soff[b];
    for (int i = threadIdx.x; i < NFFT; i += MEL_THREADS) tws[i] = tw[i];
    auto sample = [&](int f, int n) {                          // windowed sample n of frame f (0 beyond the utterance)
        if (f >= F) return 0.f;
        int64_t j = (int64_t)f * HOP + n - NFFT / 2;          // index into the hop-padded waveform
        if (j < 0) j = -j;                                    // reflection_pad1d (mel.rs:63-65)
        if (j >= npad) j = 2 * (npad - 1) - j;
        return ((j < ntrue) ? x[j] : 0.f) * hann[n];          // zero padding of mel.rs:51-53, periodic Hann
    };
    for (int idx = threadIdx.x; idx < FT * (NH + 1); idx += MEL_THREADS) {
        const int fi = idx / (NH + 1), n = idx - fi * (NH + 1);
        const float a = sample(f0 + fi, n);
        if (n == 0 || n == NH) xe[n][fi] = a;
        else {
            const float c = sample(f0 + fi, NFFT - n);
            xe[n][fi] = a + c; xo[n][fi] = a - c;
        }
    }
    __syncthreads();
    const int k = threadIdx.x;
    float re[FT], im[FT];
    if (k < NBIN) {
        const float sgn = (k & 1) ? -1.f : 1.f;
#pragma unroll
        for (int i = 0; i < FT; ++i) { re[i] = xe[0][i] + sgn * xe[NH][i]; im[i] = 0.f; }
        int r = 0;
#pragma unroll 2
        for (int n = 1; n < NH; ++n) {
            r += k; if (r >= NFFT) r -= NFFT;
            const float2 t = tws[r];
#pragma unroll
            for (int i4 = 0; i4 < FT; i4 += 4) {
                const float4 e = *reinterpret_cast<const float4*>(&xe[n][i4]);
                const float4 o = *reinterpret_cast<const float4*>(&xo[n][i4]);
                re[i4] = fmaf(e.x, t.x, re[i4]); re[i4 + 1] = fmaf(e.y, t.x, re[i4 + 1]); re[i4 + 2] = fmaf(e.z, t.x, re[i4 + 2]); re[i4 + 3] = fmaf(e.w, t.x, re[i4 + 3]);
                im[i4] = fmaf(o.x, t.y, im[i4]); im[i4 + 1] = fmaf(o.y, t.y, im[i4 + 1]); im[i4 + 2] = fmaf(o.z, t.y, im[i4 + 2]); im[i4 + 3] = fmaf(o.w, t.y, im[i4 + 3]);
            }
        }
    }
    __syncthreads();                                          // every thread is done reading xe / xo: pw may overwrite them
    if (k < NBIN) {
#pragma unroll
        for (int i = 0; i < FT; ++i) pw[i][k] = re[i] * re[i] + im[i] * im[i];   // abs().square()
    }
    __syncthreads();
    float lmax = -INFINITY;
    float* out = mel_out + (size_t)n_mels * foff[b];
    for (int idx = threadIdx.x; idx < n_mels * FT; idx += MEL_THREADS) {
        int m = idx / FT, fi = idx - m * FT;
        int f = f0 + fi;
        int k0 = krange[2 * m], k1 = krange[2 * m + 1];
        float acc = 0.f;
        for (int kk = k0; kk < k1; ++kk) acc = fmaf(fb[m * NBIN + kk], pw[fi][kk], acc);
        float v = log10f(fmaxf(acc, 1e-10f));                     // clamp_min(1e-10).log10()
        if (f < F) {
            out[(size_t)m * F + f] = v;
            lmax = fmaxf(lmax, v);
        }
    }
    lmax = block_max(lmax, red);
    if (threadIdx.x == 0) atomicMax(&maxkey[b], float_to_ordered(lmax));
}

__global__ void mel_finalize_kernel(float* __restrict__ mel, const int64_t* __restrict__ foff,
                                    const int64_t* __restrict__ n_pad, int n_mels,
                                    const int* __restrict__ maxkey) {
    const int b = blockIdx.y;
    const int64_t total = (int64_t)n_mels * (n_pad[b] / HOP);
    float* p = mel + (size_t)n_mels * foff[b];
    const float floor_v = ordered_to_float(maxkey[b]) - 8.0f;     // maximum(max - 8)  (mel.rs:91-92)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        float v = fmaxf(p[i], floor_v);
        p[i] = (v + 4.0f) / 4.0f;                                 // mel.rs:93
    }
}

__global__ void mel_init_max_kernel(int* maxkey, int batch) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) maxkey[i] = float_to_ordered(-INFINITY);
}

void launch_mel(const Model& m, const float* samples, const int64_t* d_soff, const int64_t* d_n,
                const int64_t* d_npad, const int64_t* d_foff, int batch, int max_frames,
                float* mel_out, int* d_maxkey, cudaStream_t st) {
    mel_init_max_kernel<<<(batch + 127) / 128, 128, 0, st>>>(d_maxkey, batch);
    dim3 grid((max_frames + FT - 1) / FT, batch);
    mel_power_kernel<<<grid, MEL_THREADS, 0, st>>>(samples, d_soff, d_n, d_npad, d_foff, m.hann,
                                                   reinterpret_cast<const float2*>(m.dft_tw), m.mel_fb, m.mel_krange,
                                                   m.d.c.num_mel_bins, mel_out, d_maxkey);
    dim3 g2(148, batch);
    mel_finalize_kernel<<<g2, 256, 0, st>>>(mel_out, d_foff, d_npad, m.d.c.num_mel_bins, d_maxkey);
    ASRB_CUDA_CHECK(cudaGetLastError());
}

}  // namespace asrb
