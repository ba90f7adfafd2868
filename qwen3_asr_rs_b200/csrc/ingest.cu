// ingest.cu -- GPU-side audio ingest: interleaved PCM (s16 / s32 / f32, any channel count, any rate) -> mono f32 @ 16 kHz,
// written straight into the session's sample buffer (the layout mel.cu reads), so that a WAV file's payload goes
// host -> device ONCE in its native format (2 bytes per sample for s16) and never exists as f32 on the host.
//
// Reference: load_audio_wav + resample, /root/reference/src/audio.rs:162-245 -- hound samples -> f32 (s / 2^(bits-1),
// :181-189), mono mixdown (mean over channels, :193-206), resample to 16 kHz (:209-213).  The reference resamples with
// rubato's windowed-sinc interpolator (crate rubato, not vendored; its FFmpeg path uses swresample): neither is
// reproducible bit for bit here, so -- as for the host loader (audio.py) -- the resampler is the polyphase FIR of
// scipy.signal.resample_poly (Kaiser beta = 5 window, half length 10 * max(up, down)), restated below (filter design in
// double on the host, contraction on the GPU in double) and pinned by its own golden: tests compare against scipy.
// This stage is OUTSIDE the parity point of the hot path (parity starts at the f32 16 kHz vector).
#include <algorithm>
#include <cmath>
#include <map>
#include <numeric>
#include <vector>
#include "internal.h"

namespace asrb {

struct PolyFilter { std::vector<double> h; int up = 1, down = 1, n_pre_remove = 0; };

static double bessel_i0(double x) {           // power series, converges fast for |x| <= 5
    double s = 1.0, t = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 64; ++k) { t *= q / ((double)k * k); s += t; if (t < 1e-18 * s) break; }
    return s;
}

// scipy.signal.resample_poly(x, up, down) filter: firwin(2 * half_len + 1, 1 / max(up, down), window=('kaiser', 5.0)) * up,
// zero-padded in front so that output k lines up with input k * down / up
static PolyFilter design_filter(int up, int down, int64_t n_in) {
    PolyFilter f; f.up = up; f.down = down;
    const int max_rate = std::max(up, down);
    const double cutoff = 1.0 / max_rate, beta = 5.0;
    const int half_len = 10 * max_rate, ntaps = 2 * half_len + 1;
    std::vector<double> h((size_t)ntaps);
    const double alpha = 0.5 * (ntaps - 1), pi = 3.14159265358979323846;
    double sum = 0.0;
    for (int n = 0; n < ntaps; ++n) {
        const double m = n - alpha, xm = cutoff * m;
        const double sinc = (xm == 0.0) ? 1.0 : std::sin(pi * xm) / (pi * xm);
        const double r = m / alpha;
        const double w = bessel_i0(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / bessel_i0(beta);
        h[n] = cutoff * sinc * w;
        sum += h[n];
    }
    for (double& v : h) v = v / sum * up;                       // unit DC gain (firwin scale=True), then * up
    const int n_pre_pad = down - half_len % down;
    f.n_pre_remove = (half_len + n_pre_pad) / down;
    const int64_t n_out = (n_in * up) / down + ((n_in * up) % down ? 1 : 0);
    auto output_len = [&](int64_t len_h) {                    // scipy.signal._upfirdn._output_len
        return ((n_in - 1) * up + len_h - 1) / down + 1;
    };
    int n_post_pad = 0;
    while (output_len((int64_t)ntaps + n_pre_pad + n_post_pad) < n_out + f.n_pre_remove) ++n_post_pad;
    f.h.assign((size_t)n_pre_pad, 0.0);
    f.h.insert(f.h.end(), h.begin(), h.end());
    f.h.insert(f.h.end(), (size_t)n_post_pad, 0.0);
    return f;
}

enum { PCM_S16 = 0, PCM_F32 = 1, PCM_S32 = 2 };

// mono sample n of an interleaved PCM stream (hound: i / 2^(bits-1), audio.rs:181-189; mean over channels, :193-206)
template <int FMT>
__device__ __forceinline__ float pcm_mono(const void* pcm, int64_t n, int C) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
        float v;
        if (FMT == PCM_S16) v = (float)reinterpret_cast<const int16_t*>(pcm)[n * C + c] / 32768.0f;
        else if (FMT == PCM_S32) v = (float)reinterpret_cast<const int32_t*>(pcm)[n * C + c] / 2147483648.0f;
        else v = reinterpret_cast<const float*>(pcm)[n * C + c];
        s += v;
    }
    return C > 1 ? s / (float)C : s;
}

// out[k] = sum_n x[n] * h[(k + pre) * down - n * up]; one thread per output sample; [n_out, n_out_pad) zero filled
template <int FMT>
__global__ void ingest_resample_kernel(const void* __restrict__ pcm, int64_t n_in, int C, const double* __restrict__ h, int len_h,
                                       int up, int down, int pre, float* __restrict__ out, int64_t n_out, int64_t n_out_pad) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_out_pad) return;
    if (k >= n_out) { out[k] = 0.f; return; }
    if (up == 1 && down == 1) { out[k] = pcm_mono<FMT>(pcm, k, C); return; }
    const int64_t t = (k + pre) * (int64_t)down;                 // position on the up-sampled grid
    int64_t n_hi = t / up;                                         // largest n with t - n * up >= 0
    int64_t n_lo = (t - (len_h - 1) + up - 1) / up;                // smallest n with t - n * up <= len_h - 1
    if (t - (len_h - 1) < 0) n_lo = 0;
    n_hi = min(n_hi, n_in - 1);
    double acc = 0.0;
    for (int64_t n = n_lo; n <= n_hi; ++n) acc += (double)pcm_mono<FMT>(pcm, n, C) * h[t - n * up];
    out[k] = (float)acc;
}

struct IngestState {
    std::map<std::pair<int, int>, std::pair<double*, PolyFilter>> filters;   // (up, down) -> device taps
    void* d_raw = nullptr; void* h_raw = nullptr; size_t raw_cap = 0;
    ~IngestState() {
        for (auto& kv : filters) cudaFree(kv.second.first);
        if (d_raw) cudaFree(d_raw);
        if (h_raw) cudaFreeHost(h_raw);
    }
};
IngestState* ingest_state_new() { return new IngestState(); }
void ingest_state_free(IngestState* st) { delete st; }

// Fills `d_samples` (session layout: utterance b at soff[b], zero padded to a multiple of 160) from raw PCM.
// Returns the number of 16 kHz samples per utterance in n_out.
void ingest_pcm(IngestState* st, cudaStream_t stream, const void* const* pcm, const int64_t* n_frames, const int32_t* channels,
                const int32_t* rate, const int32_t* format, int batch, float* d_samples, int64_t max_samples_per_utt,
                int64_t* n_out, int64_t* soff_out) {
    const int TARGET = 16000;
    size_t total_raw = 0;
    std::vector<size_t> roff((size_t)batch);
    for (int b = 0; b < batch; ++b) {
        ASRB_REQUIRE(pcm[b] && n_frames[b] > 0 && channels[b] >= 1 && channels[b] <= 64 && rate[b] >= 1000 && rate[b] <= 768000,
                     ASRB_ERR_INVALID, "ingest: bad PCM description");
        ASRB_REQUIRE(format[b] == PCM_S16 || format[b] == PCM_F32 || format[b] == PCM_S32, ASRB_ERR_INVALID, "ingest: format must be s16 / f32 / s32");
        const size_t bytes = (size_t)n_frames[b] * channels[b] * (format[b] == PCM_S16 ? 2 : 4);
        roff[b] = total_raw; total_raw += (bytes + 255) & ~(size_t)255;
    }
    if (total_raw > st->raw_cap) {
        if (st->d_raw) cudaFree(st->d_raw);
        if (st->h_raw) cudaFreeHost(st->h_raw);
        st->d_raw = nullptr; st->h_raw = nullptr; st->raw_cap = 0;
        ASRB_CUDA_CHECK(cudaMalloc(&st->d_raw, total_raw));
        ASRB_CUDA_CHECK(cudaMallocHost(&st->h_raw, total_raw));
        st->raw_cap = total_raw;
    }
    for (int b = 0; b < batch; ++b)
        memcpy((uint8_t*)st->h_raw + roff[b], pcm[b], (size_t)n_frames[b] * channels[b] * (format[b] == PCM_S16 ? 2 : 4));
    ASRB_CUDA_CHECK(cudaMemcpyAsync(st->d_raw, st->h_raw, total_raw, cudaMemcpyHostToDevice, stream));
    int64_t so = 0;
    for (int b = 0; b < batch; ++b) {
        const int g = std::gcd(TARGET, (int)rate[b]);
        const int up = TARGET / g, down = rate[b] / g;
        const int64_t nout = (n_frames[b] * up) / down + ((n_frames[b] * up) % down ? 1 : 0);
        ASRB_REQUIRE(nout <= max_samples_per_utt, ASRB_ERR_INVALID, "ingest: utterance exceeds session capacity");
        const int64_t npad = ((nout + 159) / 160) * 160;
        const double* d_h = nullptr; int len_h = 0, pre = 0;
        if (up != 1 || down != 1) {
            // the zero padding of the taps depends (only through n_post_pad) on the input length: key on it too when it matters
            auto key = std::make_pair(up, down);
            auto it = st->filters.find(key);
            PolyFilter f = design_filter(up, down, n_frames[b]);
            if (it == st->filters.end() || it->second.second.h.size() < f.h.size()) {
                if (it != st->filters.end()) { cudaFree(it->second.first); st->filters.erase(it); }
                // keep generous trailing zeros so that longer inputs rarely force a re-upload
                f.h.resize(f.h.size() + 4 * (size_t)std::max(up, down), 0.0);
                double* d = nullptr;
                ASRB_CUDA_CHECK(cudaMalloc(&d, f.h.size() * sizeof(double)));
                ASRB_CUDA_CHECK(cudaMemcpyAsync(d, f.h.data(), f.h.size() * sizeof(double), cudaMemcpyHostToDevice, stream));
                ASRB_CUDA_CHECK(cudaStreamSynchronize(stream));        // f.h is a temporary
                it = st->filters.emplace(key, std::make_pair(d, f)).first;
            }
            d_h = it->second.first; len_h = (int)it->second.second.h.size(); pre = it->second.second.n_pre_remove;
        }
        const void* src = (const uint8_t*)st->d_raw + roff[b];
        const int threads = 256; const unsigned blocks = (unsigned)((npad + threads - 1) / threads);
        float* dst = d_samples + so;
        if (format[b] == PCM_S16) ingest_resample_kernel<PCM_S16><<<blocks, threads, 0, stream>>>(src, n_frames[b], channels[b], d_h, len_h, up, down, pre, dst, nout, npad);
        else if (format[b] == PCM_S32) ingest_resample_kernel<PCM_S32><<<blocks, threads, 0, stream>>>(src, n_frames[b], channels[b], d_h, len_h, up, down, pre, dst, nout, npad);
        else ingest_resample_kernel<PCM_F32><<<blocks, threads, 0, stream>>>(src, n_frames[b], channels[b], d_h, len_h, up, down, pre, dst, nout, npad);
        ASRB_CUDA_CHECK(cudaGetLastError());
        n_out[b] = nout; soff_out[b] = so;
        so += npad;
    }
}

}  // namespace asrb
