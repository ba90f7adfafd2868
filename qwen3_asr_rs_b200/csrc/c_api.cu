// c_api.cu -- the extern "C" boundary (include/asr_b200.h).  Exceptions never cross it: every
// entry point returns a status int and records a thread-local message (idiom of the reference's
// own FFI layer, /root/reference/src/backend/mlx/ffi.rs:60-110).
#include <cstring>
#include "internal.h"

namespace asrb {
struct Session;
void model_set_tensor(Model* m, const char* name, int dtype, const int64_t* shape, int ndim, const void* host);
void model_finalize(Model* m);
void model_load_dir(Ctx* ctx, const char* dir, Model** out);
Session* session_create(Model* m, int max_batch, int64_t max_samples, int max_lang, int max_new);
void session_free(Session* s);
void session_mel(Session* s, const float* const* samples, const int64_t* n_samples, int batch, int64_t* n_frames_out);
void session_mel_read(Session* s, int b, float* out);
void session_encode(Session* s, int64_t* n_tokens_out);
void session_encode_read(Session* s, int b, float* out);
void session_prefill(Session* s, const int64_t* const* lang_ids, const int32_t* n_lang_ids, int64_t* seq_lens_out, float* last_logits);
void session_decode_step(Session* s, int64_t* next_ids_out, float* logits);
void session_generate(Session* s, int max_new_tokens, int32_t* ids_out, int32_t* lens_out);
void session_transcribe_ids(Session* s, const float* const* samples, const int64_t* n_samples, int batch,
                            const int64_t* const* lang_ids, const int32_t* n_lang_ids, int max_new_tokens,
                            int32_t* ids_out, int32_t* lens_out);
void session_last_timings(Session* s, float* ms6, int64_t* kernels, int64_t* steps);
void session_set_option(Session* s, const char* key, const char* value);
void session_stats(Session* s, int64_t* out, int n);
void session_ingest_pcm(Session* s, const void* const* pcm, const int64_t* n_frames, const int32_t* channels, const int32_t* rate,
                        const int32_t* format, int batch, int64_t* n_samples_out);
void session_ingested_read(Session* s, int b, float* out);
void session_device_ids(Session* s, const int32_t** ids, const int32_t** lens, int* stride, int* batch);
int decode_mega_debug_timeline(long long* out, int cap);
int decode_batch_debug_timeline(long long* out, int cap);
}  // namespace asrb

using namespace asrb;

namespace asrb {
// shared by asrb_model_create and the config.json loader: anything that would divide by zero or index out of bounds later
void validate_dims(const asrb_dims& d) {
    ASRB_REQUIRE(d.num_mel_bins == 128, ASRB_ERR_INVALID, "num_mel_bins must be 128");
    ASRB_REQUIRE(d.n_window > 0 && d.n_window_infer > 0 && d.d_model > 0 && d.encoder_layers > 0 && d.encoder_attention_heads > 0 &&
                     d.d_model % d.encoder_attention_heads == 0 && d.encoder_ffn_dim > 0 && d.downsample_hidden_size > 0 && d.output_dim > 0,
                 ASRB_ERR_INVALID, "bad audio encoder dims");
    ASRB_REQUIRE(d.hidden_size > 0 && d.intermediate_size > 0 && d.num_hidden_layers > 0 && d.num_attention_heads > 0 &&
                     d.num_key_value_heads > 0 && d.num_attention_heads % d.num_key_value_heads == 0 && d.head_dim > 0 && d.head_dim % 2 == 0,
                 ASRB_ERR_INVALID, "bad text decoder dims");
    ASRB_REQUIRE(d.vocab_size > 151676, ASRB_ERR_INVALID, "bad dims (vocab must contain the prompt special tokens)");
    ASRB_REQUIRE(d.rms_norm_eps > 0 && d.rope_theta > 1.0, ASRB_ERR_INVALID, "bad rms_norm_eps / rope_theta");
}
}  // namespace asrb

static thread_local std::string g_last_error;

template <typename F> static int guarded(F&& f) {
    try { f(); return ASRB_OK; }
    catch (const Error& e) { g_last_error = e.what(); return e.code; }
    catch (const std::bad_alloc&) { g_last_error = "host out of memory"; return ASRB_ERR_INVALID; }
    catch (const std::exception& e) { g_last_error = e.what(); return ASRB_ERR_INVALID; }
    catch (...) { g_last_error = "unknown error"; return ASRB_ERR_INVALID; }
}
#define NONNULL(p) ASRB_REQUIRE((p) != nullptr, ASRB_ERR_INVALID, "null pointer: " #p)

struct asrb_ctx { Ctx c; };
struct asrb_model { Model m; };
struct asrb_session { Session* s; };

extern "C" {

const char* asrb_last_error(void) { return g_last_error.c_str(); }
const char* asrb_version(void) { return "qwen3_asr_rs_b200 0.1 (sm_100a)"; }

int asrb_init(int device, asrb_ctx** out) {
    return guarded([&] {
        NONNULL(out);
        int n = 0;
        ASRB_CUDA_CHECK(cudaGetDeviceCount(&n));
        ASRB_REQUIRE(device >= 0 && device < n, ASRB_ERR_INVALID, "no such CUDA device");
        ASRB_CUDA_CHECK(cudaSetDevice(device));
        cudaDeviceProp p;
        ASRB_CUDA_CHECK(cudaGetDeviceProperties(&p, device));
        ASRB_REQUIRE(p.major == 10, ASRB_ERR_INVALID,
                     "this library is built for sm_100a (B200) only; found compute capability " +
                         std::to_string(p.major) + "." + std::to_string(p.minor));
        asrb_ctx* c = new asrb_ctx();
        c->c.device = device; c->c.sm_count = p.multiProcessorCount; c->c.smem_optin = p.sharedMemPerBlockOptin;
        *out = c;
    });
}
int asrb_ctx_free(asrb_ctx* ctx) { return guarded([&] { delete ctx; }); }

int asrb_dims_default(asrb_dims* d) {
    return guarded([&] {
        NONNULL(d);
        // src/config.rs:52-62, 90-99
        d->d_model = 896; d->encoder_layers = 18; d->encoder_attention_heads = 14; d->encoder_ffn_dim = 3584;
        d->num_mel_bins = 128; d->max_source_positions = 1500; d->n_window = 50; d->n_window_infer = 800;
        d->downsample_hidden_size = 480; d->output_dim = 1024;
        d->vocab_size = 151936; d->hidden_size = 1024; d->intermediate_size = 3072; d->num_hidden_layers = 28;
        d->num_attention_heads = 16; d->num_key_value_heads = 8; d->head_dim = 128; d->tie_word_embeddings = 1;
        d->rms_norm_eps = 1e-6; d->rope_theta = 1000000.0;
    });
}

int asrb_model_create(asrb_ctx* ctx, const asrb_dims* dims, asrb_model** out) {
    return guarded([&] {
        NONNULL(ctx); NONNULL(dims); NONNULL(out);
        validate_dims(*dims);
        ASRB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
        asrb_model* m = new asrb_model();
        m->m.ctx = &ctx->c; m->m.d.c = *dims; m->m.d.derive();
        *out = m;
    });
}
int asrb_model_set_tensor(asrb_model* m, const char* name, int dtype, const int64_t* shape, int ndim, const void* host) {
    return guarded([&] { NONNULL(m); NONNULL(shape); ASRB_CUDA_CHECK(cudaSetDevice(m->m.ctx->device)); model_set_tensor(&m->m, name, dtype, shape, ndim, host); });
}
int asrb_model_finalize(asrb_model* m) {
    return guarded([&] { NONNULL(m); ASRB_CUDA_CHECK(cudaSetDevice(m->m.ctx->device)); model_finalize(&m->m); });
}
int asrb_model_load(asrb_ctx* ctx, const char* model_dir, asrb_model** out) {
    return guarded([&] {
        NONNULL(ctx); NONNULL(model_dir); NONNULL(out);
        ASRB_CUDA_CHECK(cudaSetDevice(ctx->c.device));
        asrb_model* m = new asrb_model();
        m->m.ctx = &ctx->c;
        try {
            Model* mp = &m->m;
            model_load_dir(&ctx->c, model_dir, &mp);
        } catch (...) { delete m; throw; }
        *out = m;
    });
}
int asrb_model_dims(const asrb_model* m, asrb_dims* out) { return guarded([&] { NONNULL(m); NONNULL(out); *out = m->m.d.c; }); }
int asrb_model_lossy_tensors(const asrb_model* m, int* count) { return guarded([&] { NONNULL(m); NONNULL(count); *count = m->m.lossy_count; }); }
int asrb_model_free(asrb_model* m) { return guarded([&] { if (m) { cudaSetDevice(m->m.ctx->device); delete m; } }); }

int asrb_session_create(asrb_model* m, int max_batch, int64_t max_samples, int max_lang_ids, int max_new_tokens, asrb_session** out) {
    return guarded([&] {
        NONNULL(m); NONNULL(out);
        asrb_session* s = new asrb_session();
        try { s->s = session_create(&m->m, max_batch, max_samples, max_lang_ids, max_new_tokens); } catch (...) { delete s; throw; }
        *out = s;
    });
}
int asrb_session_free(asrb_session* s) { return guarded([&] { if (s) { session_free(s->s); delete s; } }); }

int asrb_transcribe_ids(asrb_session* s, const float* const* samples, const int64_t* n_samples, int batch,
                        const int64_t* const* lang_ids, const int32_t* n_lang_ids, int max_new_tokens,
                        int32_t* ids_out, int32_t* lens_out) {
    return guarded([&] { NONNULL(s); NONNULL(samples); NONNULL(n_samples);
                         session_transcribe_ids(s->s, samples, n_samples, batch, lang_ids, n_lang_ids, max_new_tokens, ids_out, lens_out); });
}
int asrb_mel(asrb_session* s, const float* const* samples, const int64_t* n_samples, int batch, int64_t* n_frames_out) {
    return guarded([&] { NONNULL(s); NONNULL(samples); NONNULL(n_samples); session_mel(s->s, samples, n_samples, batch, n_frames_out); });
}
int asrb_mel_read(asrb_session* s, int b, float* out) { return guarded([&] { NONNULL(s); NONNULL(out); session_mel_read(s->s, b, out); }); }
int asrb_encode(asrb_session* s, int64_t* n_tokens_out) { return guarded([&] { NONNULL(s); session_encode(s->s, n_tokens_out); }); }
int asrb_encode_read(asrb_session* s, int b, float* out) { return guarded([&] { NONNULL(s); NONNULL(out); session_encode_read(s->s, b, out); }); }
int asrb_prefill(asrb_session* s, const int64_t* const* lang_ids, const int32_t* n_lang_ids, int64_t* seq_lens_out, float* last_logits) {
    return guarded([&] { NONNULL(s); session_prefill(s->s, lang_ids, n_lang_ids, seq_lens_out, last_logits); });
}
int asrb_decode_step(asrb_session* s, int64_t* next_ids_out, float* logits) {
    return guarded([&] { NONNULL(s); session_decode_step(s->s, next_ids_out, logits); });
}
int asrb_generate(asrb_session* s, int max_new_tokens, int32_t* ids_out, int32_t* lens_out) {
    return guarded([&] { NONNULL(s); NONNULL(ids_out); NONNULL(lens_out); session_generate(s->s, max_new_tokens, ids_out, lens_out); });
}
int asrb_last_timings(asrb_session* s, float* ms_out6, int64_t* kernels_launched, int64_t* decode_steps) {
    return guarded([&] { NONNULL(s); session_last_timings(s->s, ms_out6, kernels_launched, decode_steps); });
}
int asrb_ingest_pcm(asrb_session* s, const void* const* pcm, const int64_t* n_frames, const int32_t* channels,
                    const int32_t* sample_rate, const int32_t* format, int batch, int64_t* n_samples_out) {
    return guarded([&] { NONNULL(s); NONNULL(pcm); NONNULL(n_frames); NONNULL(channels); NONNULL(sample_rate); NONNULL(format);
                         session_ingest_pcm(s->s, pcm, n_frames, channels, sample_rate, format, batch, n_samples_out); });
}
int asrb_ingested_read(asrb_session* s, int b, float* out) { return guarded([&] { NONNULL(s); NONNULL(out); session_ingested_read(s->s, b, out); }); }
int asrb_transcribe_ingested(asrb_session* s, const int64_t* const* lang_ids, const int32_t* n_lang_ids, int max_new_tokens,
                             int32_t* ids_out, int32_t* lens_out) {
    return guarded([&] { NONNULL(s); session_transcribe_ids(s->s, nullptr, nullptr, 0, lang_ids, n_lang_ids, max_new_tokens, ids_out, lens_out); });
}
int asrb_session_device_ids(asrb_session* s, const int32_t** ids_dev, const int32_t** lens_dev, int* row_stride, int* batch) {
    return guarded([&] { NONNULL(s); NONNULL(ids_dev); NONNULL(lens_dev); NONNULL(row_stride); NONNULL(batch);
                         session_device_ids(s->s, ids_dev, lens_dev, row_stride, batch); });
}
int asrb_session_stats(asrb_session* s, int64_t* out, int n) {
    return guarded([&] { NONNULL(s); NONNULL(out); session_stats(s->s, out, n); });
}
int asrb_session_set_option(asrb_session* s, const char* key, const char* value) {
    return guarded([&] { NONNULL(s); session_set_option(s->s, key, value); });
}

int asrb_debug_mega_timeline(long long* out, int cap) {
    int n = 0;
    guarded([&] { n = (getenv("ASRB_MEGA_DEBUG") && std::string(getenv("ASRB_MEGA_DEBUG")) == "batch") ? decode_batch_debug_timeline(out, cap)
                                                                                                            : decode_mega_debug_timeline(out, cap); });
    return n;
}

}  // extern "C"
