/*
 * asr_b200.h -- C ABI of the B200-native Qwen3-ASR hot path.
 *
 * Drop-in boundary for second-state/qwen3_asr_rs (reference paths relative to
 * /root/reference).  The reference has no backend trait: its seam is the cfg-switched
 * `struct Tensor` (src/tensor.rs:120-126) that the three hot modules call
 * (src/mel.rs, src/audio_encoder.rs, src/text_decoder.rs) from
 * `AsrInference::transcribe` (src/inference.rs:89-213).  This library replaces the
 * span steps 2-8 of that function (src/inference.rs:94-200) -- f32 samples in host
 * memory -> generated token ids in host memory -- with hand-written sm_100a kernels.
 * A third `#[cfg(feature = "b200")]` arm binds these symbols.
 *
 * Conventions follow the reference's own FFI idiom (src/backend/mlx/ffi.rs:60-110):
 * every function returns an int status (0 = ok), results come back through
 * out-parameters, handles are opaque, every handle has an explicit _free.  Nothing
 * throws or aborts across this boundary; asrb_last_error() returns a thread-local
 * message for the last non-zero status.  No torch / C++ types appear in signatures.
 *
 * Threading: one CUDA stream per session; distinct sessions may be driven from
 * distinct threads; a single session is not re-entrant (the reference is
 * single-threaded and synchronous, src/inference.rs:89).
 */
#ifndef ASR_B200_H
#define ASR_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define ASRB_API __attribute__((visibility("default")))
#else
#define ASRB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ASRB_OK 0
#define ASRB_ERR_INVALID 1   /* bad argument / shape / missing tensor            */
#define ASRB_ERR_CUDA 2      /* CUDA runtime or driver error                      */
#define ASRB_ERR_IO 3        /* model directory / safetensors / config.json       */
#define ASRB_ERR_STATE 4     /* call order (e.g. encode before mel)               */

typedef struct asrb_ctx asrb_ctx;         /* one per device                              */
typedef struct asrb_model asrb_model;     /* immutable weights, shareable by sessions   */
typedef struct asrb_session asrb_session; /* KV cache, scratch, streams for ONE batch   */

/* dtype codes for asrb_model_set_tensor (safetensors dtypes the reference accepts,
 * src/weights.rs:74-117) */
#define ASRB_DT_F32 0
#define ASRB_DT_BF16 1
#define ASRB_DT_F16 2

/* Model hyper-parameters: the fields of src/config.rs:27-113, defaults = 0.6B. */
typedef struct asrb_dims {
    /* audio encoder (AudioEncoderConfig, src/config.rs:27-62) */
    int32_t d_model, encoder_layers, encoder_attention_heads, encoder_ffn_dim;
    int32_t num_mel_bins, max_source_positions, n_window, n_window_infer;
    int32_t downsample_hidden_size, output_dim;
    /* text decoder (TextDecoderConfig, src/config.rs:66-113) */
    int32_t vocab_size, hidden_size, intermediate_size, num_hidden_layers;
    int32_t num_attention_heads, num_key_value_heads, head_dim;
    int32_t tie_word_embeddings;
    double rms_norm_eps, rope_theta;
} asrb_dims;

/* ---- context --------------------------------------------------------------------- */
/* replaces the device pick of src/main.rs:51-58 */
ASRB_API int asrb_init(int device, asrb_ctx** out);
ASRB_API int asrb_ctx_free(asrb_ctx* ctx);
ASRB_API const char* asrb_last_error(void);
ASRB_API const char* asrb_version(void);

/* ---- model ----------------------------------------------------------------------- */
/* fills *d with src/config.rs defaults (Qwen3-ASR-0.6B) */
ASRB_API int asrb_dims_default(asrb_dims* d);
/* AsrInference::load (src/inference.rs:30-86): config.json + model.safetensors or
 * model.safetensors.index.json + shards (src/weights.rs:10-58).  bf16 stays bf16. */
ASRB_API int asrb_model_load(asrb_ctx* ctx, const char* model_dir, asrb_model** out);
/* Incremental construction (what load() does internally; also used by tests to
 * build models in memory): create -> set_tensor for every HF name -> finalize. */
ASRB_API int asrb_model_create(asrb_ctx* ctx, const asrb_dims* dims, asrb_model** out);
ASRB_API int asrb_model_set_tensor(asrb_model* m, const char* name, int dtype,
                          const int64_t* shape, int ndim, const void* host_data);
ASRB_API int asrb_model_finalize(asrb_model* m);
ASRB_API int asrb_model_dims(const asrb_model* m, asrb_dims* out);
/* Matrices are kept in bf16 (lossless for the released bf16 checkpoints; the reference widens them to f32,
 * src/weights.rs:74-89).  An F32/F16 matrix that is not bf16-representable is REJECTED by set_tensor / load unless
 * ASRB_ALLOW_LOSSY_WEIGHTS=1; *count = number of matrices that were rounded under that override (0 = exact). */
ASRB_API int asrb_model_lossy_tensors(const asrb_model* m, int* count);
ASRB_API int asrb_model_free(asrb_model* m);

/* ---- session --------------------------------------------------------------------- */
/* Capacity: up to max_batch utterances of up to max_samples samples each, prompt
 * suffix of up to max_lang_ids forced-language ids, up to max_new_tokens generated
 * ids (the reference caps at 4096, src/inference.rs:153). */
ASRB_API int asrb_session_create(asrb_model* m, int max_batch, int64_t max_samples,
                        int max_lang_ids, int max_new_tokens, asrb_session** out);
ASRB_API int asrb_session_free(asrb_session* s);

/* Whole hot path, the call `transcribe()` makes once per file (src/inference.rs:94-200)
 * generalised to a batch of independent utterances:
 *   samples[b]        f32 mono 16 kHz (src/mel.rs:49), n_samples[b] of them
 *   lang_ids[b]       NULL, or the ids of tokenizer.encode("language Xxx")
 *                     (src/inference.rs:246-250) appended to the prompt
 *   ids_out           [batch][max_new_tokens] generated ids (EOS excluded)
 *   lens_out          [batch] number of ids generated
 * Greedy argmax; stops a sequence at EOS {151643,151645} (src/inference.rs:154,163)
 * or at max_new_tokens. */
ASRB_API int asrb_transcribe_ids(asrb_session* s, const float* const* samples, const int64_t* n_samples,
                        int batch, const int64_t* const* lang_ids, const int32_t* n_lang_ids,
                        int max_new_tokens, int32_t* ids_out, int32_t* lens_out);

/* ---- GPU-side audio ingest (step 1 of transcribe(), load_audio_wav + resample, src/audio.rs:162-245) ----------- */
/* Raw interleaved PCM of `batch` utterances (hound samples: s16 / s32 scaled by 2^-(bits-1), or f32; src/audio.rs:181-189)
 * -> mono mixdown (mean over channels, :193-206) -> 16 kHz (:209-213) ON THE GPU, written straight into the session's
 * sample buffer: the payload crosses PCIe once in its native format.  The resampler is the polyphase FIR of
 * scipy.signal.resample_poly (the reference's rubato / swresample interpolators are not reproducible here); this stage
 * is outside the parity point of the hot path and is pinned by its own golden (tests/test_gpu_parity.py).
 * n_samples_out[b] = number of 16 kHz samples produced.  Follow with asrb_transcribe_ingested (or asrb_mel with
 * samples == NULL). */
#define ASRB_PCM_S16 0
#define ASRB_PCM_F32 1
#define ASRB_PCM_S32 2
ASRB_API int asrb_ingest_pcm(asrb_session* s, const void* const* pcm, const int64_t* n_frames, const int32_t* channels,
                    const int32_t* sample_rate, const int32_t* format, int batch, int64_t* n_samples_out);
ASRB_API int asrb_ingested_read(asrb_session* s, int b, float* out /* [n_samples[b]] */);
/* asrb_transcribe_ids on the utterances ingested by the last asrb_ingest_pcm */
ASRB_API int asrb_transcribe_ingested(asrb_session* s, const int64_t* const* lang_ids, const int32_t* n_lang_ids,
                             int max_new_tokens, int32_t* ids_out, int32_t* lens_out);

/* Stage entry points = the calls transcribe() makes (each runs on the session stream;
 * *_read functions synchronise and copy to host, for parity tests). */
/* WhisperFeatureExtractor::extract, src/mel.rs:49-96 (called at src/inference.rs:95) */
ASRB_API int asrb_mel(asrb_session* s, const float* const* samples, const int64_t* n_samples, int batch,
             int64_t* n_frames_out);
ASRB_API int asrb_mel_read(asrb_session* s, int b, float* out /* [num_mel_bins * n_frames[b]] */);
/* AudioEncoder::forward, src/audio_encoder.rs:79-169 (src/inference.rs:100) */
ASRB_API int asrb_encode(asrb_session* s, int64_t* n_tokens_out);
ASRB_API int asrb_encode_read(asrb_session* s, int b, float* out /* [n_tokens[b] * output_dim] */);
/* build_prompt + embed + inject + MRoPE + prefill, src/inference.rs:105-149; writes the
 * last-row logits [batch][vocab] if last_logits != NULL (parity mode) */
ASRB_API int asrb_prefill(asrb_session* s, const int64_t* const* lang_ids, const int32_t* n_lang_ids,
                 int64_t* seq_lens_out, float* last_logits);
/* one greedy iteration, src/inference.rs:160-200: argmax of the pending logits ->
 * next_ids_out[b] (-1 when b already hit EOS) -> embed -> decoder forward with S=1;
 * logits [batch][vocab] if non-NULL */
ASRB_API int asrb_decode_step(asrb_session* s, int64_t* next_ids_out, float* logits);
/* remaining iterations with no per-token host sync */
ASRB_API int asrb_generate(asrb_session* s, int max_new_tokens, int32_t* ids_out, int32_t* lens_out);

/* ---- introspection for bench / tests ---------------------------------------------- */
/* per-stage device milliseconds of the last asrb_transcribe_ids (CUDA events on the session
 * stream): [0]=h2d of samples [1]=mel [2]=encoder [3]=prefill [4]=decode loop [5]=total;
 * plus counters (kernels launched, decoder forward steps) */
ASRB_API int asrb_last_timings(asrb_session* s, float* ms_out6, int64_t* kernels_launched,
                      int64_t* decode_steps);
/* Device-resident results of the last asrb_generate / asrb_transcribe_ids (valid until the next call on this
 * session; the session stream has been synchronised): ids [max_batch][max_new_tokens] int32 and lens [max_batch]
 * int32 in HBM.  This is what the multi-GPU gather (the path's only collective: one all-gather of ids over
 * NCCL / NVLink) reads directly, with no host staging. */
ASRB_API int asrb_session_device_ids(asrb_session* s, const int32_t** ids_dev, const int32_t** lens_dev,
                            int* row_stride, int* batch);
/* Path counters since session creation -- silent fallbacks made visible (bench.py asserts the fallback ones are 0):
 *   [0] decoder forwards on the batch-aware fused step   [1] on the single-sequence fused step
 *   [2] on the per-phase kernels (fallback: logits requested, unsupported dims, context beyond the fused limit)
 *   [3] GEMMs that fell back from tcgen05 to the SIMT kernel (process-wide)   [4] tcgen05 GEMM launches (process-wide)
 * writes min(n, 5) values */
ASRB_API int asrb_session_stats(asrb_session* s, int64_t* out, int n);
/* knobs: "gemm" = "tc"|"simt", "decode" = "mega"|"phases", "batch_step" = "1"|"0", "planes" = "1"|"2"|"3",
 * "resident" = "1"|"0" (1: the samples uploaded by the previous call are reused, no H2D) */
ASRB_API int asrb_session_set_option(asrb_session* s, const char* key, const char* value);

/* debug (ASRB_MEGA_DEBUG=1): clock64 timeline of the last fused decode step, CTA 0 then CTA G-1;
 * returns the number of slots per CTA (0 if disabled) */
ASRB_API int asrb_debug_mega_timeline(long long* out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* ASR_B200_H */
