// session.cu -- host driver: restates AsrInference::transcribe steps 2-8
// (/root/reference/src/inference.rs:94-200) for a BATCH of independent utterances on one GPU.
// Control flow (prompt layout, position ids, greedy loop, EOS) is the reference's; all arithmetic
// is in the kernels.  Everything is varlen: rows of all utterances are concatenated and described
// by small index arrays uploaded once per call.
#include <algorithm>
#include <cstring>
#include "internal.h"

namespace asrb {

void model_set_tensor(Model* m, const char* name, int dtype, const int64_t* shape, int ndim, const void* host);
void model_finalize(Model* m);
struct IngestState;
IngestState* ingest_state_new();
void ingest_state_free(IngestState* st);
void ingest_pcm(IngestState* st, cudaStream_t stream, const void* const* pcm, const int64_t* n_frames, const int32_t* channels,
                const int32_t* rate, const int32_t* format, int batch, float* d_samples, int64_t max_samples_per_utt,
                int64_t* n_out, int64_t* soff_out);

// token ids of the fixed prompt (inference.rs:215-257) and special tokens (tokenizer.rs:53-59)
static const int kPromptHead[9] = {151644, 8948, 198, 151645, 198, 151644, 872, 198, 151669};
static const int kPromptTail[6] = {151670, 151645, 198, 151644, 77091, 198};
static const int kAudioPad = 151676;

struct Session {
    Model* m = nullptr;
    cudaStream_t st = nullptr;
    int max_batch = 0, max_lang = 0, max_new = 0;
    int64_t max_samples = 0, max_npad = 0;
    int maxF = 0, maxC = 0, maxT = 0, maxS = 0, max_ctx = 0;
    int gemm_impl = GEMM_TC;
    int decode_mode = 1;   // 1 = fused step when available, 0 = per-phase kernels
    bool batch_step = true;   // batch >= 2: the batch-aware fused step (decode_batch.cu); false = one fused launch per sequence
    int64_t n_batch_steps = 0, n_mega_steps = 0, n_phase_steps = 0;   // decoder forwards by path since session creation
    int nplanes = 3;
    // ---- current batch plan (host) ----
    int stage = 0;         // 0 idle, 1 mel, 2 encoded, 3 prefilled
    int B = 0;
    std::vector<int64_t> n, npad, F, foff, soff;
    std::vector<int> C, T, toff, S, srow0;
    int totF = 0, totC = 0, totT = 0, totS = 0, maxlenS = 0, maxwin = 0, nwin = 0;
    // ---- device buffers ----
    std::vector<void*> owned;
    float *h_samples = nullptr, *d_samples = nullptr, *d_mel = nullptr;
    int64_t* d_i64 = nullptr;      // soff | n | npad | foff | frames  (5 * max_batch)
    int64_t* h_i64 = nullptr;
    int *d_int = nullptr, *h_int = nullptr; size_t int_cap = 0, enc_int_cap = 0;   // [encoder plan | prefill plan]
    int* d_maxkey = nullptr;
    bf16 *act1 = nullptr, *act2 = nullptr, *feat = nullptr; size_t act1_ps = 0, act2_ps = 0, feat_ps = 0;
    float *x_enc = nullptr, *enc_qkv = nullptr, *audio = nullptr;
    bf16 *enc_h = nullptr, *enc_attn = nullptr, *enc_ff = nullptr; size_t ench_ps = 0, encff_ps = 0;
    float *hid = nullptr, *dqkv = nullptr, *qrot = nullptr;
    bf16 *dh = nullptr, *dattn = nullptr, *dact = nullptr; size_t dh_ps = 0, dattn_ps = 0, dact_ps = 0;
    float *kcache = nullptr, *vcache = nullptr; size_t cache_layer_stride = 0, cache_seq_stride = 0;
    DecodeBufs db{};
    float* splitk_ws = nullptr;   // fp32 partial tiles of split-K GEMMs (gemm_tc.cu)
    MegaBufs mega{};
    int sx_nb = 0;             // NB instantiation the exchange words were last armed for (decode_batch.cu)
    unsigned mega_steps = 1;   // host mirror of the device epoch (upper bound): see launch_decode_step_mega
    int* d_lastrow = nullptr;
    int *h_done = nullptr, *h_ids = nullptr, *h_nout = nullptr, *h_next = nullptr;
    // int-plan offsets (into d_int)
    int *d_chunk_clip = nullptr, *d_chunk_f0 = nullptr, *d_rowmap = nullptr, *d_win_q0 = nullptr, *d_win_len = nullptr;
    int *d_ids = nullptr, *d_audio_row = nullptr, *d_row_seq = nullptr, *d_row_pos = nullptr, *d_seq_q0 = nullptr, *d_seq_len = nullptr;
    // decode graph
    cudaGraphExec_t step_graph = nullptr; int graph_B = 0; int graph_mode = -1;
    // stats
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float last_ms[6] = {0, 0, 0, 0, 0, 0};
    bool resident = false, timing = false;
    IngestState* ingest = nullptr;      // GPU-side audio ingest (ingest.cu)
    std::vector<int64_t> ingested_n;    // 16 kHz samples per utterance produced by the last asrb_ingest_pcm (empty: none pending)
    int64_t launches = 0, decode_steps = 0;
    int greedy_done = 0;   // greedy applications since prefill (tokens appended or EOS), bounds max_new
    ~Session();
};

Session::~Session() {
    if (step_graph) cudaGraphExecDestroy(step_graph);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
    for (void* p : owned) cudaFree(p);
    if (h_samples) cudaFreeHost(h_samples);
    if (h_i64) cudaFreeHost(h_i64);
    if (h_int) cudaFreeHost(h_int);
    if (h_done) cudaFreeHost(h_done);
    if (h_ids) cudaFreeHost(h_ids);
    if (h_nout) cudaFreeHost(h_nout);
    if (h_next) cudaFreeHost(h_next);
    if (ingest) ingest_state_free(ingest);
    if (st) cudaStreamDestroy(st);
}

template <typename T> static T* salloc(Session* s, size_t n, bool zero = false) {
    T* p = nullptr;
    ASRB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    s->owned.push_back(p);
    if (zero) ASRB_CUDA_CHECK(cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return p;
}

Session* session_create(Model* m, int max_batch, int64_t max_samples, int max_lang, int max_new) {
    ASRB_REQUIRE(m && m->finalized, ASRB_ERR_STATE, "model not finalized");
    ASRB_REQUIRE(max_batch >= 1 && max_samples > 200 && max_new >= 1 && max_lang >= 0, ASRB_ERR_INVALID, "bad session capacity");
    ASRB_CUDA_CHECK(cudaSetDevice(m->ctx->device));
    Session* s = new Session();
    try {
        const Dims& d = m->d; const asrb_dims& c = d.c;
        if (const char* e = getenv("ASRB_GEMM")) s->gemm_impl = (std::string(e) == "simt") ? GEMM_SIMT : GEMM_TC;     // debug overrides
        if (const char* e = getenv("ASRB_DECODE")) s->decode_mode = (std::string(e) == "phases") ? 0 : 1;
        if (const char* e = getenv("ASRB_PLANES")) s->nplanes = std::min(3, std::max(1, atoi(e)));
        s->m = m; s->max_batch = max_batch; s->max_samples = max_samples; s->max_lang = max_lang; s->max_new = max_new;
        s->max_npad = ((max_samples + 159) / 160) * 160;
        s->maxF = (int)(s->max_npad / 160);
        s->maxC = (s->maxF + d.chunk_frames - 1) / d.chunk_frames;
        s->maxT = s->maxC * d.tok_per_chunk;
        s->maxS = s->maxT + 15 + max_lang;
        s->max_ctx = s->maxS + max_new;
        ASRB_REQUIRE(s->max_ctx <= m->rope_max_pos, ASRB_ERR_INVALID, "context exceeds RoPE table");
        ASRB_CUDA_CHECK(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
        for (auto& e : s->ev) ASRB_CUDA_CHECK(cudaEventCreate(&e));
        const size_t Bm = max_batch;
        ASRB_CUDA_CHECK(cudaMallocHost(&s->h_samples, Bm * s->max_npad * sizeof(float)));
        s->d_samples = salloc<float>(s, Bm * s->max_npad);
        s->d_mel = salloc<float>(s, Bm * c.num_mel_bins * s->maxF);
        ASRB_CUDA_CHECK(cudaMallocHost(&s->h_i64, 5 * Bm * sizeof(int64_t)));
        s->d_i64 = salloc<int64_t>(s, 5 * Bm);
        s->d_maxkey = salloc<int>(s, Bm);
        const size_t totC = Bm * s->maxC, totT = Bm * s->maxT, totS = Bm * s->maxS;
        s->enc_int_cap = 2 * totC + totC * d.tok_per_chunk + 2 * (totC + Bm) + 16;
        s->int_cap = s->enc_int_cap + 4 * totS + 8 * Bm + 16;
        ASRB_CUDA_CHECK(cudaMallocHost(&s->h_int, s->int_cap * sizeof(int)));
        s->d_int = salloc<int>(s, s->int_cap);
        // encoder activations
        s->act1_ps = totC * 4 * d.conv_h[2] * d.conv_w[2] * d.cpad;
        s->act2_ps = totC * 4 * d.conv_h[3] * d.conv_w[3] * d.cpad;
        s->feat_ps = totC * d.tok_per_chunk * (size_t)d.feat;
        s->act1 = salloc<bf16>(s, 3 * s->act1_ps, true);     // zero: channel padding / odd-parity pad column
        s->act2 = salloc<bf16>(s, 3 * s->act2_ps, true);
        s->feat = salloc<bf16>(s, 3 * s->feat_ps);
        s->x_enc = salloc<float>(s, totT * c.d_model);
        s->enc_qkv = salloc<float>(s, totT * 3 * c.d_model);
        s->audio = salloc<float>(s, totT * c.output_dim);
        s->ench_ps = totT * c.d_model; s->encff_ps = totT * c.encoder_ffn_dim;
        s->enc_h = salloc<bf16>(s, 3 * s->ench_ps);
        s->enc_attn = salloc<bf16>(s, 3 * s->ench_ps);
        s->enc_ff = salloc<bf16>(s, 3 * s->encff_ps);
        s->splitk_ws = salloc<float>(s, SPLITK_WS_FLOATS);
        // decoder activations
        s->hid = salloc<float>(s, totS * c.hidden_size);
        s->dqkv = salloc<float>(s, totS * d.qkv_dim);
        s->qrot = salloc<float>(s, totS * d.q_dim);
        s->dh_ps = totS * c.hidden_size; s->dattn_ps = totS * d.q_dim; s->dact_ps = totS * c.intermediate_size;
        s->dh = salloc<bf16>(s, 3 * s->dh_ps);
        s->dattn = salloc<bf16>(s, 3 * s->dattn_ps);
        s->dact = salloc<bf16>(s, 3 * s->dact_ps);
        s->cache_seq_stride = (size_t)c.num_key_value_heads * s->max_ctx * c.head_dim;
        s->cache_layer_stride = Bm * s->cache_seq_stride;
        s->kcache = salloc<float>(s, c.num_hidden_layers * s->cache_layer_stride);
        s->vcache = salloc<float>(s, c.num_hidden_layers * s->cache_layer_stride);
        // decode state
        DecodeBufs& b = s->db;
        b.x = salloc<float>(s, Bm * c.hidden_size);
        b.qkv = salloc<float>(s, Bm * d.qkv_dim);
        b.attn = salloc<float>(s, Bm * d.q_dim);
        b.act = salloc<float>(s, Bm * c.intermediate_size);
        b.logits = salloc<float>(s, Bm * c.vocab_size);
        b.n_part = m->ctx->sm_count * 2;
        b.part_val = salloc<float>(s, Bm * b.n_part);
        b.part_idx = salloc<int>(s, Bm * b.n_part);
        b.pos = salloc<int>(s, Bm); b.done = salloc<int>(s, Bm); b.next_id = salloc<int>(s, Bm);
        b.ids_out = salloc<int>(s, Bm * max_new); b.n_out = salloc<int>(s, Bm);
        b.max_new = max_new;
        s->d_lastrow = salloc<int>(s, Bm);
        s->mega.bar = salloc<unsigned>(s, 4, true);
        { const unsigned one = 1; ASRB_CUDA_CHECK(cudaMemcpy(s->mega.bar + 1, &one, sizeof(one), cudaMemcpyHostToDevice)); }   // epoch 1
        const size_t part_floats = std::max(decode_mega_part_floats(*m), decode_batch_part_floats(*m));
        s->mega.part = salloc<float>(s, part_floats, true);   // zero = tag 0 = never written
        s->mega.part_bytes = part_floats * sizeof(float);
        s->mega.steps_issued = &s->mega_steps;
        if (max_batch >= 2) {   // batch-aware fused step: self-validating exchange words, all "not written yet" (0xFFFFFFFF)
            s->mega.sx_bytes = decode_batch_sx_bytes(*m);
            s->mega.sx = (uint32_t*)salloc<uint8_t>(s, s->mega.sx_bytes);
            ASRB_CUDA_CHECK(cudaMemset(s->mega.sx, 0xFF, s->mega.sx_bytes));
            s->mega.sx_nb = &s->sx_nb;
            // keep them L2-resident between steps (2-3 GB of weights and KV stream through L2 every step): persisting
            // access window on the session stream; best effort (ignored if the device refuses the set-aside)
            int maxwin = 0, maxpersist = 0;
            cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, m->ctx->device);
            cudaDeviceGetAttribute(&maxpersist, cudaDevAttrMaxPersistingL2CacheSize, m->ctx->device);
            if (maxwin > 0 && maxpersist > 0) {
                const size_t want = std::min<size_t>({s->mega.sx_bytes, (size_t)maxwin, (size_t)maxpersist});
                if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
                    cudaStreamAttrValue av{};
                    av.accessPolicyWindow.base_ptr = s->mega.sx; av.accessPolicyWindow.num_bytes = want;
                    av.accessPolicyWindow.hitRatio = 1.0f; av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                    if (cudaStreamSetAttribute(s->st, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) cudaGetLastError();
                } else cudaGetLastError();
            }
        }
        if (getenv("ASRB_MEGA_DEBUG")) s->mega.dbg = salloc<long long>(s, decode_mega_dbg_slots(), true);
        ASRB_CUDA_CHECK(cudaMallocHost(&s->h_done, Bm * sizeof(int)));
        ASRB_CUDA_CHECK(cudaMallocHost(&s->h_nout, Bm * sizeof(int)));
        ASRB_CUDA_CHECK(cudaMallocHost(&s->h_next, Bm * sizeof(int)));
        ASRB_CUDA_CHECK(cudaMallocHost(&s->h_ids, Bm * max_new * sizeof(int)));
        ASRB_CUDA_CHECK(cudaDeviceSynchronize());
    } catch (...) { delete s; throw; }
    return s;
}

void session_free(Session* s) { delete s; }

// -------------------------------------------------------------------------------------------------
// step 2: mel  (inference.rs:95)
// -------------------------------------------------------------------------------------------------
void session_mel(Session* s, const float* const* samples, const int64_t* n_samples, int batch, int64_t* n_frames_out) {
    ASRB_REQUIRE(batch >= 1 && batch <= s->max_batch, ASRB_ERR_INVALID, "batch exceeds session capacity");
    ASRB_CUDA_CHECK(cudaSetDevice(s->m->ctx->device));
    const bool ingested = (samples == nullptr);          // samples already in HBM, written by asrb_ingest_pcm
    if (ingested) ASRB_REQUIRE((int)s->ingested_n.size() == batch, ASRB_ERR_STATE, "no ingested audio for this batch");
    s->B = batch; s->stage = 0;
    s->n.assign(batch, 0); s->npad.assign(batch, 0); s->F.assign(batch, 0); s->foff.assign(batch, 0); s->soff.assign(batch, 0);
    int64_t so = 0, fo = 0; int maxF = 0;
    for (int b = 0; b < batch; ++b) {
        int64_t n = ingested ? s->ingested_n[b] : n_samples[b];
        ASRB_REQUIRE((ingested || samples[b]) && n > 0 && n <= s->max_samples, ASRB_ERR_INVALID, "n_samples out of session capacity");
        int64_t np = ((n + 159) / 160) * 160;                                       // mel.rs:51
        ASRB_REQUIRE(np > 200, ASRB_ERR_INVALID, "utterance too short for reflect padding (needs > 200 samples)");
        s->n[b] = n; s->npad[b] = np; s->F[b] = np / 160; s->soff[b] = so; s->foff[b] = fo;
        if (!s->resident && !ingested) {
            memcpy(s->h_samples + so, samples[b], n * sizeof(float));
            if (np > n) memset(s->h_samples + so + n, 0, (np - n) * sizeof(float));
        }
        so += np; fo += np / 160; maxF = std::max<int>(maxF, (int)(np / 160));
    }
    s->totF = (int)fo;
    int64_t* h = s->h_i64; const int Bm = s->max_batch;
    for (int b = 0; b < batch; ++b) { h[b] = s->soff[b]; h[Bm + b] = s->n[b]; h[2 * Bm + b] = s->npad[b]; h[3 * Bm + b] = s->foff[b]; h[4 * Bm + b] = s->F[b]; }
    ASRB_CUDA_CHECK(cudaMemcpyAsync(s->d_i64, h, 5 * Bm * sizeof(int64_t), cudaMemcpyHostToDevice, s->st));
    if (!s->resident && !ingested)
        ASRB_CUDA_CHECK(cudaMemcpyAsync(s->d_samples, s->h_samples, so * sizeof(float), cudaMemcpyHostToDevice, s->st));
    if (ingested) s->ingested_n.clear();                 // consumed
    if (s->timing) ASRB_CUDA_CHECK(cudaEventRecord(s->ev[1], s->st));
    launch_mel(*s->m, s->d_samples, s->d_i64, s->d_i64 + Bm, s->d_i64 + 2 * Bm, s->d_i64 + 3 * Bm, batch, maxF, s->d_mel,
               s->d_maxkey, s->st);
    s->launches += 3;
    if (n_frames_out) for (int b = 0; b < batch; ++b) n_frames_out[b] = s->F[b];
    s->stage = 1;
}

// step 1 on the GPU (src/audio.rs:162-245): raw interleaved PCM -> mono 16 kHz f32 in the session's sample buffer
void session_ingest_pcm(Session* s, const void* const* pcm, const int64_t* n_frames, const int32_t* channels, const int32_t* rate,
                        const int32_t* format, int batch, int64_t* n_samples_out) {
    ASRB_REQUIRE(batch >= 1 && batch <= s->max_batch, ASRB_ERR_INVALID, "batch exceeds session capacity");
    ASRB_CUDA_CHECK(cudaSetDevice(s->m->ctx->device));
    if (!s->ingest) s->ingest = ingest_state_new();
    std::vector<int64_t> n((size_t)batch), so((size_t)batch);
    ingest_pcm(s->ingest, s->st, pcm, n_frames, channels, rate, format, batch, s->d_samples, s->max_samples, n.data(), so.data());
    s->ingested_n = n;
    s->stage = 0;
    if (n_samples_out) for (int b = 0; b < batch; ++b) n_samples_out[b] = n[b];
}
void session_ingested_read(Session* s, int b, float* out) {
    ASRB_REQUIRE(b >= 0 && b < (int)s->ingested_n.size(), ASRB_ERR_STATE, "ingested_read: nothing ingested for this index");
    ASRB_CUDA_CHECK(cudaStreamSynchronize(s->st));
    int64_t so = 0;
    for (int i = 0; i < b; ++i) so += ((s->ingested_n[i] + 159) / 160) * 160;
    ASRB_CUDA_CHECK(cudaMemcpy(out, s->d_samples + so, (size_t)s->ingested_n[b] * sizeof(float), cudaMemcpyDeviceToHost));
}

void session_mel_read(Session* s, int b, float* out) {
    ASRB_REQUIRE(s->stage >= 1 && b >= 0 && b < s->B, ASRB_ERR_STATE, "mel_read: no mel for this index");
    ASRB_CUDA_CHECK(cudaStreamSynchronize(s->st));
    const int nm = s->m->d.c.num_mel_bins;
    ASRB_CUDA_CHECK(cudaMemcpy(out, s->d_mel + (size_t)nm * s->foff[b], (size_t)nm * s->F[b] * sizeof(float), cudaMemcpyDeviceToHost));
}

// -------------------------------------------------------------------------------------------------
// step 3: audio encoder  (inference.rs:100 -> audio_encoder.rs:79-169)
// -------------------------------------------------------------------------------------------------
static GemmA plainA(const bf16* a, size_t ps, int M, int K, int nplanes) {
    GemmA A; A.mode = A_PLAIN; A.a = a; A.plane_stride = ps; A.M = M; A.K = K; A.lda = K; A.nplanes = nplanes; return A;
}

void session_encode(Session* s, int64_t* n_tokens_out) {
    ASRB_REQUIRE(s->stage >= 1, ASRB_ERR_STATE, "encode called before mel");
    Model& m = *s->m; const Dims& d = m.d; const asrb_dims& c = d.c;
    const int B = s->B, tpc = d.tok_per_chunk, cf = d.chunk_frames;
    ASRB_CUDA_CHECK(cudaSetDevice(m.ctx->device));
    // ---- plan: chunks, valid tokens, windows (audio_encoder.rs:83-121,141-152,172-209) ----
    s->C.assign(B, 0); s->T.assign(B, 0); s->toff.assign(B, 0);
    int totC = 0, totT = 0;
    for (int b = 0; b < B; ++b) { s->C[b] = (int)((s->F[b] + cf - 1) / cf); totC += s->C[b]; }
    int* hi = s->h_int;
    int* chunk_clip = hi; int* chunk_f0 = chunk_clip + totC; int* rowmap = chunk_f0 + totC;
    int* win_q0 = rowmap + (size_t)totC * tpc; int* win_len = win_q0 + (totC + B);
    int ci = 0, nwin = 0, maxwin = 0;
    for (int b = 0; b < B; ++b) {
        s->toff[b] = totT;
        int wtok = 0, wq0 = totT;
        for (int k = 0; k < s->C[b]; ++k, ++ci) {
            chunk_clip[ci] = b; chunk_f0[ci] = k * cf;
            int frames = (int)std::min<int64_t>(cf, s->F[b] - (int64_t)k * cf);
            int valid = conv_out_len(conv_out_len(conv_out_len(frames)));                // feat_extract_output_length
            for (int t = 0; t < tpc; ++t) rowmap[(size_t)ci * tpc + t] = t < valid ? totT + t : -1;
            totT += valid; wtok += valid;
            bool close = d.chunks_per_window > 0 && ((k + 1) % d.chunks_per_window == 0);
            if (close || k == s->C[b] - 1) {
                if (d.chunks_per_window == 0) { /* mask None: one window per utterance */ if (k != s->C[b] - 1) continue; }
                win_q0[nwin] = wq0; win_len[nwin] = wtok; maxwin = std::max(maxwin, wtok); ++nwin;
                wq0 = totT; wtok = 0;
            }
        }
        s->T[b] = totT - s->toff[b];
    }
    s->totC = totC; s->totT = totT; s->nwin = nwin; s->maxwin = maxwin;
    const size_t nint = (size_t)(win_len + (totC + B) - hi);
    ASRB_REQUIRE(nint <= s->enc_int_cap, ASRB_ERR_INVALID, "plan exceeds session capacity");
    ASRB_CUDA_CHECK(cudaMemcpyAsync(s->d_int, hi, nint * sizeof(int), cudaMemcpyHostToDevice, s->st));
    s->d_chunk_clip = s->d_int; s->d_chunk_f0 = s->d_int + (chunk_f0 - hi); s->d_rowmap = s->d_int + (rowmap - hi);
    s->d_win_q0 = s->d_int + (win_q0 - hi); s->d_win_len = s->d_int + (win_len - hi);
    const int Bm = s->max_batch; const int np = s->nplanes; cudaStream_t st = s->st;

    // ---- conv stem ----
    launch_conv1(m, s->d_mel, s->d_chunk_clip, s->d_chunk_f0, s->d_i64 + 3 * Bm, s->d_i64 + 4 * Bm, totC, s->act1, s->act1_ps, st);
    {   // conv2d2: implicit GEMM  M = C*32*25, N = dsh, K = 9*cpad
        GemmA A; A.mode = A_CONV; A.a = s->act1; A.plane_stride = s->act1_ps; A.nplanes = np;
        A.OH = d.conv_h[2]; A.OW = d.conv_w[2]; A.Hh = d.conv_h[2]; A.Wh = d.conv_w[2]; A.cpad = d.cpad;
        A.M = totC * A.OH * A.OW; A.K = 9 * d.cpad; A.lda = A.K;
        GemmEpi E; E.mode = EPI_CONV_PARITY; E.bias = m.conv2_b; E.out_s3 = s->act2; E.s3_plane_stride = s->act2_ps;
        E.OH = A.OH; E.OW = A.OW; E.Hh2 = d.conv_h[3]; E.Wh2 = d.conv_w[3]; E.cpad = d.cpad;
        launch_gemm(A, m.conv2_w, c.downsample_hidden_size, E, s->gemm_impl, st);
    }
    {   // conv2d3
        GemmA A; A.mode = A_CONV; A.a = s->act2; A.plane_stride = s->act2_ps; A.nplanes = np;
        A.OH = d.conv_h[3]; A.OW = d.conv_w[3]; A.Hh = d.conv_h[3]; A.Wh = d.conv_w[3]; A.cpad = d.cpad;
        A.M = totC * A.OH * A.OW; A.K = 9 * d.cpad; A.lda = A.K;
        GemmEpi E; E.mode = EPI_CONV_FEAT; E.bias = m.conv3_b; E.out_s3 = s->feat; E.s3_plane_stride = s->feat_ps;
        E.OH = A.OH; E.OW = A.OW; E.lds = d.feat;
        launch_gemm(A, m.conv3_w, c.downsample_hidden_size, E, s->gemm_impl, st);
    }
    {   // conv_out + positional embedding + valid-token gather -> x_enc [totT][d_model]
        GemmA A = plainA(s->feat, s->feat_ps, totC * tpc, d.feat, np);
        GemmEpi E; E.mode = EPI_CONVOUT; E.bias = m.conv_out_b; E.out_f32 = s->x_enc; E.ldo = c.d_model;
        E.splitk_ws = s->splitk_ws; E.extra_launches = &s->launches;
        E.row_map = s->d_rowmap; E.pos = m.pos_emb; E.pos_period = tpc;
        launch_gemm(A, m.conv_out_w, c.d_model, E, s->gemm_impl, st);
    }
    s->launches += 4;
    // ---- transformer layers (layers.rs:230-242) ----
    const int dm = c.d_model;
    for (int l = 0; l < c.encoder_layers; ++l) {
        const EncLayerW& w = m.enc[l];
        launch_layernorm_s3(s->x_enc, w.ln1_w, w.ln1_b, totT, dm, 1e-5f, s->enc_h, s->ench_ps, st);
        { GemmA A = plainA(s->enc_h, s->ench_ps, totT, dm, np);
          GemmEpi E; E.bias = w.bqkv; E.out_f32 = s->enc_qkv; E.ldo = 3 * dm;
          launch_gemm(A, w.wqkv, 3 * dm, E, s->gemm_impl, st); }
        { AttnParams p{}; p.q = s->enc_qkv; p.ldq = 3 * dm; p.k = s->enc_qkv + dm; p.v = s->enc_qkv + 2 * dm;
          p.head_stride = d.enc_hd; p.ldk = 3 * dm; p.seg_stride = 0; p.keys_in_rows = 1;
          p.seg_q0 = s->d_win_q0; p.seg_len = s->d_win_len; p.nseg = nwin; p.nheads = c.encoder_attention_heads; p.group = 1;
          p.causal = 0; p.max_len = maxwin; p.out_s3 = s->enc_attn; p.plane_stride = s->ench_ps; p.ldo = dm;
          launch_attention(p, d.enc_hd, st); }
        { GemmA A = plainA(s->enc_attn, s->ench_ps, totT, dm, np);
          GemmEpi E; E.bias = w.bo; E.residual = s->x_enc; E.ldr = dm; E.out_f32 = s->x_enc; E.ldo = dm; E.splitk_ws = s->splitk_ws; E.extra_launches = &s->launches;
          launch_gemm(A, w.wo, dm, E, s->gemm_impl, st); }
        launch_layernorm_s3(s->x_enc, w.ln2_w, w.ln2_b, totT, dm, 1e-5f, s->enc_h, s->ench_ps, st);
        { GemmA A = plainA(s->enc_h, s->ench_ps, totT, dm, np);
          GemmEpi E; E.bias = w.b1; E.act = 1; E.out_s3 = s->enc_ff; E.s3_plane_stride = s->encff_ps; E.lds = c.encoder_ffn_dim;
          launch_gemm(A, w.fc1, c.encoder_ffn_dim, E, s->gemm_impl, st); }
        { GemmA A = plainA(s->enc_ff, s->encff_ps, totT, c.encoder_ffn_dim, np);
          GemmEpi E; E.bias = w.b2; E.residual = s->x_enc; E.ldr = dm; E.out_f32 = s->x_enc; E.ldo = dm; E.splitk_ws = s->splitk_ws; E.extra_launches = &s->launches;
          launch_gemm(A, w.fc2, dm, E, s->gemm_impl, st); }
        s->launches += 7;
    }
    // ---- ln_post -> proj1 + GELU -> proj2  (audio_encoder.rs:163-165) ----
    launch_layernorm_s3(s->x_enc, m.lnpost_w, m.lnpost_b, totT, dm, 1e-5f, s->enc_h, s->ench_ps, st);
    { GemmA A = plainA(s->enc_h, s->ench_ps, totT, dm, np);
      GemmEpi E; E.bias = m.proj1_b; E.act = 1; E.out_s3 = s->enc_attn; E.s3_plane_stride = s->ench_ps; E.lds = dm;
      launch_gemm(A, m.proj1, dm, E, s->gemm_impl, st); }
    { GemmA A = plainA(s->enc_attn, s->ench_ps, totT, dm, np);
      GemmEpi E; E.bias = m.proj2_b; E.out_f32 = s->audio; E.ldo = c.output_dim; E.splitk_ws = s->splitk_ws; E.extra_launches = &s->launches;
      launch_gemm(A, m.proj2, c.output_dim, E, s->gemm_impl, st); }
    s->launches += 3;
    if (n_tokens_out) for (int b = 0; b < B; ++b) n_tokens_out[b] = s->T[b];
    s->stage = 2;
}

void session_encode_read(Session* s, int b, float* out) {
    ASRB_REQUIRE(s->stage >= 2 && b >= 0 && b < s->B, ASRB_ERR_STATE, "encode_read: nothing encoded for this index");
    ASRB_CUDA_CHECK(cudaStreamSynchronize(s->st));
    const int od = s->m->d.c.output_dim;
    ASRB_CUDA_CHECK(cudaMemcpy(out, s->audio + (size_t)s->toff[b] * od, (size_t)s->T[b] * od * sizeof(float), cudaMemcpyDeviceToHost));
}

// -------------------------------------------------------------------------------------------------
// steps 4-7: prompt, embed + inject, positions, prefill  (inference.rs:105-149)
// -------------------------------------------------------------------------------------------------
void session_prefill(Session* s, const int64_t* const* lang_ids, const int32_t* n_lang_ids, int64_t* seq_lens_out,
                     float* last_logits) {
    ASRB_REQUIRE(s->stage >= 2, ASRB_ERR_STATE, "prefill called before encode");
    Model& m = *s->m; const Dims& d = m.d; const asrb_dims& c = d.c;
    const int B = s->B; cudaStream_t st = s->st; const int np = s->nplanes;
    ASRB_CUDA_CHECK(cudaSetDevice(m.ctx->device));
    s->S.assign(B, 0); s->srow0.assign(B, 0);
    int totS = 0, maxlen = 0;
    for (int b = 0; b < B; ++b) {
        int nl = (lang_ids && lang_ids[b] && n_lang_ids) ? n_lang_ids[b] : 0;
        ASRB_REQUIRE(nl >= 0 && nl <= s->max_lang, ASRB_ERR_INVALID, "language prompt exceeds session capacity");
        s->srow0[b] = totS; s->S[b] = 9 + s->T[b] + 6 + nl; totS += s->S[b]; maxlen = std::max(maxlen, s->S[b]);
    }
    s->totS = totS; s->maxlenS = maxlen;
    int* hi = s->h_int + s->enc_int_cap;      // separate region: the encoder plan upload may still be in flight
    int* di = s->d_int + s->enc_int_cap;
    int* ids = hi; int* arow = ids + totS; int* rseq = arow + totS; int* rpos = rseq + totS;
    int* sq0 = rpos + totS; int* slen = sq0 + B; int* lastrow = slen + B; int* pos0 = lastrow + B;
    for (int b = 0; b < B; ++b) {
        int r = s->srow0[b];
        for (int i = 0; i < 9; ++i, ++r) { ids[r] = kPromptHead[i]; arow[r] = -1; }
        for (int t = 0; t < s->T[b]; ++t, ++r) { ids[r] = kAudioPad; arow[r] = s->toff[b] + t; }
        for (int i = 0; i < 6; ++i, ++r) { ids[r] = kPromptTail[i]; arow[r] = -1; }
        int nl = s->S[b] - (9 + s->T[b] + 6);
        for (int i = 0; i < nl; ++i, ++r) {
            int64_t id = lang_ids[b][i];
            ASRB_REQUIRE(id >= 0 && id < c.vocab_size, ASRB_ERR_INVALID, "language id out of vocabulary");
            ids[r] = (int)id; arow[r] = -1;
        }
        for (int i = 0; i < s->S[b]; ++i) { rseq[s->srow0[b] + i] = b; rpos[s->srow0[b] + i] = i; }   // build_position_ids :259-266
        sq0[b] = s->srow0[b]; slen[b] = s->S[b]; lastrow[b] = s->srow0[b] + s->S[b] - 1; pos0[b] = s->S[b] - 1;
    }
    const size_t nint = (size_t)(pos0 + B - hi);
    ASRB_REQUIRE(s->enc_int_cap + nint <= s->int_cap, ASRB_ERR_INVALID, "plan exceeds session capacity");
    ASRB_CUDA_CHECK(cudaMemcpyAsync(di, hi, nint * sizeof(int), cudaMemcpyHostToDevice, st));
    s->d_ids = di; s->d_audio_row = di + (arow - hi); s->d_row_seq = di + (rseq - hi);
    s->d_row_pos = di + (rpos - hi); s->d_seq_q0 = di + (sq0 - hi); s->d_seq_len = di + (slen - hi);
    ASRB_CUDA_CHECK(cudaMemcpyAsync(s->d_lastrow, di + (lastrow - hi), B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    ASRB_CUDA_CHECK(cudaMemcpyAsync(s->db.pos, di + (pos0 - hi), B * sizeof(int), cudaMemcpyDeviceToDevice, st));
    ASRB_CUDA_CHECK(cudaMemsetAsync(s->db.done, 0, B * sizeof(int), st));
    ASRB_CUDA_CHECK(cudaMemsetAsync(s->db.n_out, 0, B * sizeof(int), st));

    launch_embed_inject(m.embed, c.hidden_size, s->d_ids, s->d_audio_row, s->audio, totS, s->hid, st);
    s->launches += 1;
    const int H = c.hidden_size; const float eps = (float)c.rms_norm_eps;
    for (int l = 0; l < c.num_hidden_layers; ++l) {
        const DecLayerW& w = m.dec[l];
        float* kc = s->kcache + (size_t)l * s->cache_layer_stride;
        float* vc = s->vcache + (size_t)l * s->cache_layer_stride;
        launch_rmsnorm_s3(s->hid, w.ln_in, totS, H, eps, s->dh, s->dh_ps, st);
        { GemmA A = plainA(s->dh, s->dh_ps, totS, H, np);
          GemmEpi E; E.out_f32 = s->dqkv; E.ldo = d.qkv_dim;
          launch_gemm(A, w.wqkv, d.qkv_dim, E, s->gemm_impl, st); }
        launch_qk_norm_rope(s->dqkv, totS, s->d_row_seq, s->d_row_pos, w.qnorm, w.knorm, eps, m.rope_cos, m.rope_sin,
                            c.num_attention_heads, c.num_key_value_heads, c.head_dim, s->qrot, kc, vc, s->cache_seq_stride,
                            s->max_ctx, st);
        { AttnParams p{}; p.q = s->qrot; p.ldq = d.q_dim; p.k = kc; p.v = vc; p.seg_stride = s->cache_seq_stride;
          p.head_stride = (size_t)s->max_ctx * c.head_dim; p.ldk = c.head_dim; p.keys_in_rows = 0;
          p.seg_q0 = s->d_seq_q0; p.seg_len = s->d_seq_len; p.nseg = B; p.nheads = c.num_attention_heads;
          p.group = c.num_attention_heads / c.num_key_value_heads; p.causal = 1; p.max_len = maxlen;
          p.out_s3 = s->dattn; p.plane_stride = s->dattn_ps; p.ldo = d.q_dim;
          launch_attention(p, c.head_dim, st); }
        { GemmA A = plainA(s->dattn, s->dattn_ps, totS, d.q_dim, np);
          GemmEpi E; E.residual = s->hid; E.ldr = H; E.out_f32 = s->hid; E.ldo = H; E.splitk_ws = s->splitk_ws; E.extra_launches = &s->launches;
          launch_gemm(A, w.wo, H, E, s->gemm_impl, st); }
        launch_rmsnorm_s3(s->hid, w.ln_post, totS, H, eps, s->dh, s->dh_ps, st);
        { GemmA A = plainA(s->dh, s->dh_ps, totS, H, np);
          GemmEpi E; E.mode = EPI_SWIGLU; E.out_s3 = s->dact; E.s3_plane_stride = s->dact_ps; E.lds = c.intermediate_size;
          launch_gemm(A, w.wgu, 2 * c.intermediate_size, E, s->gemm_impl, st); }
        { GemmA A = plainA(s->dact, s->dact_ps, totS, c.intermediate_size, np);
          GemmEpi E; E.residual = s->hid; E.ldr = H; E.out_f32 = s->hid; E.ldo = H; E.splitk_ws = s->splitk_ws; E.extra_launches = &s->launches;
          launch_gemm(A, w.wdown, H, E, s->gemm_impl, st); }
        s->launches += 8;
    }
    // final norm + lm_head on the last row of each utterance only (the reference computes all S rows,
    // text_decoder.rs:111-112, and uses row S-1, inference.rs:156)
    launch_lmhead_argmax(m, s->hid, s->d_lastrow, B, s->db, last_logits != nullptr, st, &s->launches);
    // greedy bookkeeping for token 0 (inference.rs:161-170): argmax, EOS check, append, embed
    launch_greedy(m, s->db, B, st, &s->launches);
    s->greedy_done = 1;
    if (seq_lens_out) for (int b = 0; b < B; ++b) seq_lens_out[b] = s->S[b];
    if (last_logits) {
        ASRB_CUDA_CHECK(cudaStreamSynchronize(st));
        ASRB_CUDA_CHECK(cudaMemcpy(last_logits, s->db.logits, (size_t)B * c.vocab_size * sizeof(float), cudaMemcpyDeviceToHost));
    }
    s->stage = 3;
}

// -------------------------------------------------------------------------------------------------
// step 8: greedy loop  (inference.rs:160-200)
// -------------------------------------------------------------------------------------------------
bool decode_mega_supported(const Model& m, int B, int max_ctx);
bool decode_batch_supported(const Model& m, int B, int max_ctx);
size_t decode_batch_part_floats(const Model& m);
void launch_decode_step_batch(const Model& m, const DecodeBufs& b, int B, float* kcache, float* vcache,
                              size_t cache_layer_stride, size_t cache_seq_stride, int max_ctx, int ctx_now, const MegaBufs& mb,
                              cudaStream_t st, int64_t* launches);
void launch_decode_step_mega(const Model& m, const DecodeBufs& b, int B, float* kcache, float* vcache,
                             size_t cache_layer_stride, size_t cache_seq_stride, int max_ctx, int ctx_now, const MegaBufs& mb,
                             cudaStream_t st, int64_t* launches);
size_t decode_mega_part_floats(const Model& m);
int decode_mega_dbg_slots();

// one iteration of the loop body: decoder forward on the pending token, then the greedy bookkeeping
// that selects / appends / embeds the next one.  (The fused kernel does both.)
// upper bound of (position + 1) for the next forward: prompt length + tokens appended so far
static int ctx_bound(const Session* s) { return std::min(s->max_ctx, s->maxlenS + s->greedy_done); }
static bool use_batch(Session* s, bool write_logits) {     // batch >= 2: weights streamed once for all sequences
    return s->decode_mode == 1 && s->batch_step && !write_logits && s->B >= 2 && decode_batch_supported(*s->m, s->B, ctx_bound(s));
}
static bool use_mega(Session* s, bool write_logits) {
    return use_batch(s, write_logits) ||
           (s->decode_mode == 1 && !write_logits && decode_mega_supported(*s->m, s->B, ctx_bound(s)));
}
static void forward_step(Session* s, bool write_logits) {
    Model& m = *s->m;
    if (use_batch(s, write_logits)) {
        launch_decode_step_batch(m, s->db, s->B, s->kcache, s->vcache, s->cache_layer_stride, s->cache_seq_stride, s->max_ctx,
                                 ctx_bound(s), s->mega, s->st, &s->launches);
        s->n_batch_steps += 1;
    } else if (use_mega(s, write_logits)) {
        launch_decode_step_mega(m, s->db, s->B, s->kcache, s->vcache, s->cache_layer_stride, s->cache_seq_stride, s->max_ctx,
                                ctx_bound(s), s->mega, s->st, &s->launches);
        s->n_mega_steps += 1;
    } else {
        s->n_phase_steps += 1;
        launch_decode_step_phases(m, s->db, s->B, s->kcache, s->vcache, s->cache_layer_stride, s->cache_seq_stride, s->max_ctx,
                                  write_logits, s->st, &s->launches);
        launch_greedy(m, s->db, s->B, s->st, &s->launches);
    }
}

void session_decode_step(Session* s, int64_t* next_ids_out, float* logits) {
    ASRB_REQUIRE(s->stage >= 3, ASRB_ERR_STATE, "decode_step called before prefill");
    Model& m = *s->m; const int B = s->B;
    ASRB_CUDA_CHECK(cudaSetDevice(m.ctx->device));
    // the id selected by the previous greedy application is the one this iteration consumes
    ASRB_CUDA_CHECK(cudaMemcpyAsync(s->h_next, s->db.next_id, B * sizeof(int), cudaMemcpyDeviceToHost, s->st));
    forward_step(s, logits != nullptr);
    s->decode_steps += 1; s->greedy_done += 1;
    ASRB_CUDA_CHECK(cudaStreamSynchronize(s->st));
    if (next_ids_out) for (int b = 0; b < B; ++b) next_ids_out[b] = s->h_next[b];
    if (logits) ASRB_CUDA_CHECK(cudaMemcpy(logits, s->db.logits, (size_t)B * m.d.c.vocab_size * sizeof(float), cudaMemcpyDeviceToHost));
}

void session_generate(Session* s, int max_new_tokens, int32_t* ids_out, int32_t* lens_out) {
    ASRB_REQUIRE(s->stage >= 3, ASRB_ERR_STATE, "generate called before prefill");
    ASRB_REQUIRE(max_new_tokens >= 1 && max_new_tokens <= s->max_new, ASRB_ERR_INVALID, "max_new_tokens exceeds session capacity");
    Model& m = *s->m; const int B = s->B; cudaStream_t st = s->st;
    ASRB_CUDA_CHECK(cudaSetDevice(m.ctx->device));
    // Token 0 was selected at the end of prefill; each further token costs one forward + greedy.  The
    // reference also runs `forward` after the last appended token and discards its logits
    // (inference.rs:160-200); that wasted forward is not issued here.
    const int steps = std::max(0, max_new_tokens - s->greedy_done);
    auto ensure_graph = [&]() {   // per-phase path: ~142 launches per step -> replay them as one CUDA graph
        const int mode_key = s->decode_mode * (s->max_batch + 1) + B;     // unique per (mode, batch)
        if (s->step_graph == nullptr || s->graph_mode != mode_key) {
            if (s->step_graph) { cudaGraphExecDestroy(s->step_graph); s->step_graph = nullptr; }
            cudaGraph_t g = nullptr;
            int64_t before = s->launches;
            ASRB_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try { forward_step(s, false); }
            catch (...) { cudaStreamEndCapture(st, &g); if (g) cudaGraphDestroy(g); s->launches = before; throw; }
            ASRB_CUDA_CHECK(cudaStreamEndCapture(st, &g));
            ASRB_CUDA_CHECK(cudaGraphInstantiate(&s->step_graph, g, 0));
            cudaGraphDestroy(g);
            s->graph_mode = mode_key; s->graph_B = (int)(s->launches - before);   // kernels per replay
            s->launches = before;
        }
    };
    const int check_every = 16;
    bool all_done = false;
    for (int it = 0; it < steps && !all_done; ++it) {
        // the fused step covers contexts up to 1152 keys; beyond that (long generations) the per-phase path takes over
        if (use_mega(s, false)) forward_step(s, false);
        else { ensure_graph(); ASRB_CUDA_CHECK(cudaGraphLaunch(s->step_graph, st)); s->launches += s->graph_B; }
        s->decode_steps += 1; s->greedy_done += 1;
        if ((it + 1) % check_every == 0 && it + 1 < steps) {
            ASRB_CUDA_CHECK(cudaMemcpyAsync(s->h_done, s->db.done, B * sizeof(int), cudaMemcpyDeviceToHost, st));
            ASRB_CUDA_CHECK(cudaStreamSynchronize(st));
            all_done = true;
            for (int b = 0; b < B; ++b) all_done = all_done && s->h_done[b];
        }
    }
    ASRB_CUDA_CHECK(cudaMemcpyAsync(s->h_nout, s->db.n_out, B * sizeof(int), cudaMemcpyDeviceToHost, st));
    ASRB_CUDA_CHECK(cudaMemcpyAsync(s->h_ids, s->db.ids_out, (size_t)B * s->max_new * sizeof(int), cudaMemcpyDeviceToHost, st));
    ASRB_CUDA_CHECK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b) {
        int n = std::min(s->h_nout[b], max_new_tokens);
        lens_out[b] = n;
        for (int i = 0; i < n; ++i) ids_out[(size_t)b * max_new_tokens + i] = s->h_ids[(size_t)b * s->max_new + i];
    }
}

void session_transcribe_ids(Session* s, const float* const* samples, const int64_t* n_samples, int batch,
                            const int64_t* const* lang_ids, const int32_t* n_lang_ids, int max_new_tokens,
                            int32_t* ids_out, int32_t* lens_out) {
    ASRB_REQUIRE(ids_out && lens_out, ASRB_ERR_INVALID, "null output");
    if (samples == nullptr) batch = (int)s->ingested_n.size();      // asrb_transcribe_ingested
    ASRB_REQUIRE(max_new_tokens >= 1 && max_new_tokens <= s->max_new, ASRB_ERR_INVALID, "max_new_tokens exceeds session capacity");
    cudaStream_t st = s->st;
    s->launches = 0; s->decode_steps = 0;
    ASRB_CUDA_CHECK(cudaEventRecord(s->ev[0], st));
    s->timing = true;
    try { session_mel(s, samples, n_samples, batch, nullptr); } catch (...) { s->timing = false; throw; }   // records ev[1] after the H2D
    s->timing = false;
    ASRB_CUDA_CHECK(cudaEventRecord(s->ev[2], st));
    session_encode(s, nullptr);
    ASRB_CUDA_CHECK(cudaEventRecord(s->ev[3], st));
    session_prefill(s, lang_ids, n_lang_ids, nullptr, nullptr);
    ASRB_CUDA_CHECK(cudaEventRecord(s->ev[4], st));
    session_generate(s, max_new_tokens, ids_out, lens_out);
    ASRB_CUDA_CHECK(cudaEventRecord(s->ev[5], st));
    ASRB_CUDA_CHECK(cudaEventSynchronize(s->ev[5]));
    for (int i = 0; i < 5; ++i) ASRB_CUDA_CHECK(cudaEventElapsedTime(&s->last_ms[i], s->ev[i], s->ev[i + 1]));
    ASRB_CUDA_CHECK(cudaEventElapsedTime(&s->last_ms[5], s->ev[0], s->ev[5]));
}

void session_last_timings(Session* s, float* ms6, int64_t* kernels, int64_t* steps) {
    if (ms6) for (int i = 0; i < 6; ++i) ms6[i] = s->last_ms[i];
    if (kernels) *kernels = s->launches;
    if (steps) *steps = s->decode_steps;
}

void session_device_ids(Session* s, const int32_t** ids, const int32_t** lens, int* stride, int* batch) {
    ASRB_REQUIRE(s->stage >= 3, ASRB_ERR_STATE, "device_ids: nothing generated yet");
    *ids = s->db.ids_out; *lens = s->db.n_out; *stride = s->max_new; *batch = s->B;
}

void session_stats(Session* s, int64_t* out, int n) {
    const int64_t v[5] = {s->n_batch_steps, s->n_mega_steps, s->n_phase_steps, g_gemm_simt_fallbacks.load(), g_gemm_tc_launches.load()};
    for (int i = 0; i < n && i < 5; ++i) out[i] = v[i];
}

void session_set_option(Session* s, const char* key, const char* value) {
    std::string k(key ? key : ""), v(value ? value : "");
    if (s->step_graph) { cudaGraphExecDestroy(s->step_graph); s->step_graph = nullptr; s->graph_mode = -1; }   // captured with the old options
    if (k == "gemm") {
        if (v == "tc") s->gemm_impl = GEMM_TC; else if (v == "simt") s->gemm_impl = GEMM_SIMT;
        else throw Error(ASRB_ERR_INVALID, "gemm must be tc|simt");
    } else if (k == "decode") {
        if (v == "mega") s->decode_mode = 1; else if (v == "phases") s->decode_mode = 0;
        else throw Error(ASRB_ERR_INVALID, "decode must be mega|phases");
    } else if (k == "batch_step") {
        s->batch_step = (v == "1");
    } else if (k == "planes") {
        int p = atoi(v.c_str());
        ASRB_REQUIRE(p >= 1 && p <= 3, ASRB_ERR_INVALID, "planes must be 1..3");
        s->nplanes = p;
    } else if (k == "resident") {
        s->resident = (v == "1");
    } else throw Error(ASRB_ERR_INVALID, "unknown option: " + k);
}

}  // namespace asrb
