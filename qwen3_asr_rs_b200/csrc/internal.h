// internal.h -- host-side structures and kernel launch prototypes (not part of the ABI).
#pragma once
#include <atomic>
#include <map>
#include <string>
#include <vector>
#include "../../include/asr_b200.h"
#include "common.cuh"

namespace asrb {

// ---------------------------------------------------------------------------------------
// dims: asrb_dims + derived geometry (reference: src/config.rs, src/audio_encoder.rs:79-134)
// ---------------------------------------------------------------------------------------
struct Dims {
    asrb_dims c;
    int enc_hd;             // d_model / heads
    int chunk_frames;       // 2 * n_window                       (audio_encoder.rs:83)
    int chunks_per_window;  // n_window_infer / chunk_frames      (audio_encoder.rs:179)
    int conv_h[4];          // freq extent  128 -> 64 -> 32 -> 16
    int conv_w[4];          // time extent  100 -> 50 -> 25 -> 13
    int tok_per_chunk;      // conv_w[3]
    int cpad;               // downsample_hidden_size rounded up to 64 (TMA / UMMA K-block)
    int feat;               // dsh * conv_h[3]  (conv_out in-features, 7680)
    int q_dim, kv_dim, qkv_dim;
    void derive();
};
void validate_dims(const asrb_dims& d);   // c_api.cu: throws ASRB_ERR_INVALID
inline int conv_out_len(int l) { return (l - 1) / 2 + 1; }   // audio_encoder.rs:263-266

struct Ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
};

struct EncLayerW {
    float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
    bf16 *wqkv, *wo, *fc1, *fc2;
    float *bqkv, *bo, *b1, *b2;
};
struct DecLayerW {
    float *ln_in, *ln_post, *qnorm, *knorm;
    bf16 *wqkv, *wo, *wgu, *wdown;   // wgu rows interleaved: 2j = gate_j, 2j+1 = up_j
};

struct RawTensor {
    void* dev = nullptr;       // bf16 for matrices, f32 for vectors
    bool is_bf16 = false;
    std::vector<int64_t> shape;
    size_t numel = 0;
};

struct Model {
    Ctx* ctx = nullptr;
    Dims d;
    bool finalized = false;
    bool lossy_weights = false;          // an F32/F16 matrix was not bf16-representable (only with ASRB_ALLOW_LOSSY_WEIGHTS=1)
    int lossy_count = 0;
    std::map<std::string, RawTensor> raw;
    std::vector<void*> owned;            // packed buffers created at finalize

    // mel constants (src/mel.rs:115-187 + periodic Hann + DFT twiddles)
    float *mel_fb = nullptr, *dft_cos = nullptr, *dft_sin = nullptr, *hann = nullptr, *dft_tw = nullptr;
    int* mel_krange = nullptr;           // [num_mels][2] non-zero bin range of each filter
    // encoder
    float *conv1_w = nullptr, *conv1_b = nullptr, *conv2_b = nullptr, *conv3_b = nullptr, *conv_out_b = nullptr;
    bf16 *conv2_w = nullptr, *conv3_w = nullptr, *conv_out_w = nullptr;   // conv: [dsh][9][cpad]
    float* pos_emb = nullptr;            // [tok_per_chunk][d_model]
    std::vector<EncLayerW> enc;
    float *lnpost_w = nullptr, *lnpost_b = nullptr, *proj1_b = nullptr, *proj2_b = nullptr;
    bf16 *proj1 = nullptr, *proj2 = nullptr;
    // decoder
    bf16 *embed = nullptr, *lm_head = nullptr;
    std::vector<DecLayerW> dec;
    float* final_norm = nullptr;
    DecLayerW* d_dec_layers = nullptr;   // device copy of `dec` for the fused decode step; its ln_in / ln_post point at
                                         // copies stored in the step's bank-conflict-free activation layout (decode_mega.cu)
    float* final_norm_sw = nullptr;      // final norm weight in the same layout
    // batch-aware fused step (decode_batch.cu): weight matrices with the 16-byte chunks of every row XOR-swizzled by
    // (row & 7) -- bank-conflict-free ldmatrix on bulk-copied rows -- and plain norm vectors; null for unsupported dims
    DecLayerW* d_dec_layers_b = nullptr;
    bf16* lm_head_b = nullptr;
    float *rope_cos = nullptr, *rope_sin = nullptr;   // [rope_max_pos][head_dim/2]
    int rope_max_pos = 0;

    ~Model();
};

// ---------------------------------------------------------------------------------------
// GEMM plumbing shared by the SIMT and tcgen05 implementations
//   D[m][n] = sum_k A(m,k) * W[n][k]      W: bf16 [N][K] row-major (HF layout, layers.rs:74-80)
// ---------------------------------------------------------------------------------------
enum { A_PLAIN = 0, A_CONV = 1 };
struct GemmA {
    int mode = A_PLAIN;
    const bf16* a = nullptr;     // split3 planes
    size_t plane_stride = 0;     // elements between planes
    int nplanes = 3;
    int M = 0, K = 0, lda = 0;
    // A_CONV: implicit 3x3 / stride 2 / pad 1 conv over the parity-split channels-last layout
    //   in[(((chunk*2+ph)*2+pw)*Hh + hh)*Wh + wh][cpad],  m = (chunk, oh, ow),  k = (tap, cin)
    int OH = 0, OW = 0, Hh = 0, Wh = 0, cpad = 0;
};
enum { EPI_PLAIN = 0, EPI_SWIGLU = 1, EPI_CONV_PARITY = 2, EPI_CONV_FEAT = 3, EPI_CONVOUT = 4 };
struct GemmEpi {
    int mode = EPI_PLAIN;
    const float* bias = nullptr;      // [N]
    int act = 0;                      // 1 = exact-erf GELU
    const float* residual = nullptr;  // fp32 [M][ldr]
    int ldr = 0;
    float* out_f32 = nullptr;
    int ldo = 0;
    bf16* out_s3 = nullptr;
    size_t s3_plane_stride = 0;
    int lds = 0;
    // conv epilogues
    int OH = 0, OW = 0, Hh2 = 0, Wh2 = 0, cpad = 0;
    // EPI_CONVOUT
    const int* row_map = nullptr;     // [M] -> token row or -1
    const float* pos = nullptr;       // [pos_period][N]
    int pos_period = 1;
    // optional fp32 workspace for split-K (EPI_PLAIN GEMMs with too few tiles to fill the GPU): >= SPLITK_WS_FLOATS floats
    float* splitk_ws = nullptr;
    int64_t* extra_launches = nullptr;   // incremented by the number of kernels launched beyond the one GEMM kernel
    long long* dbg = nullptr;            // ASRB_GEMM_DEBUG timeline (gemm_tc.cu), null in production
};
static constexpr size_t SPLITK_WS_FLOATS = (size_t)4 * 64 * 128 * 128;   // 4 splits x 64 tiles of 128 x 128
enum { GEMM_SIMT = 0, GEMM_TC = 1 };
extern std::atomic<int64_t> g_gemm_simt_fallbacks, g_gemm_tc_launches;   // gemm_simt.cu: tcgen05 requested but SIMT ran / tcgen05 launches
void launch_gemm(const GemmA& A, const bf16* W, int N, const GemmEpi& E, int impl, cudaStream_t st);
// tcgen05 implementation (gemm_tc.cu); returns false when the shape is unsupported
bool launch_gemm_tc(const GemmA& A, const bf16* W, int N, const GemmEpi& E, cudaStream_t st);
void launch_gemm_simt(const GemmA& A, const bf16* W, int N, const GemmEpi& E, cudaStream_t st);

// ---------------------------------------------------------------------------------------
// kernels (each .cu exposes launchers; all asynchronous on `st`)
// ---------------------------------------------------------------------------------------
// mel.cu
void launch_mel(const Model& m, const float* samples, const int64_t* d_soff, const int64_t* d_n,
                const int64_t* d_npad, const int64_t* d_foff, int batch, int max_frames,
                float* mel_out, int* d_maxkey, cudaStream_t st);
// elementwise.cu
void launch_conv1(const Model& m, const float* mel, const int* d_chunk_clip, const int* d_chunk_f0,
                  const int64_t* d_foff, const int64_t* d_frames, int n_chunks,
                  bf16* out_s3, size_t plane_stride, cudaStream_t st);
void launch_layernorm_s3(const float* x, const float* w, const float* b, int rows, int dim, float eps,
                         bf16* out_s3, size_t plane_stride, cudaStream_t st);
void launch_rmsnorm_s3(const float* x, const float* w, int rows, int dim, float eps,
                       bf16* out_s3, size_t plane_stride, cudaStream_t st);
void launch_embed_inject(const bf16* embed, int hidden, const int* d_ids, const int* d_audio_row,
                         const float* audio, int rows, float* out, cudaStream_t st);
void launch_qk_norm_rope(const float* qkv, int rows, const int* d_row_seq, const int* d_row_pos,
                         const float* qnorm, const float* knorm, float eps,
                         const float* rope_cos, const float* rope_sin,
                         int nq, int nkv, int hd, float* q_out, float* kcache, float* vcache,
                         size_t cache_seq_stride, int max_ctx, cudaStream_t st);
// attention.cu
struct AttnParams {
    const float* q; int ldq;            // q row r, head h at q + r*ldq + h*hd
    const float* k; const float* v;     // key j of segment s, kv-head g at k + s*seg_stride + g*head_stride + j*ldk
    size_t seg_stride, head_stride; int ldk;
    const int* seg_q0;                  // [nseg] first q row of segment (also index base of keys when keys_in_rows)
    const int* seg_len;                 // [nseg]
    int keys_in_rows;                   // 1: key j lives at row (seg_q0+j) of the k/v buffers (encoder qkv buffer)
    int nseg, nheads, group;            // q head h uses kv head h / group
    int causal; int max_len;
    bf16* out_s3; size_t plane_stride; int ldo;
};
void launch_attention(const AttnParams& p, int hd, cudaStream_t st);
// decode.cu  (per-phase kernels, any batch <= 8) -- see decode_mega.cu for the fused step
struct DecodeBufs {
    float* x;        // [B][H] residual stream
    float* qkv;      // [B][qkv_dim]
    float* attn;     // [B][q_dim]
    float* act;      // [B][I]
    float* logits;   // [B][V] or null
    float* part_val; int* part_idx; int n_part;   // argmax partials [B][n_part]
    int* pos;        // [B] position of the token being processed (= ctx length before append)
    int* done;       // [B]
    int* next_id;    // [B]
    int* ids_out;    // [B][max_new]
    int* n_out;      // [B]
    int max_new;
};
void launch_decode_step_phases(const Model& m, const DecodeBufs& b, int B, float* kcache, float* vcache,
                               size_t cache_layer_stride, size_t cache_seq_stride, int max_ctx,
                               bool write_logits, cudaStream_t st, int64_t* launches);
// final-norm + lm_head + argmax on arbitrary rows of a residual stream (prefill last rows);
// also performs the greedy bookkeeping of src/inference.rs:161-170 (EOS check, append, embed)
struct MegaBufs { unsigned* bar = nullptr; float* part = nullptr; long long* dbg = nullptr; size_t part_bytes = 0; unsigned* steps_issued = nullptr;
                  uint32_t* sx = nullptr; size_t sx_bytes = 0; int* sx_nb = nullptr; };   // sx: decode_batch.cu's self-validating words   // per-session state of the fused step
size_t decode_mega_part_floats(const Model& m);
size_t decode_batch_part_floats(const Model& m);
size_t decode_batch_sx_bytes(const Model& m);   // decode_batch.cu: NB sequences per fused launch
bool decode_batch_supported(const Model& m, int B, int ctx);
int decode_mega_dbg_slots();
void launch_greedy(const Model& m, const DecodeBufs& b, int B, cudaStream_t st, int64_t* launches);
void launch_lmhead_argmax(const Model& m, const float* x_rows, const int* d_row_idx, int B,
                          const DecodeBufs& b, bool write_logits, cudaStream_t st, int64_t* launches);

}  // namespace asrb
