// loader.cu -- AsrInference::load for the hot path: config.json + safetensors -> device model.
//
// Reference: AsrConfig::from_file (/root/reference/src/config.rs:116-120, serde defaults :27-113),
// load_model_weights (src/weights.rs:10-58: model.safetensors, else model.safetensors.index.json +
// shards) and load_safetensors (src/weights.rs:62-120).  The reference widens every tensor to f32 on
// the host; here bf16 matrices stay bf16 (model.cu).  Files are mmap'ed and handed to
// model_set_tensor without an intermediate copy.  A ~100-line JSON reader is enough for both
// config.json and the safetensors header.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <cstring>
#include <fstream>
#include <memory>
#include <set>
#include <sstream>
#include "internal.h"

namespace asrb {

void model_set_tensor(Model* m, const char* name, int dtype, const int64_t* shape, int ndim, const void* host);
void model_finalize(Model* m);

namespace {

struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false; double num = 0; std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const std::string& k) const {
        if (kind != Obj) return nullptr;
        for (auto& kv : obj) if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p; const char* e;
    [[noreturn]] void fail(const char* what) { throw Error(ASRB_ERR_IO, std::string("JSON parse error: ") + what); }
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    JVal parse() { ws(); JVal v = value(); ws(); return v; }
    JVal value() {
        if (p >= e) fail("unexpected end");
        JVal v;
        char c = *p;
        if (c == '{') {
            v.kind = JVal::Obj; ++p; ws();
            if (p < e && *p == '}') { ++p; return v; }
            for (;;) {
                ws(); if (p >= e || *p != '"') fail("expected key");
                std::string k = string(); ws();
                if (p >= e || *p != ':') fail("expected ':'");
                ++p; ws();
                v.obj.emplace_back(std::move(k), value()); ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = JVal::Arr; ++p; ws();
            if (p < e && *p == ']') { ++p; return v; }
            for (;;) {
                ws(); v.arr.push_back(value()); ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') { v.kind = JVal::Str; v.str = string(); }
        else if (c == 't' && e - p >= 4 && !strncmp(p, "true", 4)) { v.kind = JVal::Bool; v.b = true; p += 4; }
        else if (c == 'f' && e - p >= 5 && !strncmp(p, "false", 5)) { v.kind = JVal::Bool; v.b = false; p += 5; }
        else if (c == 'n' && e - p >= 4 && !strncmp(p, "null", 4)) { p += 4; }
        else {
            char* end = nullptr;
            v.kind = JVal::Num; v.num = strtod(p, &end);
            if (end == p) fail("bad value");
            p = end;
        }
        return v;
    }
    std::string string() {
        std::string s; ++p;
        while (p < e && *p != '"') {
            if (*p == '\\' && p + 1 < e) {
                ++p;
                switch (*p) {
                    case 'n': s += '\n'; break; case 't': s += '\t'; break; case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break; case 'f': s += '\f'; break;
                    case 'u': { if (e - p < 5) fail("bad \\u"); unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                                if (cp < 0x80) s += (char)cp; else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
                                else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
                                p += 4; break; }
                    default: s += *p;
                }
                ++p;
            } else s += *p++;
        }
        if (p >= e) fail("unterminated string");
        ++p; return s;
    }
};

std::string read_text(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(ASRB_ERR_IO, "cannot open " + path);
    std::stringstream ss; ss << f.rdbuf(); return ss.str();
}
bool exists(const std::string& path) { struct stat st; return stat(path.c_str(), &st) == 0; }

struct Mapped {
    void* base = nullptr; size_t size = 0; int fd = -1;
    explicit Mapped(const std::string& path) {
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) throw Error(ASRB_ERR_IO, "cannot open " + path);
        struct stat st; if (fstat(fd, &st) != 0) { close(fd); throw Error(ASRB_ERR_IO, "cannot stat " + path); }
        size = (size_t)st.st_size;
        base = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (base == MAP_FAILED) { close(fd); throw Error(ASRB_ERR_IO, "cannot mmap " + path); }
    }
    ~Mapped() { if (base && base != MAP_FAILED) munmap(base, size); if (fd >= 0) close(fd); }
};

void geti(const JVal* o, const char* k, int32_t& dst) { if (o) if (const JVal* v = o->get(k)) if (v->kind == JVal::Num) dst = (int32_t)v->num; }
void getd(const JVal* o, const char* k, double& dst) { if (o) if (const JVal* v = o->get(k)) if (v->kind == JVal::Num) dst = v->num; }

void load_safetensors_file(Model* m, const std::string& path) {          // weights.rs:62-120
    Mapped mp(path);
    ASRB_REQUIRE(mp.size >= 8, ASRB_ERR_IO, "safetensors file too small: " + path);
    uint64_t hlen = 0; memcpy(&hlen, mp.base, 8);
    ASRB_REQUIRE(hlen <= mp.size - 8, ASRB_ERR_IO, "bad safetensors header length: " + path);
    const char* hp = (const char*)mp.base + 8;
    JParser jp{hp, hp + hlen};
    JVal hdr = jp.parse();
    ASRB_REQUIRE(hdr.kind == JVal::Obj, ASRB_ERR_IO, "safetensors header is not an object");
    const uint8_t* data = (const uint8_t*)mp.base + 8 + hlen;
    const size_t data_size = mp.size - 8 - hlen;
    for (auto& kv : hdr.obj) {
        if (kv.first == "__metadata__") continue;
        const JVal* dt = kv.second.get("dtype"); const JVal* sh = kv.second.get("shape"); const JVal* off = kv.second.get("data_offsets");
        ASRB_REQUIRE(dt && sh && off && off->arr.size() == 2, ASRB_ERR_IO, "malformed tensor entry: " + kv.first);
        int code; size_t esz;
        if (dt->str == "BF16") { code = ASRB_DT_BF16; esz = 2; }
        else if (dt->str == "F16") { code = ASRB_DT_F16; esz = 2; }
        else if (dt->str == "F32") { code = ASRB_DT_F32; esz = 4; }
        else if (dt->str == "I64") continue;                               // accepted by the reference, unused by the path
        else throw Error(ASRB_ERR_IO, "unsupported dtype " + dt->str + " for " + kv.first);
        if (kv.first.rfind("thinker.", 0) != 0) continue;                  // tensors outside the thinker are never read
        int64_t shape[4]; int nd = (int)sh->arr.size(); size_t numel = 1;
        if (nd == 0) continue;
        ASRB_REQUIRE(nd <= 4, ASRB_ERR_IO, "rank > 4: " + kv.first);
        for (int i = 0; i < nd; ++i) { shape[i] = (int64_t)sh->arr[i].num; numel *= (size_t)shape[i]; }
        const size_t b = (size_t)off->arr[0].num, e = (size_t)off->arr[1].num;
        ASRB_REQUIRE(e <= data_size && b <= e && e - b == numel * esz, ASRB_ERR_IO, "bad data_offsets: " + kv.first);
        model_set_tensor(m, kv.first.c_str(), code, shape, nd, data + b);
    }
}

}  // namespace

void model_load_dir(Ctx* ctx, const char* dir_c, Model** out) {
    Model* m = *out;
    const std::string dir(dir_c);
    // ---- config.json (config.rs) ----
    asrb_dims d;
    {   // defaults == 0.6B (config.rs:52-62, 90-99)
        d.d_model = 896; d.encoder_layers = 18; d.encoder_attention_heads = 14; d.encoder_ffn_dim = 3584; d.num_mel_bins = 128;
        d.max_source_positions = 1500; d.n_window = 50; d.n_window_infer = 800; d.downsample_hidden_size = 480; d.output_dim = 1024;
        d.vocab_size = 151936; d.hidden_size = 1024; d.intermediate_size = 3072; d.num_hidden_layers = 28; d.num_attention_heads = 16;
        d.num_key_value_heads = 8; d.head_dim = 128; d.tie_word_embeddings = 1; d.rms_norm_eps = 1e-6; d.rope_theta = 1000000.0;
    }
    const std::string cfg_text = read_text(dir + "/config.json");
    JParser jp{cfg_text.data(), cfg_text.data() + cfg_text.size()};
    JVal cfg = jp.parse();
    const JVal* th = cfg.get("thinker_config");
    ASRB_REQUIRE(th && th->kind == JVal::Obj, ASRB_ERR_IO, "config.json: missing thinker_config");
    const JVal* a = th->get("audio_config"); const JVal* t = th->get("text_config");
    geti(a, "d_model", d.d_model); geti(a, "encoder_layers", d.encoder_layers); geti(a, "encoder_attention_heads", d.encoder_attention_heads);
    geti(a, "encoder_ffn_dim", d.encoder_ffn_dim); geti(a, "num_mel_bins", d.num_mel_bins); geti(a, "max_source_positions", d.max_source_positions);
    geti(a, "n_window", d.n_window); geti(a, "n_window_infer", d.n_window_infer); geti(a, "downsample_hidden_size", d.downsample_hidden_size);
    geti(a, "output_dim", d.output_dim);
    geti(t, "vocab_size", d.vocab_size); geti(t, "hidden_size", d.hidden_size); geti(t, "intermediate_size", d.intermediate_size);
    geti(t, "num_hidden_layers", d.num_hidden_layers); geti(t, "num_attention_heads", d.num_attention_heads);
    geti(t, "num_key_value_heads", d.num_key_value_heads); geti(t, "head_dim", d.head_dim);
    getd(t, "rms_norm_eps", d.rms_norm_eps); getd(t, "rope_theta", d.rope_theta);
    if (t) if (const JVal* tw = t->get("tie_word_embeddings")) if (tw->kind == JVal::Bool) d.tie_word_embeddings = tw->b ? 1 : 0;
    validate_dims(d);                      // a config.json with n_window = 0 etc. must be a status code, not a SIGFPE
    m->ctx = ctx; m->d.c = d; m->d.derive();
    // ---- weights (weights.rs:10-58) ----
    const std::string single = dir + "/model.safetensors", index = dir + "/model.safetensors.index.json";
    if (exists(single)) {
        load_safetensors_file(m, single);
    } else if (exists(index)) {
        const std::string it = read_text(index);
        JParser ip{it.data(), it.data() + it.size()};
        JVal idx = ip.parse();
        const JVal* wm = idx.get("weight_map");
        ASRB_REQUIRE(wm && wm->kind == JVal::Obj, ASRB_ERR_IO, "Missing weight_map in index");
        std::set<std::string> shards;
        for (auto& kv : wm->obj) if (kv.second.kind == JVal::Str) shards.insert(kv.second.str);
        for (auto& sname : shards) load_safetensors_file(m, dir + "/" + sname);
    } else {
        throw Error(ASRB_ERR_IO, "No model weights found in " + dir + " (expected model.safetensors or model.safetensors.index.json)");
    }
    model_finalize(m);
}

}  // namespace asrb
