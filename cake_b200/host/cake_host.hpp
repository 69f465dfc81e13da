// cake_host.hpp — C++ host side above the C ABI (include/cake_b200.h), mirroring the reference's interface for the
// block-forward path with the reference's names, argument meaning and error behaviour:
//
//   Forwarder        cake-core/src/cake/mod.rs:510-556      load / forward / forward_mut / forward_batch / goodbye /
//                                                            layer_name / ident
//   Transformer      models/common/transformer.rs:14-150    the local block (here backed by a cake_b200_block)
//   Cache            models/common/cache.rs:9-254           clear / as_new
//   Context          cake/mod.rs:41-65                      config, dtype, device, var_builder, cache
//   VarBuilder       utils/mod.rs:255-384                   mmapped safetensors (single file or index.json shards)
//   TextModelBase    models/common/text_model.rs:133-530    load / forward / prepare_prompt / next_token / reset
//   Master           cake/sharding/master.rs:14-191         generate_text and its tok/s definition
//
// cake is Rust and no Rust toolchain exists in the authoring environment, so the compiled host side is C++.
// Errors are exceptions carrying cake_b200_last_error() (the reference returns anyhow::Result with context).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cake_b200.h"

namespace cake_host {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};
inline void check(int rc, const std::string &what) {
  if (rc != CAKE_B200_OK) throw Error(what + ": " + cake_b200_last_error());
}

// ---------------------------------------------------------------------------------------------- tiny JSON
struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  double num = 0;
  bool b = false;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;
  const Json *get(const std::string &k) const {
    for (auto &kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  double number(const std::string &k, double dflt) const {
    const Json *j = get(k);
    return (j && j->kind == Num) ? j->num : dflt;
  }
  bool boolean(const std::string &k, bool dflt) const {
    const Json *j = get(k);
    return (j && j->kind == Bool) ? j->b : dflt;
  }
};
class JsonParser {
  const char *p, *e;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
  [[noreturn]] void fail(const char *m) { throw Error(std::string("json: ") + m); }
  std::string string() {
    if (*p != '"') fail("expected string");
    p++;
    std::string s;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) {
        p++;
        s += (*p == 'n') ? '\n' : (*p == 't') ? '\t' : *p;
      } else s += *p;
      p++;
    }
    p++;
    return s;
  }
 public:
  JsonParser(const char *b, size_t n) : p(b), e(b + n) {}
  Json value() {
    ws();
    Json j;
    if (p >= e) fail("eof");
    if (*p == '{') {
      j.kind = Json::Obj;
      p++;
      ws();
      while (*p != '}') {
        ws();
        std::string k = string();
        ws();
        if (*p != ':') fail("expected ':'");
        p++;
        j.obj.emplace_back(k, value());
        ws();
        if (*p == ',') p++;
        ws();
      }
      p++;
    } else if (*p == '[') {
      j.kind = Json::Arr;
      p++;
      ws();
      while (*p != ']') {
        j.arr.push_back(value());
        ws();
        if (*p == ',') p++;
        ws();
      }
      p++;
    } else if (*p == '"') {
      j.kind = Json::Str;
      j.str = string();
    } else if (!strncmp(p, "true", 4)) { j.kind = Json::Bool; j.b = true; p += 4;
    } else if (!strncmp(p, "false", 5)) { j.kind = Json::Bool; p += 5;
    } else if (!strncmp(p, "null", 4)) { p += 4;
    } else {
      char *end;
      j.kind = Json::Num;
      j.num = strtod(p, &end);
      if (end == p) fail("bad number");
      p = end;
    }
    return j;
  }
};
inline std::string slurp(const std::string &path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw Error("can't read " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

// ---------------------------------------------------------------------------------------------- config
// models/common/config.rs:87-150 via the into_config of the dense architectures whose block is this path
struct Config {
  cake_b200_config c{};
  std::string model_prefix = "model";
  std::vector<uint32_t> eos;
  std::string arch;
  bool fused_qkv_proj = false, fused_gate_up_proj = false;  // attention.rs:90-94, mlp.rs:38-40
  // the sibling block structures (models/{olmo2,gemma3,exaone4}/block.rs)
  std::string block_kind = "llama";   // "llama" (common/transformer.rs) | "olmo2" | "gemma3" | "exaone4"
  std::vector<bool> global_layers;    // config.rs:126-130: per-layer schedule, true = global (full context)
  bool residual_rms_norm = false;     // config.rs:113: norm weights stored as deltas, (1 + w) in f32 at load (config.rs:155-173)
  int layer_window = 0;               // the local layers' sliding window (already 0 when it can never bite)
  struct Variant { bool pre_norms, post_norms; int window; bool no_rope; };
  // norm placement and attention mode of layer i, as each block.rs sets them up (see cake_b200_block_set_variant)
  Variant layer_variant(int i) const {
    const bool g = i < (int)global_layers.size() && global_layers[i];
    if (block_kind == "olmo2") return {false, true, -1, false};                        // olmo2/block.rs:62-90
    if (block_kind == "gemma3") return {true, true, g ? 0 : layer_window, !g};          // gemma3/block.rs:60-66: local = window, no RoPE
    if (block_kind == "exaone4") return {true, false, g ? 0 : layer_window, g};         // exaone4/block.rs:50-58: global = no RoPE
    return {true, false, -1, false};
  }
  bool standard_blocks() const { return block_kind == "llama"; }
  static Config from_path(const std::string &path, int dtype, int max_seq_override = 0) {
    std::string txt = slurp(path);
    Json j = JsonParser(txt.data(), txt.size()).value();
    Config k;
    if (const Json *a = j.get("architectures"))  // config.rs:175-190: the first *string* entry
      for (auto &e : a->arr)
        if (e.kind == Json::Str) { k.arch = e.str; break; }
    // Per-architecture serde defaults and hard-wired flags of each into_config (llama3/config.rs:62-98,
    // qwen2/config.rs:69-105, qwen3/config.rs:55-93, mistral/config.rs:56-93, falcon3/config.rs:53-90).
    // Unknown strings fall back to Llama (cake/mod.rs:81-109); known architectures with another block are refused.
    struct Arch { const char *name; double rope; int max_pos; bool bias, qk_norm, head_dim, window, phi; };
    static const Arch archs[] = {
        {"LlamaForCausalLM", 500000.0, 4096, false, false, false, false, false},
        {"Qwen2ForCausalLM", 1000000.0, 32768, true, false, false, false, false},
        {"Qwen3ForCausalLM", 1000000.0, 40960, false, true, true, false, false},
        {"MistralForCausalLM", 1000000.0, 131072, false, false, true, true, false},
        {"FalconForCausalLM", 500000.0, 131072, false, false, true, false, false},
        // phi4/config.rs:63-100: pre-fused qkv_proj / gate_up_proj tensors, partial rotary
        {"Phi3ForCausalLM", 1000000.0, 131072, false, false, true, false, true},
        {"Phi4ForCausalLM", 1000000.0, 131072, false, false, true, false, true},
    };
    static const char *other_blocks[] = {"Qwen3_5ForConditionalGeneration", "Qwen3MoeForCausalLM",
        "Qwen3_5MoeForConditionalGeneration", "LuxTTSForTextToSpeech"};
    // cake/mod.rs:99-105: the dense sibling blocks; serde defaults of models/{gemma3,olmo2,exaone4}/config.rs
    static const Arch siblings[] = {
        {"Gemma3ForCausalLM", 10000.0, 131072, false, true, true, false, false},
        {"OLMo2ForCausalLM", 500000.0, 4096, false, true, true, false, false},
        {"Olmo2ForCausalLM", 500000.0, 4096, false, true, true, false, false},
        {"ExaoneForCausalLM", 500000.0, 131072, false, true, true, false, false},
    };
    for (const char *o : other_blocks)
      if (k.arch == o) throw Error("architecture " + k.arch + " is outside the block-forward path built here");
    const Arch *ar = &archs[0];
    for (auto &a : archs)
      if (k.arch == a.name) ar = &a;
    for (auto &a : siblings)
      if (k.arch == a.name) {
        ar = &a;
        k.block_kind = k.arch == "Gemma3ForCausalLM" ? "gemma3" : k.arch == "ExaoneForCausalLM" ? "exaone4" : "olmo2";
      }
    auto &c = k.c;
    // the fields every *Config struct of the reference declares without a serde default: absent -> parse error
    auto required = [&](const char *name) -> int {
      const Json *v = j.get(name);
      if (!v || v->kind != Json::Num) throw Error("can't parse " + path + ": missing field `" + name + "`");
      if (!(v->num >= 1.0 && v->num <= 2147483647.0)) throw Error("can't parse " + path + ": field `" + name + "` is out of range");
      return (int)v->num;
    };
    c.hidden = required("hidden_size");
    c.inter = required("intermediate_size");
    c.vocab = required("vocab_size");
    c.n_layers = required("num_hidden_layers");
    c.n_heads = required("num_attention_heads");
    if (!j.get("rms_norm_eps") || j.get("rms_norm_eps")->kind != Json::Num) throw Error("can't parse " + path + ": missing field `rms_norm_eps`");
    c.n_kv_heads = (int)j.number("num_key_value_heads", c.n_heads);
    const int hd_default = c.n_heads ? c.hidden / c.n_heads : 0;  // attention.rs:85
    c.head_dim = ar->head_dim ? (int)j.number("head_dim", hd_default) : hd_default;
    c.rms_eps = (float)j.number("rms_norm_eps", 1e-5);
    c.rope_theta = (float)j.number("rope_theta", ar->rope);
    c.partial_rotary = ar->phi ? (float)j.number("partial_rotary_factor", 1.0) : 1.0f;
    k.fused_qkv_proj = k.fused_gate_up_proj = ar->phi;
    c.max_seq = max_seq_override ? max_seq_override : (int)j.number("max_position_embeddings", ar->max_pos);
    c.tie_embeddings = j.boolean("tie_word_embeddings", false);
    c.qk_norm = ar->qk_norm;
    c.qkv_bias = ar->bias;
    if (ar->window) {
      const int w = (int)j.number("sliding_window", 0);  // null / absent -> 0
      c.sliding_window = (w > 0 && w < c.max_seq) ? w : 0;  // cache.rs:173-205: limit = min(window, max_seq_len)
    }
    if (k.block_kind != "llama") {
      const int n = c.n_layers;
      if (k.block_kind == "olmo2") {  // olmo2/config.rs:52-90
        c.pre_reshape_qk_norm = 1;
      } else {
        const int w = (int)j.number("sliding_window", k.block_kind == "gemma3" ? 1024 : 4096);
        k.layer_window = (w > 0 && w < c.max_seq) ? w : 0;
        k.global_layers.assign(n, false);
        if (k.block_kind == "exaone4") {  // exaone4/config.rs:62-65: every `global_layer_period`-th layer is global
          const int period = (int)j.number("global_layer_period", 4);
          for (int i = 0; i < n; i++) k.global_layers[i] = period > 0 && (i + 1) % period == 0;
        } else {                          // gemma3/config.rs:76-90: explicit schedule (true = global) or every pattern-th layer
          const Json *sched = j.get("sliding_window_attention_schedule");
          const int pattern = (int)j.number("sliding_window_pattern", 6);
          for (int i = 0; i < n; i++)
            k.global_layers[i] = (sched && sched->kind == Json::Arr && !sched->arr.empty())
                                     ? (i < (int)sched->arr.size() && sched->arr[i].kind == Json::Bool && sched->arr[i].b)
                                     : (pattern > 0 && (i + 1) % pattern == 0);
          c.tie_embeddings = 1;             // gemma3/config.rs:103
          k.residual_rms_norm = true;       // :110
          c.use_gelu_mlp = 1;               // :117
          c.embed_scale = sqrtf((float)c.hidden);  // :118
        }
      }
    }
    c.dtype = dtype;
    c.rope_factor = 1.f; c.rope_low = 1.f; c.rope_high = 4.f;
    if (const Json *rs = j.get("rope_scaling")) {
      if (rs->kind == Json::Obj) {
        const Json *t = rs->get("rope_type");
        const int orig = (int)rs->number("original_max_position_embeddings", 0);
        if (t && t->str == "llama3" && orig > 0) {  // cache.rs:49-80
          c.rope_llama3 = 1;
          c.rope_factor = (float)rs->number("factor", 0.0);            // serde defaults of the reference's RopeScaling
          c.rope_low = (float)rs->number("low_freq_factor", 0.0);
          c.rope_high = (float)rs->number("high_freq_factor", 0.0);
          c.rope_orig_max = orig;
        }
      }
    }
    if (const Json *e = j.get("eos_token_id")) {  // config.rs:6-19 EosTokenId single | array
      if (e->kind == Json::Num) k.eos.push_back((uint32_t)e->num);
      for (auto &x : e->arr) k.eos.push_back((uint32_t)x.num);
    }
    return k;
  }
  bool is_eos(uint32_t t) const {
    for (auto x : eos)
      if (x == t) return true;
    return false;
  }
  std::string layer_name(int i) const { return model_prefix + ".layers." + std::to_string(i); }  // text_model.rs:205
};

// ---------------------------------------------------------------------------------------------- D <-> f32 (load-time only)
inline float d_to_f32(uint16_t h, int dtype) {
  uint32_t u;
  if (dtype == CAKE_B200_BF16) u = (uint32_t)h << 16;
  else {
    const uint32_t s = (h >> 15) & 1u, e = (h >> 10) & 31u;
    uint32_t m = h & 1023u;
    if (e == 0) {
      if (!m) u = s << 31;
      else {  // subnormal: renormalise
        int ex = 113;
        while (!(m & 1024u)) { m <<= 1; ex--; }
        u = (s << 31) | ((uint32_t)ex << 23) | ((m & 1023u) << 13);
      }
    } else if (e == 31) u = (s << 31) | 0x7f800000u | (m << 13);
    else u = (s << 31) | ((e + 112u) << 23) | (m << 13);
  }
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint16_t f32_to_d(float f, int dtype) {  // round to nearest even, as candle's / torch's to_dtype
  uint32_t u;
  memcpy(&u, &f, 4);
  if (dtype == CAKE_B200_BF16) {
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN
    return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
  }
  const uint32_t s = (u >> 16) & 0x8000u;
  const int32_t e = (int32_t)((u >> 23) & 255u) - 127 + 15;
  uint32_t m = u & 0x7fffffu;
  if (((u >> 23) & 255u) == 255u) return (uint16_t)(s | 0x7c00u | (m ? 0x200u : 0u));
  if (e >= 31) return (uint16_t)(s | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)s;
    m |= 0x800000u;
    const int sh = 14 - e;
    uint32_t r = m >> sh;
    const uint32_t rem = m & ((1u << sh) - 1u), half = 1u << (sh - 1);
    if (rem > half || (rem == half && (r & 1u))) r++;
    return (uint16_t)(s | r);
  }
  uint32_t r = ((uint32_t)e << 10) | (m >> 13);
  const uint32_t rem = m & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) r++;  // may carry into the exponent: still the right encoding
  return (uint16_t)(s | r);
}

// ---------------------------------------------------------------------------------------------- VarBuilder
// mmapped safetensors: 8-byte little-endian header length, JSON {name: {dtype, shape, data_offsets}}, raw data
// (utils/mod.rs:255-272 single file; :333-384 model.safetensors.index.json -> the shards that hold the names).
struct TensorView {
  const void *data = nullptr;
  std::string dtype;
  std::vector<int64_t> shape;
  size_t bytes = 0;
};
class VarBuilder {
  struct Map { void *p; size_t n; };
  std::vector<Map> maps_;
  std::map<std::string, TensorView> t_;
  mutable std::map<std::string, std::vector<uint16_t>> converted_;  // tensors whose checkpoint dtype differs from the model dtype
  void add_file(const std::string &path) {
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) throw Error("can't open " + path);
    struct stat st;
    fstat(fd, &st);
    void *p = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (p == MAP_FAILED) throw Error("mmap failed for " + path);
    maps_.push_back({p, (size_t)st.st_size});
    const size_t size = (size_t)st.st_size;
    if (size < 8) throw Error(path + ": shorter than a safetensors header");
    uint64_t hl;
    memcpy(&hl, p, 8);
    if (hl > 100ull * 1024 * 1024 || 8 + hl > size) throw Error(path + ": header length " + std::to_string(hl) + " does not fit the file");
    const char *hdr = (const char *)p + 8;
    Json j = JsonParser(hdr, (size_t)hl).value();
    if (j.kind != Json::Obj) throw Error(path + ": header is not a JSON object");
    const char *base = hdr + hl;
    const size_t data_len = size - 8 - (size_t)hl;
    for (auto &kv : j.obj) {
      if (kv.first == "__metadata__") continue;
      const Json *dt = kv.second.get("dtype"), *sh = kv.second.get("shape"), *off = kv.second.get("data_offsets");
      if (!dt || dt->kind != Json::Str || !sh || sh->kind != Json::Arr || !off || off->kind != Json::Arr || off->arr.size() != 2 ||
          off->arr[0].kind != Json::Num || off->arr[1].kind != Json::Num)
        throw Error(path + ": bad entry for tensor " + kv.first);
      TensorView v;
      v.dtype = dt->str;
      double numel = 1;
      for (auto &d : sh->arr) {
        if (d.kind != Json::Num || d.num < 0) throw Error(path + ": bad shape for tensor " + kv.first);
        v.shape.push_back((int64_t)d.num);
        numel *= d.num;
      }
      const double b = off->arr[0].num, e = off->arr[1].num;
      const size_t item = (v.dtype == "F64" || v.dtype == "I64") ? 8 : (v.dtype == "F32" || v.dtype == "I32") ? 4
                          : (v.dtype == "BF16" || v.dtype == "F16" || v.dtype == "I16") ? 2 : 1;
      if (!(b >= 0 && b <= e && e <= (double)data_len) || e - b != numel * (double)item)
        throw Error(path + ": tensor " + kv.first + " has offsets that do not match its shape");
      v.data = base + (size_t)b;
      v.bytes = (size_t)(e - b);
      t_[kv.first] = v;
    }
  }
 public:
  // `only_under`: layer prefixes a worker serves — keep a shard iff it holds a tensor under one of them
  // (utils/mod.rs:334-384 load_var_builder_for_specific_layers); empty = every shard (utils/mod.rs:250-267).
  explicit VarBuilder(const std::string &dir, const std::vector<std::string> &only_under = {}) {
    const std::string idx = dir + "/model.safetensors.index.json";
    if (access(idx.c_str(), R_OK) == 0) {
      std::string txt = slurp(idx);
      Json j = JsonParser(txt.data(), txt.size()).value();
      const Json *wm = j.get("weight_map");
      if (!wm || wm->kind != Json::Obj) throw Error("no weight map in " + idx);
      std::map<std::string, bool> files;
      for (auto &kv : wm->obj) {
        if (kv.second.kind != Json::Str) continue;
        bool need = only_under.empty();
        for (auto &p : only_under)
          if (kv.first.compare(0, p.size() + 1, p + ".") == 0) need = true;
        if (need) files[kv.second.str] = true;
      }
      for (auto &f : files) add_file(dir + "/" + f.first);
    } else {
      add_file(dir + "/model.safetensors");
    }
  }
  size_t n_files() const { return maps_.size(); }
  const std::map<std::string, TensorView> &tensors() const { return t_; }
  // cake/mod.rs:335-357: the text before the first ".layers.0." key of the index wins over the configured prefix
  static std::string detect_model_prefix(const std::string &dir, const std::string &configured) {
    const std::string idx = dir + "/model.safetensors.index.json";
    if (access(idx.c_str(), R_OK) != 0) return configured;
    try {
      std::string txt = slurp(idx);
      Json j = JsonParser(txt.data(), txt.size()).value();
      if (const Json *wm = j.get("weight_map"))
        for (auto &kv : wm->obj) {
          size_t pos = kv.first.find(".layers.0.");
          if (pos != std::string::npos) return kv.first.substr(0, pos);
        }
    } catch (const std::exception &) {
    }
    return configured;
  }
  ~VarBuilder() {
    for (auto &m : maps_) munmap(m.p, m.n);
  }
  VarBuilder(const VarBuilder &) = delete;
  const TensorView *find(const std::string &name) const {
    auto it = t_.find(name);
    return it == t_.end() ? nullptr : &it->second;
  }
  // `shape`: what the loader is about to read through the pointer (candle's vb.get(shape, name) check)
  const void *get(const std::string &name, const std::string &want_dtype, const std::vector<int64_t> &shape) const {
    const TensorView *v = find(name);
    if (!v) throw Error("cannot find tensor " + name);  // candle VarBuilder error text
    if (v->shape != shape) {
      auto str = [](const std::vector<int64_t> &s) { std::string o = "["; for (size_t i = 0; i < s.size(); i++) o += (i ? ", " : "") + std::to_string(s[i]); return o + "]"; };
      throw Error("shape mismatch for " + name + ", expected: " + str(shape) + ", got: " + str(v->shape));
    }
    if (v->dtype == want_dtype) return v->data;
    // The reference converts on load (candle VarBuilder::get casts to the builder's dtype; utils/mod.rs:250-267 builds it
    // with the --dtype of the run): F32 / F16 / BF16 checkpoints are accepted for either model dtype, rounded to nearest even.
    auto known = [](const std::string &d) { return d == "BF16" || d == "F16" || d == "F32"; };
    if (!known(v->dtype) || !known(want_dtype) || want_dtype == "F32")
      throw Error("tensor " + name + " is " + v->dtype + ", expected " + want_dtype);
    auto it = converted_.find(name);
    if (it == converted_.end()) {
      size_t n = 1;
      for (auto d : v->shape) n *= (size_t)d;
      std::vector<uint16_t> out(n);
      const int to = want_dtype == "BF16" ? CAKE_B200_BF16 : CAKE_B200_F16;
      if (v->dtype == "F32") {
        const float *src = (const float *)v->data;
        for (size_t i = 0; i < n; i++) out[i] = f32_to_d(src[i], to);
      } else {
        const uint16_t *src = (const uint16_t *)v->data;
        const int from = v->dtype == "BF16" ? CAKE_B200_BF16 : CAKE_B200_F16;
        for (size_t i = 0; i < n; i++) out[i] = f32_to_d(d_to_f32(src[i], from), to);
      }
      it = converted_.emplace(name, std::move(out)).first;
    }
    return it->second.data();
  }
  const void *get_opt(const std::string &name, const std::string &want_dtype, const std::vector<int64_t> &shape) const {
    return find(name) ? get(name, want_dtype, shape) : nullptr;
  }
};

// ---------------------------------------------------------------------------------------------- Cache / Context
struct Context;
class Cache {  // cache.rs:9
 public:
  cake_b200_cache *h = nullptr;
  cake_b200_ctx *ctx;
  int batch, max_seq;
  Cache(cake_b200_ctx *c, int b, int ms) : ctx(c), batch(b), max_seq(ms) { check(cake_b200_cache_create(c, b, ms, &h), "cache"); }
  ~Cache() { cake_b200_cache_free(h); }
  Cache(const Cache &) = delete;
  void clear() { check(cake_b200_cache_clear(h), "cache.clear"); }                          // cache.rs:247-253
  std::unique_ptr<Cache> as_new() const { return std::make_unique<Cache>(ctx, batch, max_seq); }  // cache.rs:241-245
  int len(int block_idx) const { return cake_b200_cache_len(h, block_idx); }
};

struct Context {  // cake/mod.rs:41-65
  Config config;
  std::unique_ptr<VarBuilder> var_builder;
  cake_b200_ctx *h = nullptr;
  std::unique_ptr<Cache> cache;
  std::string dtype_name;
  // `worker_layers`: non-empty on a worker — only the shards that hold those layers are mapped
  Context(const std::string &model_dir, int device, int dtype, int max_seq = 0, const std::vector<std::string> &worker_layers = {})
      : config(Config::from_path(model_dir + "/config.json", dtype, max_seq)), var_builder(new VarBuilder(model_dir, worker_layers)) {
    config.model_prefix = VarBuilder::detect_model_prefix(model_dir, config.model_prefix);
    dtype_name = dtype == CAKE_B200_BF16 ? "BF16" : "F16";
    check(cake_b200_ctx_create(device, &config.c, &h), "ctx_create");
    cache.reset(new Cache(h, 1, config.c.max_seq));
  }
  ~Context() {
    cache.reset();
    if (h) cake_b200_ctx_destroy(h);
  }
  Context(const Context &) = delete;
  // config.rs:155-173 load_rms_norm_weight: with residual_rms_norm the checkpoint stores deltas and the forward weight is
  // (1 + w), added in f32 and cast back to the model dtype at load time.  Returns the tensor itself otherwise.
  const void *rms_norm_weight(const std::string &name, int64_t n) {
    const void *p = var_builder->get(name, dtype_name, {n});
    if (!config.residual_rms_norm) return p;
    norm_store.emplace_back((size_t)n);
    const uint16_t *src = (const uint16_t *)p;
    for (int64_t i = 0; i < n; i++) norm_store.back()[(size_t)i] = f32_to_d(d_to_f32(src[i], config.c.dtype) + 1.0f, config.c.dtype);
    return norm_store.back().data();
  }
 private:
  std::vector<std::vector<uint16_t>> norm_store;  // the library copies at load; kept until the ctx goes for simplicity
};

// ---------------------------------------------------------------------------------------------- Forwarder
class Forwarder {  // cake/mod.rs:510-556
 public:
  virtual ~Forwarder() = default;
  // x, y: host buffers (batch, seq, hidden) in the model dtype; the _dev variants take device pointers
  virtual void forward(const void *x_host, void *y_host, int batch, int seq, size_t index_pos, size_t block_idx, Context &ctx) = 0;
  virtual void forward_mut(const void *x_host, void *y_host, int batch, int seq, size_t index_pos, size_t block_idx, Context &ctx) {
    forward(x_host, y_host, batch, seq, index_pos, block_idx, ctx);
  }
  virtual void goodbye() {}
  virtual const std::string &layer_name() const = 0;
  virtual std::string ident() const { return "local"; }
  virtual cake_b200_block *handle() const { return nullptr; }
};

class Transformer : public Forwarder {  // transformer.rs:14-150, backed by the CUDA block
  std::string name_;
  cake_b200_block *h_ = nullptr;
 public:
  static std::unique_ptr<Transformer> load(const std::string &name, Context &ctx) {  // transformer.rs:79-101
    const VarBuilder &vb = *ctx.var_builder;
    const std::string &dt = ctx.dtype_name;
    const auto &c = ctx.config.c;
    const int64_t H = c.hidden, I = c.inter, sq = (int64_t)c.n_heads * c.head_dim, skv = (int64_t)c.n_kv_heads * c.head_dim, hd = c.head_dim;
    typedef std::vector<int64_t> Sh;
    auto g = [&](const char *s, const Sh &shape) { return vb.get(name + "." + s, dt, shape); };
    int layer = std::stoi(name.substr(name.rfind('.') + 1));
    auto t = std::unique_ptr<Transformer>(new Transformer());
    t->name_ = name;
    const size_t es = 2;  // bf16 / f16
    const void *q, *k, *v, *gate, *up;
    if (ctx.config.fused_qkv_proj) {  // one tensor = cat(q, k, v): the three are contiguous row ranges of it
      const char *w = (const char *)g("self_attn.qkv_proj.weight", Sh{sq + 2 * skv, H});
      q = w; k = w + (size_t)sq * H * es; v = w + (size_t)(sq + skv) * H * es;
    } else {
      q = g("self_attn.q_proj.weight", Sh{sq, H}); k = g("self_attn.k_proj.weight", Sh{skv, H}); v = g("self_attn.v_proj.weight", Sh{skv, H});
    }
    if (ctx.config.fused_gate_up_proj) {
      const char *w = (const char *)g("mlp.gate_up_proj.weight", Sh{2 * I, H});
      gate = w; up = w + (size_t)I * H * es;
    } else {
      gate = g("mlp.gate_proj.weight", Sh{I, H}); up = g("mlp.up_proj.weight", Sh{I, H});
    }
    // norm vectors by ROLE: ln1 = pre-attention, ln2 = pre-MLP, post_attn / post_ffn = the sandwich norms.  Checkpoint names:
    //   llama / exaone4 (transformer.rs:84-90, exaone4/block.rs:74-77): input_layernorm | post_attention_layernorm (= pre-MLP)
    //   gemma3 (gemma3/block.rs:84-91): input_layernorm | post_attention_layernorm (POST-attention) | pre_feedforward_layernorm
    //                                   | post_feedforward_layernorm
    //   olmo2 (olmo2/block.rs:49-52): post_attention_layernorm, post_feedforward_layernorm only
    const Config::Variant var = ctx.config.layer_variant(layer);
    const std::string &kind = ctx.config.block_kind;
    auto nw = [&](const char *s) { return ctx.rms_norm_weight(name + "." + s, H); };
    const void *ln1 = nullptr, *ln2 = nullptr, *post_attn = nullptr, *post_ffn = nullptr;
    if (kind == "olmo2") {
      post_attn = nw("post_attention_layernorm.weight"); post_ffn = nw("post_feedforward_layernorm.weight");
    } else if (kind == "gemma3") {
      ln1 = nw("input_layernorm.weight"); post_attn = nw("post_attention_layernorm.weight");
      ln2 = nw("pre_feedforward_layernorm.weight"); post_ffn = nw("post_feedforward_layernorm.weight");
    } else {
      ln1 = nw("input_layernorm.weight"); ln2 = nw("post_attention_layernorm.weight");
    }
    const int64_t qn_dim = c.pre_reshape_qk_norm ? sq : hd, kn_dim = c.pre_reshape_qk_norm ? skv : hd;  // attention.rs:121-122
    check(cake_b200_block_load(ctx.h, layer, q, k, v, g("self_attn.o_proj.weight", Sh{H, sq}),
                               gate, up, g("mlp.down_proj.weight", Sh{H, I}), ln1, ln2,
                               c.qkv_bias ? g("self_attn.q_proj.bias", Sh{sq}) : nullptr, c.qkv_bias ? g("self_attn.k_proj.bias", Sh{skv}) : nullptr,
                               c.qkv_bias ? g("self_attn.v_proj.bias", Sh{skv}) : nullptr,
                               c.qk_norm ? ctx.rms_norm_weight(name + ".self_attn.q_norm.weight", qn_dim) : nullptr,   // attention.rs:120-129
                               c.qk_norm ? ctx.rms_norm_weight(name + ".self_attn.k_norm.weight", kn_dim) : nullptr, &t->h_),
          name);
    if (var.post_norms || var.window >= 0 || var.no_rope) {  // load_custom(vb, cfg, use_qk_norm, sliding_window, use_rope)
      cake_b200_block_variant bv{var.window, var.no_rope ? 0 : 1, post_attn, post_ffn};
      check(cake_b200_block_set_variant(t->h_, &bv), name);
    }
    return t;
  }
  ~Transformer() override { cake_b200_block_free(h_); }
  void forward(const void *x, void *y, int batch, int seq, size_t index_pos, size_t block_idx, Context &ctx) override {
    if (!ctx.cache) throw Error("No cache specified");  // transformer.rs:120
    cake_b200_block *blocks[1] = {h_};
    int idx[1] = {(int)block_idx};
    check(cake_b200_forward_batch_host(ctx.h, blocks, idx, 1, ctx.cache->h, x, y, batch, seq, (int)index_pos), "attention/mlp " + name_);
  }
  const std::string &layer_name() const override { return name_; }
  cake_b200_block *handle() const override { return h_; }
};

// ---------------------------------------------------------------------------------------------- TextModelBase / Master
struct Token {
  uint32_t id;
  bool is_end_of_stream;
};

class TextModelBase {  // text_model.rs:133-530 for token-id prompts
 public:
  Context &ctx;
  std::vector<std::unique_ptr<Forwarder>> blocks;
  std::vector<uint32_t> tokens;
  size_t index_pos = 0, generated = 0, prompt_len = 0;
  float repeat_penalty = 1.0f;
  size_t repeat_last_n = 128;
  bool graph_ready = false;

  explicit TextModelBase(Context &c) : ctx(c) {}
  static std::unique_ptr<TextModelBase> load(Context &ctx) {  // text_model.rs:150-264 (all layers local)
    auto m = std::make_unique<TextModelBase>(ctx);
    const VarBuilder &vb = *ctx.var_builder;
    const std::string p = ctx.config.model_prefix, &dt = ctx.dtype_name;
    const int64_t V = ctx.config.c.vocab, H = ctx.config.c.hidden;
    check(cake_b200_head_load(ctx.h, vb.get(p + ".embed_tokens.weight", dt, {V, H}), ctx.rms_norm_weight(p + ".norm.weight", H),
                              ctx.config.c.tie_embeddings ? nullptr : vb.get("lm_head.weight", dt, {V, H})),
          "head_load");
    for (int i = 0; i < ctx.config.c.n_layers; i++) m->blocks.push_back(Transformer::load(ctx.config.layer_name(i), ctx));
    return m;
  }
  std::vector<cake_b200_block *> handles() const {
    std::vector<cake_b200_block *> v;
    for (auto &b : blocks) v.push_back(b->handle());
    return v;
  }
  void prepare_prompt(const std::vector<uint32_t> &ids) {  // text_model.rs:371-395
    tokens = ids;
    ctx.cache->clear();
    index_pos = 0;
    prompt_len = ids.size();
    graph_ready = false;
  }
  // text_model.rs:397-495.  index 0: whole prompt at position 0 through the block walk (prefill kernels);
  // index > 0: the greedy path replays the decode graph (one persistent kernel), other settings go through logits.
  Token next_token(size_t index) {
    const auto &c = ctx.config.c;
    uint32_t next = 0;
    if (index == 0 || repeat_penalty != 1.0f || !ctx.config.standard_blocks()) {  // the decode graph covers the standard block only
      std::vector<uint32_t> in = (index > 0) ? std::vector<uint32_t>{tokens.back()} : tokens;
      const size_t pos = (index > 0) ? index_pos : 0;
      if (in.empty()) throw Error("empty prompt");
      (void)c;
      // embed -> contiguous local blocks -> ln_f / lm_head, all on the device (text_model.rs:266-352)
      check(cake_b200_embed(ctx.h, in.data(), 1, (int)in.size(), scratch(in.size())), "embedding");
      auto hs = handles();
      std::vector<int> idx(hs.size());
      for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
      check(cake_b200_forward_batch(ctx.h, hs.data(), idx.data(), (int)hs.size(), ctx.cache->h, scratch(in.size()), scratch(in.size()), 1,
                                    (int)in.size(), (int)pos),
            "forward");
      index_pos += in.size();
      if (repeat_penalty == 1.0f) {
        check(cake_b200_logits(ctx.h, scratch(in.size()), 1, (int)in.size(), nullptr, &next), "lm_head");
      } else {
        check(cake_b200_logits(ctx.h, scratch(in.size()), 1, (int)in.size(), logits_dev(), nullptr), "lm_head");
        const size_t ngen = tokens.size() - prompt_len, start = ngen > repeat_last_n ? ngen - repeat_last_n : 0;  // :435-452
        check(cake_b200_repeat_penalty_argmax(ctx.h, logits_dev(), repeat_penalty, tokens.data() + prompt_len + start, (int)(ngen - start), &next),
              "sample");
      }
    } else {
      if (!graph_ready) {
        auto hs = handles();
        std::vector<int> idx(hs.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
        check(cake_b200_decode_build(ctx.h, hs.data(), idx.data(), (int)hs.size(), ctx.cache->h, 0, 1), "decode_build");
        check(cake_b200_decode_begin(ctx.h, tokens.back(), (int)index_pos), "decode_begin");
        graph_ready = true;
      }
      check(cake_b200_decode_step_host(ctx.h, tokens.back(), &next), "decode_step");
      index_pos += 1;
    }
    generated++;
    tokens.push_back(next);
    return Token{next, ctx.config.is_eos(next)};
  }
  void reset() {  // text_model.rs:497-515
    tokens.clear();
    ctx.cache->clear();
    index_pos = generated = prompt_len = 0;
    graph_ready = false;
  }
  void goodbye() {
    for (auto &b : blocks) b->goodbye();
  }
  ~TextModelBase() { release_scratch(); }

 private:
  // device scratch for (seq, hidden) activations and (vocab) logits
  void *scratch_ = nullptr, *logits_ = nullptr;
  size_t scratch_rows_ = 0;
  void *scratch(size_t rows) {
    if (rows > scratch_rows_) {
      check(cake_b200_sync(ctx.h), "sync");
      if (scratch_) cake_b200_dev_free(ctx.h, scratch_);
      check(cake_b200_dev_alloc(ctx.h, rows * (size_t)ctx.config.c.hidden * 2, &scratch_), "dev_alloc");
      scratch_rows_ = rows;
    }
    return scratch_;
  }
  void *logits_dev() {
    if (!logits_) check(cake_b200_dev_alloc(ctx.h, (size_t)ctx.config.c.vocab * 2, &logits_), "dev_alloc");
    return logits_;
  }
  void release_scratch() {
    if (scratch_) cake_b200_dev_free(ctx.h, scratch_);
    if (logits_) cake_b200_dev_free(ctx.h, logits_);
  }
};

class Master {  // sharding/master.rs:14-191
 public:
  TextModelBase &model;
  explicit Master(TextModelBase &m) : model(m) {}
  struct Result {
    std::vector<uint32_t> tokens;
    double tok_s;
  };
  // master.rs:109-168: tok/s = (generated - 1) / time since the first generated token
  Result generate_text(const std::vector<uint32_t> &prompt, size_t sample_len, const std::function<void(const Token &)> &stream = nullptr) {
    model.prepare_prompt(prompt);
    Result r;
    auto start = std::chrono::steady_clock::now();
    for (size_t index = 0; index < sample_len; index++) {
      if (index == 1) start = std::chrono::steady_clock::now();
      Token t = model.next_token(index);
      if (t.is_end_of_stream) break;
      r.tokens.push_back(t.id);
      if (stream) stream(t);
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    r.tok_s = model.generated > 1 ? (double)(model.generated - 1) / dt : 0.0;
    return r;
  }
};

}  // namespace cake_host
