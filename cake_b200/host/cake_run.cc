// cake_run — minimal `cake run` for the block-forward path on a B200, built on cake_host.hpp (the C++ mirror of
// cake's interface) and libcake_b200.so.  It takes token ids instead of text (tokenizer / chat template are outside
// the path):   cake_run <model_dir> --prompt-ids 1,2,3 [-n 32] [--dtype bf16|f16] [--repeat-penalty 1.0]
// <model_dir> holds config.json + model.safetensors (or model.safetensors.index.json + shards), HF layout.
// Prints one line:  tokens: t0 t1 ...   and   tok/s as the reference defines it (master.rs:160-166).
// --show-config: parse config.json only and print the resolved block configuration (needs no GPU).
#include <cstdio>
#include <cstdlib>

#include "cake_host.hpp"

using namespace cake_host;

int main(int argc, char **argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <model_dir> --prompt-ids a,b,c [-n N] [--dtype bf16|f16] [--repeat-penalty P] [--max-seq S]\n", argv[0]);
    return 2;
  }
  std::string dir = argv[1];
  std::vector<uint32_t> prompt;
  size_t n = 32;
  int dtype = CAKE_B200_BF16, max_seq = 0;
  float rp = 1.0f;
  bool show_config = false;
  for (int i = 2; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> std::string { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
    if (a == "--prompt-ids") {
      std::stringstream ss(next());
      std::string t;
      while (std::getline(ss, t, ',')) prompt.push_back((uint32_t)std::stoul(t));
    } else if (a == "-n") n = std::stoul(next());
    else if (a == "--dtype") dtype = (next() == "f16") ? CAKE_B200_F16 : CAKE_B200_BF16;
    else if (a == "--repeat-penalty") rp = std::stof(next());
    else if (a == "--max-seq") max_seq = std::stoi(next());
    else if (a == "--show-config") show_config = true;
  }
  try {
    if (show_config) {
      Config k = Config::from_path(dir + "/config.json", dtype, max_seq);
      const auto &c = k.c;
      printf("arch=%s hidden=%d inter=%d heads=%d kv_heads=%d head_dim=%d layers=%d vocab=%d max_seq=%d rms_eps=%g "
             "rope_theta=%g qkv_bias=%d qk_norm=%d tie=%d rope_llama3=%d n_eos=%zu partial_rotary=%g fused=%d sliding_window=%d "
             "block_kind=%s pre_reshape_qk_norm=%d gelu=%d embed_scale=%g residual_rms_norm=%d layer_window=%d globals=%s\n",
             k.arch.c_str(), c.hidden, c.inter, c.n_heads, c.n_kv_heads, c.head_dim, c.n_layers, c.vocab, c.max_seq,
             (double)c.rms_eps, (double)c.rope_theta, c.qkv_bias, c.qk_norm, c.tie_embeddings, c.rope_llama3, k.eos.size(),
             (double)c.partial_rotary, (int)(k.fused_qkv_proj && k.fused_gate_up_proj), c.sliding_window,
             k.block_kind.c_str(), c.pre_reshape_qk_norm, c.use_gelu_mlp, (double)c.embed_scale, (int)k.residual_rms_norm, k.layer_window,
             [&]() { std::string g; for (bool b : k.global_layers) g += b ? '1' : '0'; return g.empty() ? std::string("-") : g; }().c_str());
      return 0;
    }
    Context ctx(dir, 0, dtype, max_seq);
    auto model = TextModelBase::load(ctx);
    model->repeat_penalty = rp;
    Master master(*model);
    auto r = master.generate_text(prompt, n);
    printf("tokens:");
    for (auto t : r.tokens) printf(" %u", t);
    printf("\ntok/s: %.2f\n", r.tok_s);
    model.reset();
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
