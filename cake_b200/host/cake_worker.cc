// cake_worker — a cake worker endpoint on a B200 (cake's `cake run --worker`, worker.rs:79-597, for the block-forward
// path): listens for an unmodified cake master on cake's TCP protocol (cake_wire.hpp) and runs the requested layer
// range through libcake_b200.so (cake_host.hpp).
//   cake_worker <model_dir> --layers model.layers.16-31 [--address 0.0.0.0:10128] [--cluster-key K]
//               [--dtype bf16|f16] [--max-seq S] [--device 0] [--connections N]
//   cake_worker <model_dir> --topology topology.yml --name worker1 ...   (layers from the master's topology file)
//   cake_worker --topology topology.yml --name worker1                    (dry run: print the expanded layer list)
//   cake_worker <model_dir> [--layers ...] --list-tensors                 (dry run: tensors the node's VarBuilder maps)
//   cake_worker --echo [--reflect] [--address 127.0.0.1:0] [--cluster-key K] [--connections N]    (no GPU: protocol only)
// --layers takes the topology file's syntax (topology.rs:13,143-168): names or inclusive ranges, comma separated.
// Prints "listening on <host>:<port>" once the socket is bound.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>

#include "cake_host.hpp"
#include "cake_wire.hpp"

using namespace cake_host;
namespace cw = cake_wire;

// topology.rs:143-168: "model.layers.0-5" -> model.layers.0 .. model.layers.5 (inclusive); regex ^(.+[^\d])(\d+)-(\d+)$
static std::vector<std::string> expand_layers(const std::string &spec) {
  std::vector<std::string> out;
  std::stringstream ss(spec);
  std::string item;
  while (std::getline(ss, item, ',')) {
    if (item.empty()) continue;
    size_t dash = item.rfind('-');
    bool range = dash != std::string::npos && dash + 1 < item.size();
    size_t b = dash;
    if (range) {
      for (size_t i = dash + 1; i < item.size(); i++) range = range && isdigit((unsigned char)item[i]);
      while (b > 0 && isdigit((unsigned char)item[b - 1])) b--;
      range = range && b < dash && b > 0;  // digits before the dash, and a non-digit base in front of them
    }
    if (!range) { out.push_back(item); continue; }
    const std::string base = item.substr(0, b);
    const long start = std::stol(item.substr(b, dash - b)), stop = std::stol(item.substr(dash + 1));
    if (stop < start) throw Error("invalid range expression " + item + ", end must be >= start");
    for (long n = start; n <= stop; n++) out.push_back(base + std::to_string(n));
  }
  return out;
}

// The subset of YAML a cake topology file uses (topology.rs:15-40,126-168): a map  worker-name -> { host: str,
// description: str, layers: [str, ...] (block list "- item" or inline "[a, b]"), other scalars ignored }.
struct TopologyNode {
  std::string host;
  std::vector<std::string> layers;  // as written (ranges not expanded)
};
static std::string yaml_scalar(std::string s) {
  size_t hash = std::string::npos;
  bool in_s = false, in_d = false;
  for (size_t i = 0; i < s.size(); i++) {
    if (s[i] == '\'' && !in_d) in_s = !in_s;
    else if (s[i] == '"' && !in_s) in_d = !in_d;
    else if (s[i] == '#' && !in_s && !in_d && (i == 0 || isspace((unsigned char)s[i - 1]))) { hash = i; break; }
  }
  if (hash != std::string::npos) s = s.substr(0, hash);
  size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
  if (a == std::string::npos) return "";
  s = s.substr(a, b - a + 1);
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) s = s.substr(1, s.size() - 2);
  return s;
}
static std::map<std::string, TopologyNode> parse_topology(const std::string &text) {
  std::map<std::string, TopologyNode> topo;
  std::stringstream ss(text);
  std::string line, cur, key;
  while (std::getline(ss, line)) {
    const size_t ind = line.find_first_not_of(" \t");
    if (ind == std::string::npos || line[ind] == '#' || line.compare(ind, 3, "---") == 0) continue;
    std::string body = line.substr(ind);
    if (ind == 0) {  // "name:"
      const size_t c = body.find(':');
      if (c == std::string::npos) throw Error("topology: expected 'name:' at '" + body + "'");
      cur = yaml_scalar(body.substr(0, c));
      topo[cur];
      key.clear();
      continue;
    }
    if (cur.empty()) throw Error("topology: indented line before any worker name");
    if (body[0] == '-') {  // list item of the last key
      if (key == "layers") topo[cur].layers.push_back(yaml_scalar(body.substr(1)));
      continue;
    }
    const size_t c = body.find(':');
    if (c == std::string::npos) continue;
    key = yaml_scalar(body.substr(0, c));
    std::string val = yaml_scalar(body.substr(c + 1));
    if (key == "host") topo[cur].host = val;
    else if (key == "layers" && !val.empty() && val.front() == '[') {  // inline list
      std::stringstream ls(val.substr(1, val.rfind(']') == std::string::npos ? std::string::npos : val.rfind(']') - 1));
      std::string item;
      while (std::getline(ls, item, ',')) {
        item = yaml_scalar(item);
        if (!item.empty()) topo[cur].layers.push_back(item);
      }
    }
  }
  return topo;
}
static std::string join(const std::vector<std::string> &v) {
  std::string s;
  for (auto &x : v) s += (s.empty() ? "" : ",") + x;
  return s;
}

// The worker's blocks behind the wire: host buffers in, cake_b200_forward_batch_host, host buffers out.  Every
// connection has its own KV cache (K/V pages are allocated per layer on first use, so idle sessions cost nothing);
// forwards share the ctx's stream and scratch and are serialised by a mutex.
struct B200Backend : cw::Backend {
  Context &ctx;
  std::map<std::string, std::unique_ptr<Transformer>> blocks;
  int device_;
  std::mutex mu;
  B200Backend(Context &c, const std::vector<std::string> &names, int device) : ctx(c), device_(device) {
    for (auto &n : names) blocks[n] = Transformer::load(n, ctx);
  }
  std::string dtype() const override { return ctx.dtype_name; }
  std::string device() const override { return "cuda"; }
  uint64_t device_idx() const override { return (uint64_t)device_; }

  struct S : cw::Session {
    B200Backend &be;
    std::unique_ptr<Cache> cache;
    explicit S(B200Backend &b) : be(b) {
      std::lock_guard<std::mutex> g(be.mu);
      cache = be.ctx.cache->as_new();  // worker.rs:60-75
    }
    ~S() override {
      std::lock_guard<std::mutex> g(be.mu);
      cake_b200_sync(be.ctx.h);
      cache.reset();
    }
    void clear_cache() override {
      std::lock_guard<std::mutex> g(be.mu);
      cache->clear();
    }
    cw::RawTensor forward_ops(const cw::RawTensor &x, const std::vector<cw::Op> &ops) override {
      Context &ctx = be.ctx;
      for (auto &o : ops)
        if (!be.blocks.count(std::get<0>(o))) throw Error("could not find layer " + std::get<0>(o));  // worker.rs:513
      const uint8_t want = ctx.dtype_name == "BF16" ? cw::BF16 : cw::F16;
      auto where = [&](size_t i) { return "forward pass failed for layer " + std::get<0>(ops[i]) + " (block_idx=" + std::to_string(std::get<2>(ops[i])) + "): "; };
      if (x.dtype != want) throw Error(where(0) + "activation dtype tag " + std::to_string((int)x.dtype) + " is not the model dtype " + ctx.dtype_name);
      if (x.shape.size() != 3 || x.shape[2] != (uint64_t)ctx.config.c.hidden) throw Error(where(0) + "unexpected activation shape");
      // u64 fields from an untrusted master are narrowed to int below: bound them first (a [0, 2^32+5, H] shape has numel 0
      // and passes RawTensor::validate)
      const uint64_t lim = (uint64_t)ctx.config.c.max_seq;
      if (x.shape[0] < 1 || x.shape[0] > 1024 || x.shape[1] < 1 || x.shape[1] > lim) throw Error(where(0) + "activation batch / sequence out of range");
      for (size_t k = 0; k < ops.size(); k++)
        if (std::get<1>(ops[k]) > lim || std::get<2>(ops[k]) > (uint64_t)ctx.config.c.n_layers)
          throw Error(where(k) + "index_pos / block_idx out of range");
      cw::RawTensor cur = x, next = x;
      std::lock_guard<std::mutex> g(be.mu);
      size_t i = 0;
      while (i < ops.size()) {  // consecutive ops that share index_pos go down in one call (text_model.rs:298-321)
        size_t j = i;
        std::vector<cake_b200_block *> hs;
        std::vector<int> idx;
        while (j < ops.size() && std::get<1>(ops[j]) == std::get<1>(ops[i])) {
          hs.push_back(be.blocks[std::get<0>(ops[j])]->handle());
          idx.push_back((int)std::get<2>(ops[j]));
          j++;
        }
        int rc = cake_b200_forward_batch_host(ctx.h, hs.data(), idx.data(), (int)hs.size(), cache->h, cur.data.data(), next.data.data(),
                                              (int)x.shape[0], (int)x.shape[1], (int)std::get<1>(ops[i]));
        if (rc != 0) throw Error(where(i) + cake_b200_last_error());
        std::swap(cur, next);
        i = j;
      }
      return cur;
    }
  };
  std::unique_ptr<cw::Session> new_session() override { return std::unique_ptr<cw::Session>(new S(*this)); }
};

int main(int argc, char **argv) {
  std::string dir, layers, address = "127.0.0.1:10128", key, topology, name;
  bool echo = false, reflect = false, has_key = false, list_tensors = false;
  int dtype = CAKE_B200_BF16, max_seq = 0, device = 0, connections = -1;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> std::string { return i + 1 < argc ? std::string(argv[++i]) : std::string(); };
    if (a == "--layers") layers = next();
    else if (a == "--address") address = next();
    else if (a == "--cluster-key") { key = next(); has_key = true; }
    else if (a == "--dtype") dtype = (next() == "f16") ? CAKE_B200_F16 : CAKE_B200_BF16;
    else if (a == "--max-seq") max_seq = std::stoi(next());
    else if (a == "--device") device = std::stoi(next());
    else if (a == "--connections") connections = std::stoi(next());
    else if (a == "--topology") topology = next();
    else if (a == "--name") name = next();
    else if (a == "--echo") echo = true;
    else if (a == "--reflect") reflect = true;
    else if (a == "--list-tensors") list_tensors = true;
    else if (a == "--expand") {  // print the expansion of a --layers expression and exit (no GPU)
      try {
        for (auto &n : expand_layers(next())) printf("%s\n", n.c_str());
      } catch (const std::exception &e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
      }
      return 0;
    } else if (dir.empty() && a[0] != '-') dir = a;
  }
  if (!topology.empty()) {  // `cake run --mode worker --name N --topology T` picks the node's layers (worker.rs:150-154)
    try {
      auto topo = parse_topology(slurp(topology));
      if (!topo.count(name)) throw Error("could not find topology node for worker '" + name + "'");
      if (layers.empty()) layers = join(topo[name].layers);
      if (dir.empty() && !echo) {  // --expand-topology style dry run: print what this worker would serve
        for (auto &n : expand_layers(layers)) printf("%s\n", n.c_str());
        return 0;
      }
    } catch (const std::exception &e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
  }
  if (list_tensors && !dir.empty()) {  // dry run (no GPU): what the mmapped VarBuilder of this node holds
    try {
      const std::vector<std::string> names = layers.empty() ? std::vector<std::string>{} : expand_layers(layers);
      VarBuilder vb(dir, names);
      printf("files %zu prefix %s\n", vb.n_files(), VarBuilder::detect_model_prefix(dir, "model").c_str());
      for (auto &kv : vb.tensors()) {
        uint64_t hsh = 1469598103934665603ull;  // FNV-1a over the tensor bytes
        const unsigned char *p = (const unsigned char *)kv.second.data;
        for (size_t i = 0; i < kv.second.bytes; i++) { hsh ^= p[i]; hsh *= 1099511628211ull; }
        std::string shape;
        for (auto d : kv.second.shape) shape += (shape.empty() ? "" : "x") + std::to_string(d);
        printf("%s %s [%s] %zu %016llx\n", kv.first.c_str(), kv.second.dtype.c_str(), shape.c_str(), kv.second.bytes, (unsigned long long)hsh);
      }
    } catch (const std::exception &e) {
      fprintf(stderr, "error: %s\n", e.what());
      return 1;
    }
    return 0;
  }
  if (!echo && (dir.empty() || layers.empty())) {
    fprintf(stderr, "usage: %s <model_dir> --layers model.layers.A-B [--address host:port] [--cluster-key K] [--dtype bf16|f16] "
                    "[--max-seq S] [--device N] [--connections N]\n       %s --echo [--reflect] [--address host:port] [--cluster-key K]\n", argv[0], argv[0]);
    return 2;
  }
  try {
    const size_t colon = address.rfind(':');
    if (colon == std::string::npos) throw Error("--address needs host:port");
    const std::string host = address.substr(0, colon);
    const int port = std::stoi(address.substr(colon + 1));
    std::unique_ptr<Context> ctx;
    std::unique_ptr<cw::Backend> be;
    if (echo) {
      be.reset(new cw::EchoBackend());
    } else {
      const std::vector<std::string> names = expand_layers(layers);
      ctx.reset(new Context(dir, device, dtype, max_seq, names));
      be.reset(new B200Backend(*ctx, names, device));
    }
    cw::WireWorker w(*be, host, port, has_key ? &key : nullptr);
    w.reflect = reflect;
    printf("listening on %s\n", w.address.c_str());
    fflush(stdout);
    w.serve(connections);
    be.reset();
    ctx.reset();
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
