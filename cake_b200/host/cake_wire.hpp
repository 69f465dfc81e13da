// cake_wire.hpp — cake's TCP wire protocol in C++ (header-only, no dependencies), so that a B200 box can be one worker
// of an unmodified remote cake master.  Same scope and reference anchors as cake_b200/wire.py:
//
//   proto/mod.rs:3-10          PROTO_MAGIC, MESSAGE_MAX_SIZE
//   proto/message.rs:7-48      dtype tags, RawTensor { data, dtype, shape }
//   proto/message.rs:171-247   WorkerInfo, enum Message (declaration order = wire tag)
//   proto/message.rs:334-394   framing: magic u32 | payload length u32 (big endian) | payload
//   auth.rs:1-118              mutual HMAC-SHA256 challenge-response before any framing
//   worker.rs:298-575          one master connection: Hello -> WorkerInfo; SingleOp / Batch -> Tensor | WorkerError;
//                              Goodbye -> cache clear + WorkerInfo; LayerAssignment as first message -> Ack + WorkerReady
//
// Payloads use the `speedy` crate's BigEndian encoding (restated from its published format; the reference holds no
// golden bytes, so byte-level parity with a real cake peer is unpinned): integers big endian, usize as u64, u128 as
// 16 bytes, bool as u8, String / Vec<T> as u32 count + elements, fields in declaration order, enum tag u32.
#pragma once
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/utsname.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <random>
#include <stdexcept>
#include <string>
#include <tuple>
#include <list>
#include <memory>
#include <vector>

namespace cake_wire {

constexpr uint32_t PROTO_MAGIC = 0x0104F4C7u;
constexpr uint32_t MESSAGE_MAX_SIZE = 512u * 1024u * 1024u;

struct ProtocolError : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct ConnectionClosed : std::runtime_error {
  using std::runtime_error::runtime_error;
};

// ------------------------------------------------------------------------------------------ dtype tags (message.rs:7-36)
enum DTypeTag : uint8_t { U8 = 0, U32 = 1, I64 = 2, BF16 = 3, F16 = 4, F32 = 5, F64 = 6, F8E4M3 = 7 };
inline size_t dtype_size(uint8_t tag) {
  switch (tag) {
    case U8: case F8E4M3: return 1;
    case BF16: case F16: return 2;
    case U32: case F32: return 4;
    case I64: case F64: return 8;
    default: throw ProtocolError("unknown dtype tag: " + std::to_string((int)tag));
  }
}

// ------------------------------------------------------------------------------------------ speedy (BigEndian)
class Writer {
 public:
  std::vector<uint8_t> buf;
  void u8(uint8_t v) { buf.push_back(v); }
  void u32(uint32_t v) { for (int s = 24; s >= 0; s -= 8) buf.push_back((uint8_t)(v >> s)); }
  void u64(uint64_t v) { for (int s = 56; s >= 0; s -= 8) buf.push_back((uint8_t)(v >> s)); }
  void u128(uint64_t hi, uint64_t lo) { u64(hi); u64(lo); }
  void boolean(bool v) { u8(v ? 1 : 0); }
  void blob(const uint8_t *p, size_t n) {
    if (n > 0xffffffffull) throw ProtocolError("length does not fit u32");
    u32((uint32_t)n);
    buf.insert(buf.end(), p, p + n);
  }
  void string(const std::string &s) { blob((const uint8_t *)s.data(), s.size()); }
};

class Reader {
  const uint8_t *p_;
  size_t n_, pos_ = 0;
  const uint8_t *take(size_t n) {
    if (n > n_ - pos_) throw ProtocolError("truncated message: need " + std::to_string(n) + " bytes at offset " + std::to_string(pos_));
    const uint8_t *r = p_ + pos_;
    pos_ += n;
    return r;
  }
 public:
  Reader(const uint8_t *p, size_t n) : p_(p), n_(n) {}
  uint8_t u8() { return *take(1); }
  uint32_t u32() { const uint8_t *b = take(4); return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | b[3]; }
  uint64_t u64() { uint64_t hi = u32(); return (hi << 32) | u32(); }
  bool boolean() { return u8() != 0; }
  std::vector<uint8_t> blob() { uint32_t n = u32(); const uint8_t *b = take(n); return std::vector<uint8_t>(b, b + n); }
  std::string string() { uint32_t n = u32(); const uint8_t *b = take(n); return std::string((const char *)b, n); }
  void done() const { if (pos_ != n_) throw ProtocolError(std::to_string(n_ - pos_) + " trailing bytes after message"); }
};

// ------------------------------------------------------------------------------------------ RawTensor / WorkerInfo / Message
struct RawTensor {  // message.rs:39-48; data = elements in native (little-endian) order, row-major
  std::vector<uint8_t> data;
  uint8_t dtype = F16;
  std::vector<uint64_t> shape;
  uint64_t numel() const {  // saturating: a hostile shape must not wrap around to a plausible byte count
    uint64_t n = 1;
    for (auto d : shape) {
      if (d != 0 && n > (UINT64_MAX >> 8) / d) return UINT64_MAX >> 8;
      n *= d;
    }
    return n;
  }
  void validate() const {
    if (data.size() != numel() * dtype_size(dtype)) throw ProtocolError("tensor shape and dtype do not match " + std::to_string(data.size()) + " bytes");
  }
  void write(Writer &w) const {
    w.blob(data.data(), data.size());
    w.u8(dtype);
    w.u32((uint32_t)shape.size());
    for (auto d : shape) w.u64(d);
  }
  static RawTensor read(Reader &r) {
    RawTensor t;
    t.data = r.blob();
    t.dtype = r.u8();
    uint32_t n = r.u32();
    for (uint32_t i = 0; i < n; i++) t.shape.push_back(r.u64());
    return t;
  }
};

struct WorkerInfo {  // message.rs:171-188
  std::string version, dtype, os, arch, device;
  uint64_t device_idx = 0;
  uint64_t latency_hi = 0, latency_lo = 0;  // u128 milliseconds
  void write(Writer &w) const {
    w.string(version); w.string(dtype); w.string(os); w.string(arch); w.string(device);
    w.u64(device_idx);
    w.u128(latency_hi, latency_lo);
  }
  static WorkerInfo read(Reader &r) {
    WorkerInfo i;
    i.version = r.string(); i.dtype = r.string(); i.os = r.string(); i.arch = r.string(); i.device = r.string();
    i.device_idx = r.u64();
    i.latency_hi = r.u64(); i.latency_lo = r.u64();
    return i;
  }
};

typedef std::tuple<std::string, uint64_t, uint64_t> Op;  // (layer_name, index_pos, block_idx)

struct Message {  // message.rs:190-247
  enum Kind : uint32_t { Hello, WorkerInfoMsg, SingleOp, Batch, Tensor, Goodbye, LayerAssignment, LayerAssignmentAck,
                         ModelDataChunk, ModelDataDone, ModelDataResume, WorkerReady, WorkerError, N_KINDS };
  Kind kind = Hello;
  WorkerInfo info;                  // WorkerInfo
  RawTensor x;                      // SingleOp, Batch, Tensor
  std::string layer_name;           // SingleOp
  uint64_t index_pos = 0, block_idx = 0;
  std::vector<Op> batch;            // Batch
  std::vector<std::string> layers;  // LayerAssignment
  std::string model_hash;
  bool needs_data = false;          // LayerAssignmentAck
  std::string filename;             // ModelDataChunk, ModelDataResume
  uint64_t offset = 0, total_size = 0;
  bool compressed = false;
  uint32_t checksum = 0;
  std::vector<uint8_t> data;
  std::string message;              // WorkerError

  static Message of(Kind k) { Message m; m.kind = k; return m; }
  static Message tensor(RawTensor t) { Message m = of(Tensor); m.x = std::move(t); return m; }
  static Message error(const std::string &s) { Message m = of(WorkerError); m.message = s; return m; }
  static Message worker_info(const WorkerInfo &i) { Message m = of(WorkerInfoMsg); m.info = i; return m; }

  std::vector<uint8_t> to_bytes() const {
    Writer w;
    w.u32((uint32_t)kind);
    switch (kind) {
      case WorkerInfoMsg: info.write(w); break;
      case SingleOp: w.string(layer_name); x.write(w); w.u64(index_pos); w.u64(block_idx); break;
      case Batch:
        x.write(w);
        w.u32((uint32_t)batch.size());
        for (auto &o : batch) { w.string(std::get<0>(o)); w.u64(std::get<1>(o)); w.u64(std::get<2>(o)); }
        break;
      case Tensor: x.write(w); break;
      case LayerAssignment:
        w.u32((uint32_t)layers.size());
        for (auto &l : layers) w.string(l);
        w.string(model_hash);
        break;
      case LayerAssignmentAck: w.boolean(needs_data); break;
      case ModelDataChunk:
        w.string(filename); w.u64(offset); w.u64(total_size); w.boolean(compressed); w.u32(checksum);
        w.blob(data.data(), data.size());
        break;
      case ModelDataResume: w.string(filename); w.u64(offset); break;
      case WorkerError: w.string(message); break;
      default: break;  // Hello, Goodbye, ModelDataDone, WorkerReady: tag only
    }
    return std::move(w.buf);
  }

  static Message from_bytes(const uint8_t *p, size_t n) {
    Reader r(p, n);
    uint32_t tag = r.u32();
    if (tag >= N_KINDS) throw ProtocolError("unknown message tag " + std::to_string(tag));
    Message m = of((Kind)tag);
    switch (m.kind) {
      case WorkerInfoMsg: m.info = WorkerInfo::read(r); break;
      case SingleOp: m.layer_name = r.string(); m.x = RawTensor::read(r); m.index_pos = r.u64(); m.block_idx = r.u64(); break;
      case Batch: {
        m.x = RawTensor::read(r);
        uint32_t k = r.u32();
        for (uint32_t i = 0; i < k; i++) { std::string s = r.string(); uint64_t a = r.u64(), b = r.u64(); m.batch.emplace_back(s, a, b); }
        break;
      }
      case Tensor: m.x = RawTensor::read(r); break;
      case LayerAssignment: {
        uint32_t k = r.u32();
        for (uint32_t i = 0; i < k; i++) m.layers.push_back(r.string());
        m.model_hash = r.string();
        break;
      }
      case LayerAssignmentAck: m.needs_data = r.boolean(); break;
      case ModelDataChunk:
        m.filename = r.string(); m.offset = r.u64(); m.total_size = r.u64(); m.compressed = r.boolean(); m.checksum = r.u32();
        m.data = r.blob();
        break;
      case ModelDataResume: m.filename = r.string(); m.offset = r.u64(); break;
      case WorkerError: m.message = r.string(); break;
      default: break;
    }
    r.done();
    return m;
  }
};

// ------------------------------------------------------------------------------------------ socket I/O + framing
inline void send_all(int fd, const uint8_t *p, size_t n) {
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) throw ConnectionClosed("send failed");
    p += k; n -= (size_t)k;
  }
}
inline void recv_exact(int fd, uint8_t *p, size_t n) {
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) throw ConnectionClosed("connection closed by peer");
    p += k; n -= (size_t)k;
  }
}
inline size_t write_message(int fd, const Message &m) {  // message.rs:372-394
  std::vector<uint8_t> payload = m.to_bytes();
  if (payload.size() > MESSAGE_MAX_SIZE) throw ProtocolError("request size " + std::to_string(payload.size()) + " > MESSAGE_MAX_SIZE");
  Writer h;
  h.u32(PROTO_MAGIC);
  h.u32((uint32_t)payload.size());
  h.buf.insert(h.buf.end(), payload.begin(), payload.end());
  send_all(fd, h.buf.data(), h.buf.size());
  return h.buf.size();
}
inline Message read_message(int fd, std::vector<uint8_t> &buf) {  // message.rs:334-362
  uint8_t hdr[8];
  recv_exact(fd, hdr, 8);
  Reader r(hdr, 8);
  uint32_t magic = r.u32(), size = r.u32();
  if (magic != PROTO_MAGIC) throw ProtocolError("invalid magic value: " + std::to_string(magic));
  if (size > MESSAGE_MAX_SIZE) throw ProtocolError("request size " + std::to_string(size) + " > MESSAGE_MAX_SIZE");
  buf.resize(size);
  recv_exact(fd, buf.data(), size);
  return Message::from_bytes(buf.data(), size);
}

// ------------------------------------------------------------------------------------------ SHA-256 / HMAC (auth.rs)
struct Sha256 {
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  uint8_t blk[64];
  size_t fill = 0;
  uint64_t total = 0;
  static uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
  void compress(const uint8_t *b) {
    static const uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)b[4 * i] << 24) | ((uint32_t)b[4 * i + 1] << 16) | ((uint32_t)b[4 * i + 2] << 8) | b[4 * i + 3];
    for (int i = 16; i < 64; i++) {
      uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], bb = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
      uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g), t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), mj = (a & bb) ^ (a & c) ^ (bb & c), t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = bb; bb = a; a = t1 + t2;
    }
    h[0] += a; h[1] += bb; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  void update(const uint8_t *p, size_t n) {
    total += n;
    while (n) {
      size_t k = std::min(n, 64 - fill);
      memcpy(blk + fill, p, k);
      fill += k; p += k; n -= k;
      if (fill == 64) { compress(blk); fill = 0; }
    }
  }
  void finish(uint8_t out[32]) {
    uint64_t bits = total * 8;
    uint8_t pad = 0x80;
    update(&pad, 1);
    uint8_t z = 0;
    while (fill != 56) update(&z, 1);
    uint8_t len[8];
    for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i));
    update(len, 8);
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(h[i] >> (24 - 8 * j));
  }
};

inline void hmac_sha256(const uint8_t *key, size_t klen, const uint8_t *data, size_t dlen, uint8_t out[32]) {
  uint8_t k[64] = {0};
  if (klen > 64) { Sha256 s; s.update(key, klen); s.finish(k); } else memcpy(k, key, klen);
  uint8_t ipad[64], opad[64], inner[32];
  for (int i = 0; i < 64; i++) { ipad[i] = k[i] ^ 0x36; opad[i] = k[i] ^ 0x5c; }
  Sha256 a; a.update(ipad, 64); a.update(data, dlen); a.finish(inner);
  Sha256 b; b.update(opad, 64); b.update(inner, 32); b.finish(out);
}
inline bool constant_time_eq(const uint8_t *a, const uint8_t *b, size_t n) {  // auth.rs:41-50
  uint8_t acc = 0;
  for (size_t i = 0; i < n; i++) acc |= a[i] ^ b[i];
  return acc == 0;
}
inline void random_nonce(uint8_t out[32]) {
  std::random_device rd;
  for (int i = 0; i < 32; i += 4) { uint32_t v = rd(); memcpy(out + i, &v, 4); }
}
inline void authenticate_as_worker(int fd, const std::string &key) {  // auth.rs:89-118
  uint8_t master_nonce[32], resp[64], master_hmac[32], expect[32];
  recv_exact(fd, master_nonce, 32);
  hmac_sha256((const uint8_t *)key.data(), key.size(), master_nonce, 32, resp);
  random_nonce(resp + 32);
  send_all(fd, resp, 64);
  recv_exact(fd, master_hmac, 32);
  hmac_sha256((const uint8_t *)key.data(), key.size(), resp + 32, 32, expect);
  if (!constant_time_eq(master_hmac, expect, 32)) throw ProtocolError("master authentication failed: invalid HMAC");
}
inline void authenticate_as_master(int fd, const std::string &key) {  // auth.rs:54-84
  uint8_t nonce[32], resp[64], expect[32], mine[32];
  random_nonce(nonce);
  send_all(fd, nonce, 32);
  recv_exact(fd, resp, 64);
  hmac_sha256((const uint8_t *)key.data(), key.size(), nonce, 32, expect);
  if (!constant_time_eq(resp, expect, 32)) throw ProtocolError("worker authentication failed: invalid HMAC");
  hmac_sha256((const uint8_t *)key.data(), key.size(), resp + 32, 32, mine);
  send_all(fd, mine, 32);
}

// ------------------------------------------------------------------------------------------ worker side
// One Session per master connection = one KV cache (the reference clones a fresh cache per connection, worker.rs:60-75;
// a cake master opens one connection per remote LAYER, text_model.rs:211-227, so many are open at once).
struct Session {
  virtual ~Session() = default;
  virtual void clear_cache() = 0;
  virtual RawTensor forward_ops(const RawTensor &x, const std::vector<Op> &ops) = 0;  // throws std::exception with the reason
};
struct Backend {  // the compute behind a worker endpoint
  virtual ~Backend() = default;
  virtual std::string dtype() const = 0;                 // "BF16" | "F16" (Debug form of candle's DType, worker.rs:55)
  virtual std::string device() const = 0;                // "cuda" | "cpu"
  virtual uint64_t device_idx() const { return 0; }
  virtual std::unique_ptr<Session> new_session() = 0;    // called from the connection's thread
};

class WireWorker {  // worker.rs:79-597; one thread per master connection (the reference: one tokio task)
  Backend &be_;
  std::string key_;
  bool has_key_;
  int lfd_ = -1;
 public:
  std::string address;
  std::atomic<size_t> served{0};
  bool reflect = false;  // test aid: after the handshake, answer every message with the same message re-encoded

  WireWorker(Backend &be, const std::string &host, int port, const std::string *cluster_key) : be_(be), key_(cluster_key ? *cluster_key : ""), has_key_(cluster_key != nullptr) {
    lfd_ = ::socket(AF_INET, SOCK_STREAM, 0);
    if (lfd_ < 0) throw std::runtime_error("socket() failed");
    int one = 1;
    setsockopt(lfd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
    sockaddr_in a{};
    a.sin_family = AF_INET;
    a.sin_port = htons((uint16_t)port);
    if (inet_pton(AF_INET, host.c_str(), &a.sin_addr) != 1) throw std::runtime_error("bad listen address " + host);
    if (::bind(lfd_, (sockaddr *)&a, sizeof a) != 0) throw std::runtime_error("can't bind " + host + ":" + std::to_string(port));
    if (::listen(lfd_, 64) != 0) throw std::runtime_error("listen() failed");
    socklen_t len = sizeof a;
    getsockname(lfd_, (sockaddr *)&a, &len);
    address = host + ":" + std::to_string(ntohs(a.sin_port));
  }
  ~WireWorker() { if (lfd_ >= 0) ::close(lfd_); }

  WorkerInfo to_info(uint64_t latency_ms) const {  // worker.rs:47-57
    utsname u{};
    uname(&u);
    WorkerInfo i;
    i.version = "cake-b200";
    i.dtype = be_.dtype();
    i.os = "linux";
    i.arch = u.machine;
    i.device = be_.device();
    i.device_idx = be_.device_idx();
    i.latency_lo = latency_ms;
    return i;
  }

  void handle_master_client(int fd) {  // worker.rs:298-575
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
    if (has_key_) authenticate_as_worker(fd, key_);
    std::vector<uint8_t> buf;
    auto t0 = std::chrono::steady_clock::now();
    Message first = read_message(fd, buf);
    auto ms = [&](std::chrono::steady_clock::time_point s) {
      return (uint64_t)std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - s).count();
    };
    if (first.kind == Message::LayerAssignment) {  // master re-running setup against a running worker (:316-329)
      Message ack = Message::of(Message::LayerAssignmentAck);
      write_message(fd, ack);
      write_message(fd, Message::of(Message::WorkerReady));
      return;
    }
    if (first.kind != Message::Hello) throw ProtocolError("unexpected first message (expected Hello)");
    std::unique_ptr<Session> sess = be_.new_session();  // a fresh, empty cache for this connection
    write_message(fd, Message::worker_info(to_info(ms(t0))));
    for (;;) {
      auto t1 = std::chrono::steady_clock::now();
      Message m;
      try {
        m = read_message(fd, buf);
      } catch (const ConnectionClosed &) {
        return;  // the reference's `while let Ok(..)` ends the same way
      }
      if (reflect) { write_message(fd, m); continue; }
      if (m.kind == Message::Goodbye) {  // :363-383
        sess->clear_cache();
        write_message(fd, Message::worker_info(to_info(ms(t1))));
        continue;
      }
      std::vector<Op> ops;
      if (m.kind == Message::SingleOp) ops.emplace_back(m.layer_name, m.index_pos, m.block_idx);
      else if (m.kind == Message::Batch) ops = m.batch;
      else throw ProtocolError("unhandled message in loop");
      try {
        if (ops.empty()) throw std::runtime_error("empty batch");
        m.x.validate();
        RawTensor y = sess->forward_ops(m.x, ops);
        write_message(fd, Message::tensor(std::move(y)));
        served++;
      } catch (const ConnectionClosed &) {
        throw;
      } catch (const std::exception &e) {  // :490-520: report, keep the connection
        write_message(fd, Message::error(e.what()));
      }
    }
  }

  // Accept loop (worker.rs:577-597): one thread per connection.  max_connections < 0: forever; otherwise return once
  // that many connections have been accepted and have ended.
  void serve(int max_connections = -1) {
    // Each connection thread raises its `done` flag as its last action; the accept loop only ever joins threads
    // whose flag is up, so a live (persistent) master connection is never waited on from here.
    struct Conn {
      std::thread th;
      std::shared_ptr<std::atomic<bool>> done;
    };
    std::list<Conn> conns;
    auto reap = [&conns]() {
      for (auto it = conns.begin(); it != conns.end();) {
        if (it->done->load(std::memory_order_acquire)) {
          it->th.join();
          it = conns.erase(it);
        } else {
          ++it;
        }
      }
    };
    for (int n = 0; max_connections < 0 || n < max_connections; n++) {
      int fd = ::accept(lfd_, nullptr, nullptr);
      if (fd < 0) break;
      auto done = std::make_shared<std::atomic<bool>>(false);
      conns.push_back(Conn{std::thread([this, fd, done]() {
        try {
          handle_master_client(fd);
        } catch (const std::exception &e) {
          fprintf(stderr, "[worker] connection ended: %s\n", e.what());
        }
        ::close(fd);
        done->store(true, std::memory_order_release);
      }), done});
      if (conns.size() > 64) reap();  // finished connections only
    }
    for (auto &c : conns) c.th.join();
  }
};

// A backend that echoes the activation (tests/protocol.rs MockWorker); "model.layers.99" is reported missing.
struct EchoBackend : Backend {
  std::atomic<size_t> sessions{0};
  struct S : Session {
    void clear_cache() override {}
    RawTensor forward_ops(const RawTensor &x, const std::vector<Op> &ops) override {
      for (auto &o : ops)
        if (std::get<0>(o) == "model.layers.99") throw std::runtime_error("could not find layer " + std::get<0>(o));
      return x;
    }
  };
  std::string dtype() const override { return "F16"; }
  std::string device() const override { return "cpu"; }
  std::unique_ptr<Session> new_session() override { sessions++; return std::unique_ptr<Session>(new S()); }
};

}  // namespace cake_wire
