// cake_b200.cu — C ABI of libcake_b200.so (see include/cake_b200.h for the contract and the
// reference interfaces each entry replaces).  Host-side logic only: handles, buffers, kernel plans,
// the CUDA graph of the decode step and the NCCL p2p hand-off.  All math is in the .cuh kernels.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <algorithm>
#include <chrono>
#include <mutex>
#include <thread>
#include <set>
#include <string>
#include <vector>

#include "../../include/cake_b200.h"
#include "attn_decode.cuh"
#include "attn_prefill.cuh"
#include "attn_prefill_tc.cuh"
#include "common.cuh"
#include "decode_mega.cuh"
#include "gemm_tc.cuh"
#include "gemv.cuh"
#include "prefill.cuh"
#include "sample.cuh"

using namespace cake;
typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) return fail(CAKE_B200_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)
#define RC(call)              \
  do {                        \
    int rc_ = (call);         \
    if (rc_ != CAKE_B200_OK) return rc_; \
  } while (0)

struct cake_b200_ctx;
static int wait_loads(cake_b200_ctx *c);
extern "C" const char *cake_b200_last_error(void) { return g_err; }
extern "C" const char *cake_b200_version(void) { return "cake_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------ NCCL (dlopen, like the reference's rocm/ffi.rs function table)
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
};
static NcclApi g_nccl;
static int nccl_load() {
  if (g_nccl.lib) return CAKE_B200_OK;
  const char *names[] = {getenv("CAKE_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    if (!n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(CAKE_B200_ENCCL, "cannot dlopen libnccl.so.2 (set CAKE_B200_NCCL_LIB): %s", dlerror());
#define SYM(field, name)                                                        \
  *(void **)(&g_nccl.field) = dlsym(h, name);                                   \
  if (!g_nccl.field) return fail(CAKE_B200_ENCCL, "libnccl lacks symbol %s", name);
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GetErrorString, "ncclGetErrorString");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
#undef SYM
  g_nccl.lib = h;
  return CAKE_B200_OK;
}
#define NC(call)                                                                                           \
  do {                                                                                                     \
    ncclResult_t r_ = (call);                                                                              \
    if (r_ != ncclSuccess) return fail(CAKE_B200_ENCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); \
  } while (0)

// ------------------------------------------------------------------------------------------ handles
struct cake_b200_ctx {
  int device = 0;
  cake_b200_config cfg{};
  int es = 2, sm_count = 148, rot = 0, nqkv = 0;
  cudaStream_t stream = nullptr;
  void *cos_t = nullptr, *sin_t = nullptr;
  // head
  void *embed = nullptr, *ln_f = nullptr, *lm_head = nullptr;
  // decode workspaces
  void *xa = nullptr, *xb = nullptr, *x0 = nullptr, *qkv = nullptr, *y = nullptr, *mm = nullptr, *logits = nullptr;
  float *ws_ml = nullptr, *ws_acc = nullptr, *part_val = nullptr;
  int *part_idx = nullptr, *d_step = nullptr;
  unsigned *attn_counters = nullptr, *argmax_counter = nullptr;
  uint32_t *d_token = nullptr, *token_ring = nullptr, *d_ids = nullptr, *d_pen = nullptr;
  // weight upload pipe: mmapped safetensors -> two pinned staging buffers (filled by a few host
  // threads) -> cudaMemcpy(2D)Async on a copy stream straight into the fused device layouts
  cudaStream_t load_stream = nullptr;
  void *stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_free[2] = {nullptr, nullptr}, load_done = nullptr;
  int stage_cur = 0;
  bool load_pending = false;
  double load_bytes = 0, load_seconds = 0;
  float *samp_p = nullptr, *samp_noise = nullptr;  // sampler scratch: probabilities (vocab) / supplied uniforms
  cake_b200_sampling sampling{};                   // sampler of the decode graph (kind 0 = greedy)
  size_t d_ids_cap = 0, d_pen_cap = 0;
  uint32_t *h_pin = nullptr;  // pinned staging (tokens)
  void *h_pin_x = nullptr;
  size_t h_pin_x_cap = 0;
  int nsplit = 1;
  // prefill workspaces
  size_t pf_rows = 0;
  void *pf_h = nullptr, *pf_qkv = nullptr, *pf_y = nullptr, *pf_x1 = nullptr, *pf_gu = nullptr, *pf_mm = nullptr;
  void *io_x = nullptr;
  size_t io_x_cap = 0;
  // comm
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // decode graph
  cudaGraphExec_t gexec = nullptr;
  cudaGraph_t graph = nullptr;
  cake_b200_cache *g_cache = nullptr;
  std::vector<int> g_block_idx;
  uint64_t g_kernels = 0;
  int g_rank = 0, g_world = 1;
  long steps_done = 0;
  uint64_t launches = 0;
  bool capturing = false;
  // persistent decode megakernel (decode_mega.cuh); CAKE_B200_PER_OP=1 selects the per-op kernels instead
  bool use_mega = true;
  unsigned long long *gbar = nullptr;   // [0] grid-barrier counter, [1] launch epoch
  MkLayer *mk_tab_dev = nullptr, *mk_tab_host = nullptr, *g_tab_dev = nullptr;
  std::vector<const void *> mk_sig;
  unsigned long long *trace = nullptr;  // CAKE_B200_MEGA_TRACE=1
  unsigned long long *step_trace = nullptr;  // per-step entry / input-acquired / exit stamps of the decode kernels
  unsigned *tickets = nullptr;          // per-phase work-claim counters of the megakernel
  // ring hand-off over NVLink peer memory: inbox = {counter (u64), pad, x[hidden]} in our memory, written by rank-1
  unsigned char *inbox = nullptr, *peer_inbox = nullptr;
  unsigned long long *ring_seq = nullptr;
};
constexpr int TOKEN_RING = 1 << 16;

// Live contexts: a cache may outlive its ctx (handles are freed in any order by foreign hosts), so cache_free only
// touches ctx state when the ctx is still registered here.
static std::mutex g_live_mu;
static std::set<cake_b200_ctx *> g_live_ctx;
static std::atomic<uint64_t> g_cache_gen{1};

struct cake_b200_block {
  cake_b200_ctx *ctx;
  int layer;
  int device = 0;  // copied so that block_free never dereferences a ctx that was destroyed first
  void *wqkv = nullptr, *wo = nullptr, *wgu = nullptr, *wd = nullptr, *ln1 = nullptr, *ln2 = nullptr;
  void *bqkv = nullptr, *qn = nullptr, *kn = nullptr;
  // cake_b200_block_set_variant: the sibling block structures (olmo2 / gemma3 / exaone4 block.rs)
  void *pan = nullptr, *pfn = nullptr;  // post-attention / post-feedforward RmsNorm vectors
  int window = -1;                      // -1: the config's
  bool use_rope = true;
  bool variant() const { return !ln1 || !ln2 || pan || pfn || window >= 0 || !use_rope || ctx->cfg.pre_reshape_qk_norm; }
};

struct cake_b200_cache {
  cake_b200_ctx *ctx;
  int batch, cap;
  int device = 0;
  uint64_t gen = 0;  // unique per cache object: part of the megakernel layer-table signature (addresses can be reused)
  std::vector<void *> k, v;
  std::vector<int> len;
  int *d_pos = nullptr;
};

template <typename U> struct TypeTag { typedef U type; };
template <typename F> static inline int dispatch_T(int dtype, F &&f) {
  if (dtype == CAKE_B200_BF16) return f(TypeTag<__nv_bfloat16>{});
  return f(TypeTag<__half>{});
}
// usage: DISPATCH_T(dtype, [&](auto tag_) -> int { typedef typename decltype(tag_)::type T; ... })
#define DISPATCH_T(dtype, ...) dispatch_T((dtype), __VA_ARGS__)
#define T_LAMBDA [&](auto tag_) -> int

// (head_dim, query heads per kv head) combinations the megakernel is instantiated for
#define MK_FOR_ALL(X) X(16, 2) X(16, 4) X(64, 2) X(64, 4) X(128, 1) X(128, 2) X(128, 4) X(128, 8)

// ------------------------------------------------------------------------------------------ launch helper (PDL)
template <typename... KArgs, typename... Args>
static int launch_pdl(cake_b200_ctx *c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  CU(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
  if (c->capturing) c->g_kernels++;
  else c->launches++;
  return CAKE_B200_OK;
}

// the same with a thread-block cluster of `cluster` CTAs along x (TMA multicast between the CTAs of a pair)
template <typename... KArgs, typename... Args>
static int launch_cluster_pdl(cake_b200_ctx *c, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, unsigned cluster, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  at[1].id = cudaLaunchAttributeClusterDimension;
  at[1].val.clusterDim.x = cluster;
  at[1].val.clusterDim.y = 1;
  at[1].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 2;
  CU(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
  if (c->capturing) c->g_kernels++;
  else c->launches++;
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ GEMV plan + launch
struct GemvPlan {
  int KC, RS, WPR, RPW, n_stages, max_rows, grid;
  size_t smem;
};
static int plan_gemv(const cake_b200_ctx *c, int N, int K, int G, GemvPlan *p) {
  const int es = c->es;
  if (K % 64 != 0 || N % G != 0) return fail(CAKE_B200_EINVAL, "gemv: K=%d must be a multiple of 64, N=%d of %d", K, N, G);
  // segment: whole rows up to 16 KB, else the row is halved until it is (K=14336 -> 2 x 14 KB)
  int KC = K;
  while ((size_t)KC * es > 16384 && (KC / 2) % 64 == 0) KC /= 2;
  const size_t seg = (size_t)KC * es;
  // column slices: keep >= 64 16-byte vectors per warp per row where the row is long enough
  int WPR = 1;
  while (WPR < 8 && (KC / 8) / (WPR * 2) >= 64) WPR *= 2;
  if (WPR > 4) WPR = 4;
  const int slots = GEMV_CONSUMER_WARPS / WPR;
  int RPW = 1;  // rows per warp per stage: grow the stage towards 32 KB
  while (RPW < 4 && (size_t)(slots * RPW * 2) * seg <= 32768) RPW *= 2;
  const int RS = slots * RPW;
  const int units = N / G;
  const int grid = units < c->sm_count ? units : c->sm_count;
  const int max_rows = (units / grid + 1) * G;
  // two CTAs (this kernel + its PDL successor) must fit in one SM's 227 KB
  const size_t budget = ((size_t)K * es > 16384 ? 116 : 108) * 1024;
  int ns = GEMV_MAX_STAGES;
  while (ns > 2 && gemv_smem_bytes(K, KC, RS, WPR, ns, max_rows, es) > budget) ns--;
  const size_t smem = gemv_smem_bytes(K, KC, RS, WPR, ns, max_rows, es);
  if (smem > 227 * 1024) return fail(CAKE_B200_EINVAL, "gemv: N=%d K=%d needs %zu B of shared memory", N, K, smem);
  *p = GemvPlan{KC, RS, WPR, RPW, ns, max_rows, grid, smem};
  return CAKE_B200_OK;
}

template <typename T, int EPI> static int launch_gemv_T(cake_b200_ctx *c, GemvArgs a, const GemvPlan &p) {
  a.KC = p.KC;
  a.RS = p.RS;
  a.WPR = p.WPR;
  a.n_stages = p.n_stages;
  a.max_rows = p.max_rows;
  dim3 grid(p.grid), block(GEMV_THREADS);
  switch (p.RPW) {
    case 1: return launch_pdl(c, gemv_kernel<T, EPI, 1>, grid, block, p.smem, a);
    case 2: return launch_pdl(c, gemv_kernel<T, EPI, 2>, grid, block, p.smem, a);
    default: return launch_pdl(c, gemv_kernel<T, EPI, 4>, grid, block, p.smem, a);
  }
}
template <int EPI> static int launch_gemv(cake_b200_ctx *c, GemvArgs a) {
  GemvPlan p;
  RC(plan_gemv(c, a.N, a.K, EPI == EPI_SWIGLU ? 2 : 1, &p));
  return DISPATCH_T(c->cfg.dtype, T_LAMBDA {
      typedef typename decltype(tag_)::type T; return launch_gemv_T<T, EPI>(c, a, p); });
}

template <typename T> static int set_smem_attrs_T() {
  const int maxs = 227 * 1024;
  // max-shared carveout so that a kernel and its programmatic-dependent successor co-reside on an SM
#define SETA1(EPI, R)                                                                                          \
  CU(cudaFuncSetAttribute(gemv_kernel<T, EPI, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs));         \
  CU(cudaFuncSetAttribute(gemv_kernel<T, EPI, R>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
#define SETA(EPI) SETA1(EPI, 1) SETA1(EPI, 2) SETA1(EPI, 4)
  SETA(EPI_PLAIN) SETA(EPI_RESIDUAL) SETA(EPI_SWIGLU) SETA(EPI_ARGMAX)
#undef SETA
#undef SETA1
#define SETB(HD)                                                                                                  \
  CU(cudaFuncSetAttribute(attn_decode_kernel<T, HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs - 4096)); /* static smem counts too */         \
  CU(cudaFuncSetAttribute(attn_decode_kernel<T, HD>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  SETB(16) SETB(32) SETB(64) SETB(128) SETB(256)
#undef SETB
#define SETC(HD, G)                                                                                                  \
  CU(cudaFuncSetAttribute(decode_mega_kernel<T, HD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxs - 4096));  \
  CU(cudaFuncSetAttribute(decode_mega_kernel<T, HD, G>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  MK_FOR_ALL(SETC)
#undef SETC
  CU(cudaFuncSetAttribute(attn_prefill_mma_kernel<T, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * FA_BN * 128 * 2));
  CU(cudaFuncSetAttribute(attn_prefill_mma_kernel<T, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * FA_BN * 64 * 2));
  CU(cudaFuncSetAttribute(attn_prefill_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, FT_SMEM_BYTES));
  CU(cudaFuncSetAttribute(attn_prefill_tc_kernel<T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
#define SETT(EPI)                                                                                                                  \
  CU(cudaFuncSetAttribute(gemm_tc_kernel<T, EPI, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM_BYTES));      \
  CU(cudaFuncSetAttribute(gemm_tc_kernel<T, EPI, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES));      \
  CU(cudaFuncSetAttribute(gemm_tc_kernel<T, EPI, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES));
  SETT(TCE_PLAIN) SETT(TCE_RESIDUAL) SETT(TCE_SWIGLU)
#undef SETT
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ context
static void rope_tables_host(const cake_b200_config &c, std::vector<float> &cs, std::vector<float> &sn, int rot) {
  // cache.rs:43-96 (same arithmetic the reference performs on the host; computed once per ctx)
  const int half = rot / 2;
  std::vector<float> theta(half > 0 ? half : 1);
  for (int j = 0; j < half; j++) theta[j] = 1.0f / powf(c.rope_theta, (float)(2 * j) / (float)rot);
  if (c.rope_llama3 && c.rope_orig_max > 0) {
    const float old = (float)c.rope_orig_max;
    const float low_wl = old / c.rope_low, high_wl = old / c.rope_high;
    for (int j = 0; j < half; j++) {
      const float wl = 2.0f * 3.14159265358979323846f / theta[j];
      if (wl < high_wl) {
      } else if (wl > low_wl) {
        theta[j] /= c.rope_factor;
      } else {
        const float smooth = (old / wl - c.rope_low) / (c.rope_high - c.rope_low);
        theta[j] = (1.0f - smooth) * (theta[j] / c.rope_factor) + smooth * theta[j];
      }
    }
  }
  cs.resize((size_t)c.max_seq * half);
  sn.resize((size_t)c.max_seq * half);
  for (int p = 0; p < c.max_seq; p++)
    for (int j = 0; j < half; j++) {
      const float ang = (float)p * theta[j];
      cs[(size_t)p * half + j] = cosf(ang);
      sn[(size_t)p * half + j] = sinf(ang);
    }
}
static void f32_to_D(const std::vector<float> &src, std::vector<uint16_t> &dst, int dtype) {
  dst.resize(src.size());
  for (size_t i = 0; i < src.size(); i++) {
    if (dtype == CAKE_B200_BF16) {
      bf16 v = __float2bfloat16_rn(src[i]);
      memcpy(&dst[i], &v, 2);
    } else {
      __half v = __float2half_rn(src[i]);
      memcpy(&dst[i], &v, 2);
    }
  }
}

extern "C" int cake_b200_ctx_create(int device, const cake_b200_config *cfg, cake_b200_ctx **out) {
  if (!cfg || !out) return fail(CAKE_B200_EINVAL, "null argument");
  if (cfg->dtype != CAKE_B200_BF16 && cfg->dtype != CAKE_B200_F16)
    return fail(CAKE_B200_EINVAL, "dtype %d unsupported (bf16=0, f16=1)", cfg->dtype);
  if (cfg->hidden < 1 || cfg->inter < 1 || cfg->n_heads < 1 || cfg->n_kv_heads < 1 || cfg->n_layers < 1 || cfg->vocab < 1 || cfg->max_seq < 1)
    return fail(CAKE_B200_EINVAL, "config dimensions must be positive");
  if (cfg->n_heads % cfg->n_kv_heads != 0 || cfg->n_heads / cfg->n_kv_heads > ATTN_MAX_G)
    return fail(CAKE_B200_EINVAL, "n_heads/n_kv_heads must be an integer <= %d", ATTN_MAX_G);
  if (cfg->head_dim != 16 && cfg->head_dim != 32 && cfg->head_dim != 64 && cfg->head_dim != 128 && cfg->head_dim != 256)
    return fail(CAKE_B200_EINVAL, "head_dim %d unsupported (16/32/64/128/256)", cfg->head_dim);
  if (cfg->hidden % 64 || cfg->inter % 64 || (cfg->n_heads * cfg->head_dim) % 64)
    return fail(CAKE_B200_EINVAL, "hidden, intermediate and n_heads*head_dim must be multiples of 64");
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (ndev == 0) return fail(CAKE_B200_ECUDA, "no CUDA device: libcake_b200 has no CPU fallback");
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(CAKE_B200_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  auto *c = new cake_b200_ctx();
  // any failure below returns through CU / RC: release what was set up so far (ctx_destroy tolerates a half-built ctx)
  struct Guard { cake_b200_ctx *c; ~Guard() { if (c) { cake_b200_ctx_destroy(c); (void)cudaGetLastError(); } } } guard{c};
  c->device = device;
  c->cfg = *cfg;
  c->es = 2;
  c->sm_count = prop.multiProcessorCount;
  c->rot = (int)((float)cfg->head_dim * cfg->partial_rotary);
  c->nqkv = (cfg->n_heads + 2 * cfg->n_kv_heads) * cfg->head_dim;
  CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  RC(set_smem_attrs_T<__nv_bfloat16>());
  RC(set_smem_attrs_T<__half>());
  // RoPE tables -> D
  std::vector<float> cs, sn;
  rope_tables_host(*cfg, cs, sn, c->rot);
  std::vector<uint16_t> cd, sd;
  f32_to_D(cs, cd, cfg->dtype);
  f32_to_D(sn, sd, cfg->dtype);
  const size_t tb = cd.size() * 2 + 16;
  CU(cudaMalloc(&c->cos_t, tb));
  CU(cudaMalloc(&c->sin_t, tb));
  CU(cudaMemcpy(c->cos_t, cd.data(), cd.size() * 2, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(c->sin_t, sd.data(), sd.size() * 2, cudaMemcpyHostToDevice));
  // decode workspaces
  const int H = cfg->hidden, I = cfg->inter, nh = cfg->n_heads, hd = cfg->head_dim;
  c->nsplit = c->sm_count / cfg->n_kv_heads;
  if (c->nsplit < 1) c->nsplit = 1;
  if (c->nsplit > ATTN_MAX_SPLIT) c->nsplit = ATTN_MAX_SPLIT;
  CU(cudaMalloc(&c->xa, (size_t)H * 2));
  CU(cudaMalloc(&c->xb, (size_t)H * 2));
  CU(cudaMalloc(&c->x0, (size_t)H * 2));
  CU(cudaMalloc(&c->qkv, (size_t)c->nqkv * 2));
  CU(cudaMalloc(&c->y, (size_t)nh * hd * 2));
  CU(cudaMalloc(&c->mm, (size_t)I * 2));
  CU(cudaMalloc(&c->logits, (size_t)cfg->vocab * 2));
  CU(cudaMalloc(&c->ws_ml, (size_t)nh * c->nsplit * 2 * 4));
  CU(cudaMalloc(&c->ws_acc, (size_t)nh * c->nsplit * hd * 4));
  CU(cudaMalloc(&c->attn_counters, (size_t)cfg->n_kv_heads * 4));
  CU(cudaMemset(c->attn_counters, 0, (size_t)cfg->n_kv_heads * 4));
  CU(cudaMalloc(&c->part_val, 1024 * 4));
  CU(cudaMalloc(&c->part_idx, 1024 * 4));
  CU(cudaMalloc(&c->argmax_counter, 4));
  CU(cudaMemset(c->argmax_counter, 0, 4));
  CU(cudaMalloc(&c->d_token, 4));
  CU(cudaMalloc(&c->d_step, 4));
  CU(cudaMemset(c->d_step, 0, 4));
  CU(cudaMalloc(&c->token_ring, (size_t)TOKEN_RING * 4));
  CU(cudaMallocHost(&c->h_pin, 64 * 4));
  CU(cudaMalloc(&c->gbar, 16));
  CU(cudaMemset(c->gbar, 0, 16));
  CU(cudaMalloc(&c->tickets, sizeof(unsigned) * (4 * MK_MAX_LAYERS + 4)));
  CU(cudaMalloc(&c->ring_seq, 8));
  CU(cudaMemset(c->ring_seq, 0, 8));
  CU(cudaMalloc(&c->step_trace, sizeof(unsigned long long) * MK_STEP_RING * 8));
  CU(cudaMemset(c->step_trace, 0, sizeof(unsigned long long) * MK_STEP_RING * 8));
  CU(cudaMalloc(&c->mk_tab_dev, sizeof(MkLayer) * MK_MAX_LAYERS));
  CU(cudaMalloc(&c->g_tab_dev, sizeof(MkLayer) * MK_MAX_LAYERS));
  CU(cudaMallocHost(&c->mk_tab_host, sizeof(MkLayer) * MK_MAX_LAYERS));
  {
    const char *e = getenv("CAKE_B200_PER_OP");
    c->use_mega = !(e && e[0] == '1');
    const char *t = getenv("CAKE_B200_MEGA_TRACE");
    if (t && t[0] == '1') {
      CU(cudaMalloc(&c->trace, 8 * 4096));
      CU(cudaMemset(c->trace, 0, 8 * 4096));
    }
  }
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live_ctx.insert(c);
  }
  guard.c = nullptr;
  *out = c;
  return CAKE_B200_OK;
}

extern "C" void cake_b200_ctx_destroy(cake_b200_ctx *c) {
  if (!c) return;
  {
    std::lock_guard<std::mutex> lk(g_live_mu);
    g_live_ctx.erase(c);
  }
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->gexec) cudaGraphExecDestroy(c->gexec);
  if (c->graph) cudaGraphDestroy(c->graph);
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (c->peer_inbox) cudaIpcCloseMemHandle(c->peer_inbox);
  void *bufs[] = {c->cos_t, c->sin_t, c->embed, c->ln_f, c->xa, c->xb, c->qkv, c->y, c->mm, c->logits, c->ws_ml,
                  c->ws_acc, c->part_val, c->part_idx, c->d_step, c->attn_counters, c->argmax_counter, c->d_token,
                  c->token_ring, c->d_ids, c->d_pen, c->pf_h, c->pf_qkv, c->pf_y, c->pf_x1, c->pf_gu, c->pf_mm, c->io_x,
                  c->gbar, c->mk_tab_dev, c->g_tab_dev, c->tickets, c->ring_seq, c->inbox, c->step_trace, c->trace,
                  c->samp_p, c->samp_noise, c->x0};
  for (void *b : bufs)
    if (b) cudaFree(b);
  if (c->lm_head && c->lm_head != c->embed) cudaFree(c->lm_head);
  for (int i = 0; i < 2; i++) {
    if (c->stage[i]) cudaFreeHost(c->stage[i]);
    if (c->stage_free[i]) cudaEventDestroy(c->stage_free[i]);
  }
  if (c->load_done) cudaEventDestroy(c->load_done);
  if (c->load_stream) { cudaStreamSynchronize(c->load_stream); cudaStreamDestroy(c->load_stream); }
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->h_pin_x) cudaFreeHost(c->h_pin_x);
  if (c->mk_tab_host) cudaFreeHost(c->mk_tab_host);
  cudaStreamDestroy(c->stream);
  delete c;
}
extern "C" int cake_b200_sync(cake_b200_ctx *c) {
  if (!c) return fail(CAKE_B200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  RC(wait_loads(c));
  if (c->load_stream) CU(cudaStreamSynchronize(c->load_stream));
  CU(cudaStreamSynchronize(c->stream));
  return CAKE_B200_OK;
}
extern "C" void *cake_b200_stream(cake_b200_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" int cake_b200_launch_count(cake_b200_ctx *c, uint64_t *k) {
  if (!c || !k) return fail(CAKE_B200_EINVAL, "null argument");
  *k = c->launches;
  return CAKE_B200_OK;
}

extern "C" int cake_b200_dev_alloc(cake_b200_ctx *c, size_t bytes, void **out) {
  if (!c || !out) return fail(CAKE_B200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaMalloc(out, bytes ? bytes : 16));
  return CAKE_B200_OK;
}
extern "C" int cake_b200_dev_free(cake_b200_ctx *c, void *p) {
  if (!c) return fail(CAKE_B200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaFree(p));
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ weight upload pipe
constexpr size_t STAGE_BYTES = (size_t)64 << 20;
constexpr int STAGE_THREADS = 6;
static int pipe_init(cake_b200_ctx *c) {
  if (c->load_stream) return CAKE_B200_OK;
  CU(cudaStreamCreateWithFlags(&c->load_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    CU(cudaMallocHost(&c->stage[i], STAGE_BYTES));
    CU(cudaEventCreateWithFlags(&c->stage_free[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&c->load_done, cudaEventDisableTiming));
  return CAKE_B200_OK;
}
// host rows [r0, r1) of width bytes (source pitch spitch) -> packed staging buffer, split over a few threads: one thread
// tops out at ~6 GB/s out of the page cache, far below what the PCIe link takes
static void stage_rows(char *dst, const char *src, size_t spitch, size_t width, size_t rows) {
  const size_t total = rows * width;
  int nt = total >= ((size_t)4 << 20) ? STAGE_THREADS : 1;
  if (nt == 1) {
    for (size_t r = 0; r < rows; r++) memcpy(dst + r * width, src + r * spitch, width);
    return;
  }
  std::vector<std::thread> th;
  if (spitch == width) {  // contiguous: split by bytes
    for (int t = 0; t < nt; t++) {
      const size_t a = total * t / nt, b = total * (t + 1) / nt;
      th.emplace_back([=]() { memcpy(dst + a, src + a, b - a); });
    }
  } else {
    for (int t = 0; t < nt; t++) {
      const size_t a = rows * t / nt, b = rows * (t + 1) / nt;
      th.emplace_back([=]() { for (size_t r = a; r < b; r++) memcpy(dst + r * width, src + r * spitch, width); });
    }
  }
  for (auto &x : th) x.join();
}
// dst[r * dpitch .. + width) = src[r * spitch .. + width) for r < rows (rows == 1: a plain copy).  Device sources
// (synthetic checkpoints generated on the GPU) are copied device-to-device on the same stream.
static int pipe_copy(cake_b200_ctx *c, void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows) {
  RC(pipe_init(c));
  cudaPointerAttributes at{};
  const bool dev_src = (cudaPointerGetAttributes(&at, src) == cudaSuccess) && (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
  (void)cudaGetLastError();
  if (dev_src) {
    // written by someone else's stream (torch): copy on the legacy default stream, which orders after that work, and
    // wait — the caller may free the source as soon as we return
    CU(cudaMemcpy2D(dst, dpitch, src, spitch, width, rows, cudaMemcpyDeviceToDevice));
    CU(cudaStreamSynchronize(0));
    return CAKE_B200_OK;
  }
  const auto t0 = std::chrono::steady_clock::now();
  if (rows == 1 && width > STAGE_BYTES) {  // a long contiguous tensor: treat it as rows of 1 MB
    const size_t rw = (size_t)1 << 20, full = width / rw;
    if (full) RC(pipe_copy(c, dst, rw, src, rw, rw, full));
    if (width % rw) RC(pipe_copy(c, (char *)dst + full * rw, width % rw, (const char *)src + full * rw, width % rw, width % rw, 1));
    return CAKE_B200_OK;
  }
  const size_t rows_per = std::max<size_t>(1, STAGE_BYTES / width);
  for (size_t r0 = 0; r0 < rows; r0 += rows_per) {
    const size_t nr = std::min(rows_per, rows - r0);
    const int b = c->stage_cur;
    CU(cudaEventSynchronize(c->stage_free[b]));  // the DMA that last read this buffer has finished
    stage_rows((char *)c->stage[b], (const char *)src + r0 * spitch, spitch, width, nr);
    CU(cudaMemcpy2DAsync((char *)dst + r0 * dpitch, dpitch, c->stage[b], width, width, nr, cudaMemcpyHostToDevice, c->load_stream));
    CU(cudaEventRecord(c->stage_free[b], c->load_stream));
    c->stage_cur ^= 1;
  }
  c->load_bytes += (double)width * rows;
  c->load_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return CAKE_B200_OK;
}
static int pipe_mark(cake_b200_ctx *c) {  // compute must not start before the uploads issued so far have landed
  CU(cudaEventRecord(c->load_done, c->load_stream));
  c->load_pending = true;
  return CAKE_B200_OK;
}
static int wait_loads(cake_b200_ctx *c) {
  if (c->load_pending) {
    CU(cudaStreamWaitEvent(c->stream, c->load_done, 0));
    c->load_pending = false;
  }
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ blocks
static int upload(cake_b200_ctx *c, void **dst, const void *src, size_t bytes) {
  CU(cudaMalloc(dst, bytes + 16));
  return pipe_copy(c, *dst, bytes, src, bytes, bytes, 1);
}

extern "C" void cake_b200_block_free(cake_b200_block *b);
extern "C" int cake_b200_block_load(cake_b200_ctx *c, int layer_idx, const void *q, const void *k, const void *v,
                                    const void *o, const void *gate, const void *up, const void *down, const void *ln1,
                                    const void *ln2, const void *q_bias, const void *k_bias, const void *v_bias,
                                    const void *q_norm, const void *k_norm, cake_b200_block **out) {
  if (!c || !q || !k || !v || !o || !gate || !up || !down || !out) return fail(CAKE_B200_EINVAL, "null weight pointer");  // ln1 / ln2 may be NULL (OLMo2: no pre-norms)
  if (c->cfg.qkv_bias && (!q_bias || !k_bias || !v_bias)) return fail(CAKE_B200_EINVAL, "config has qkv_bias but a bias pointer is null");
  if (c->cfg.qk_norm && (!q_norm || !k_norm)) return fail(CAKE_B200_EINVAL, "config has qk_norm but a norm pointer is null");
  CU(cudaSetDevice(c->device));
  const size_t es = c->es, H = c->cfg.hidden, I = c->cfg.inter, hd = c->cfg.head_dim;
  const size_t sq = (size_t)c->cfg.n_heads * hd, skv = (size_t)c->cfg.n_kv_heads * hd;
  auto *b = new cake_b200_block{c, layer_idx, c->device};
  const int rc = [&]() -> int {  // on any failure (e.g. out of memory half way through a shard) nothing is leaked
  // attention.rs:109-113: Wqkv = cat([q,k,v], 0)
  CU(cudaMalloc(&b->wqkv, (sq + 2 * skv) * H * es + 16));
  RC(pipe_copy(c, b->wqkv, sq * H * es, q, sq * H * es, sq * H * es, 1));
  RC(pipe_copy(c, (char *)b->wqkv + sq * H * es, skv * H * es, k, skv * H * es, skv * H * es, 1));
  RC(pipe_copy(c, (char *)b->wqkv + (sq + skv) * H * es, skv * H * es, v, skv * H * es, skv * H * es, 1));
  RC(upload(c, &b->wo, o, H * sq * es));
  // mlp.rs:43-45 fuses gate|up by cat; here the fused matrix is ROW-INTERLEAVED (row 2i = gate_i,
  // row 2i+1 = up_i) so that silu(gate_i)*up_i is local to one warp's epilogue.  Numerically identical.
  CU(cudaMalloc(&b->wgu, 2 * I * H * es + 16));
  RC(pipe_copy(c, b->wgu, 2 * H * es, gate, H * es, H * es, I));               // the row interleave is done by the DMA engine
  RC(pipe_copy(c, (char *)b->wgu + H * es, 2 * H * es, up, H * es, H * es, I));
  RC(upload(c, &b->wd, down, H * I * es));
  if (ln1) RC(upload(c, &b->ln1, ln1, H * es));
  if (ln2) RC(upload(c, &b->ln2, ln2, H * es));
  if (c->cfg.qkv_bias) {
    CU(cudaMalloc(&b->bqkv, (sq + 2 * skv) * es + 16));
    RC(pipe_copy(c, b->bqkv, sq * es, q_bias, sq * es, sq * es, 1));
    RC(pipe_copy(c, (char *)b->bqkv + sq * es, skv * es, k_bias, skv * es, skv * es, 1));
    RC(pipe_copy(c, (char *)b->bqkv + (sq + skv) * es, skv * es, v_bias, skv * es, skv * es, 1));
  }
  if (c->cfg.qk_norm) {
    RC(upload(c, &b->qn, q_norm, (c->cfg.pre_reshape_qk_norm ? sq : hd) * es));   // attention.rs:121-122
    RC(upload(c, &b->kn, k_norm, (c->cfg.pre_reshape_qk_norm ? skv : hd) * es));
  }
  return pipe_mark(c);
  }();
  if (rc != CAKE_B200_OK) {
    cake_b200_block_free(b);
    return rc;
  }
  *out = b;
  return CAKE_B200_OK;
}
extern "C" void cake_b200_block_free(cake_b200_block *b) {
  if (!b) return;
  cudaSetDevice(b->device);
  void *bufs[] = {b->wqkv, b->wo, b->wgu, b->wd, b->ln1, b->ln2, b->bqkv, b->qn, b->kn, b->pan, b->pfn};
  for (void *p : bufs)
    if (p) cudaFree(p);  // cudaFree synchronises with outstanding work
  (void)cudaGetLastError();
  delete b;
}
extern "C" int cake_b200_block_layer(const cake_b200_block *b) { return b ? b->layer : -1; }
extern "C" int cake_b200_block_set_variant(cake_b200_block *b, const cake_b200_block_variant *v) {
  if (!b || !v) return fail(CAKE_B200_EINVAL, "null argument");
  if (v->sliding_window < -1) return fail(CAKE_B200_EINVAL, "sliding_window %d (use -1 for the config's, 0 for none)", v->sliding_window);
  cake_b200_ctx *c = b->ctx;
  CU(cudaSetDevice(c->device));
  const size_t bytes = (size_t)c->cfg.hidden * c->es;
  if (v->post_attention_norm) {
    if (b->pan) { cudaFree(b->pan); b->pan = nullptr; }
    RC(upload(c, &b->pan, v->post_attention_norm, bytes));
  }
  if (v->post_feedforward_norm) {
    if (b->pfn) { cudaFree(b->pfn); b->pfn = nullptr; }
    RC(upload(c, &b->pfn, v->post_feedforward_norm, bytes));
  }
  b->window = v->sliding_window;
  b->use_rope = v->use_rope != 0;
  return pipe_mark(c);
}

// ------------------------------------------------------------------------------------------ cache
extern "C" int cake_b200_cache_create(cake_b200_ctx *c, int batch, int max_seq, cake_b200_cache **out) {
  if (!c || !out || batch < 1 || max_seq < 1) return fail(CAKE_B200_EINVAL, "bad cache arguments");
  if (max_seq > c->cfg.max_seq) return fail(CAKE_B200_EINVAL, "cache max_seq %d exceeds config max_seq %d (RoPE table rows)", max_seq, c->cfg.max_seq);
  CU(cudaSetDevice(c->device));
  auto *k = new cake_b200_cache{c, batch, max_seq, c->device};
  k->gen = g_cache_gen.fetch_add(1);
  k->k.assign(c->cfg.n_layers, nullptr);
  k->v.assign(c->cfg.n_layers, nullptr);
  k->len.assign(c->cfg.n_layers, 0);
  cudaError_t e = cudaMalloc(&k->d_pos, 4);
  if (e == cudaSuccess) e = cudaMemset(k->d_pos, 0, 4);
  if (e != cudaSuccess) {
    if (k->d_pos) cudaFree(k->d_pos);
    delete k;
    (void)cudaGetLastError();
    return fail(CAKE_B200_ECUDA, "cache_create: %s", cudaGetErrorString(e));
  }
  *out = k;
  return CAKE_B200_OK;
}
static int cache_ensure(cake_b200_cache *k, int layer) {
  if (layer < 0 || layer >= (int)k->k.size()) return fail(CAKE_B200_EINVAL, "block_idx %d out of range", layer);
  if (k->k[layer]) return CAKE_B200_OK;
  const size_t bytes = (size_t)k->batch * k->ctx->cfg.n_kv_heads * k->cap * k->ctx->cfg.head_dim * k->ctx->es;
  CU(cudaMalloc(&k->k[layer], bytes + 16));
  CU(cudaMalloc(&k->v[layer], bytes + 16));
  return CAKE_B200_OK;
}
extern "C" int cake_b200_cache_clear(cake_b200_cache *k) {
  if (!k) return fail(CAKE_B200_EINVAL, "null argument");
  for (auto &l : k->len) l = 0;
  return CAKE_B200_OK;
}
extern "C" void cake_b200_cache_free(cake_b200_cache *k) {
  if (!k) return;
  cudaSetDevice(k->device);
  {  // drop everything in the ctx that points into this cache: the cached layer table and the decode graph
    std::lock_guard<std::mutex> lk(g_live_mu);
    if (g_live_ctx.count(k->ctx)) {
      cake_b200_ctx *c = k->ctx;
      cudaStreamSynchronize(c->stream);
      c->mk_sig.clear();
      if (c->g_cache == k) {
        if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
        if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
        c->g_cache = nullptr;
        c->g_block_idx.clear();
      }
    }
  }
  for (void *p : k->k)
    if (p) cudaFree(p);
  for (void *p : k->v)
    if (p) cudaFree(p);
  cudaFree(k->d_pos);
  (void)cudaGetLastError();
  delete k;
}
extern "C" int cake_b200_cache_len(const cake_b200_cache *k, int block_idx) {
  if (!k || block_idx < 0 || block_idx >= (int)k->len.size()) return -1;
  return k->len[block_idx];
}
extern "C" int cake_b200_cache_read(cake_b200_cache *k, int block_idx, int which, void *out_host, size_t bytes) {
  if (!k || !out_host) return fail(CAKE_B200_EINVAL, "null argument");
  cake_b200_ctx *c = k->ctx;
  CU(cudaSetDevice(c->device));
  if (block_idx < 0 || block_idx >= (int)k->len.size() || !k->k[block_idx]) return fail(CAKE_B200_EINVAL, "layer %d has no cache", block_idx);
  const int len = k->len[block_idx], hd = c->cfg.head_dim, nkv = c->cfg.n_kv_heads;
  const size_t need = (size_t)k->batch * nkv * len * hd * c->es;
  if (bytes < need) return fail(CAKE_B200_EINVAL, "cache_read needs %zu bytes", need);
  CU(cudaStreamSynchronize(c->stream));
  const char *src = (const char *)(which ? k->v[block_idx] : k->k[block_idx]);
  CU(cudaMemcpy2D(out_host, (size_t)len * hd * c->es, src, (size_t)k->cap * hd * c->es, (size_t)len * hd * c->es,
                  (size_t)k->batch * nkv, cudaMemcpyDeviceToHost));
  return CAKE_B200_OK;
}
extern "C" int cake_b200_cache_fill_synthetic(cake_b200_cache *k, const int *block_idx, int n_blocks, int len,
                                              uint32_t seed) {
  if (!k || (!block_idx && n_blocks > 0) || n_blocks < 0 || len < 0) return fail(CAKE_B200_EINVAL, "bad cache_fill_synthetic arguments");
  cake_b200_ctx *c = k->ctx;
  CU(cudaSetDevice(c->device));
  if (len > k->cap) return fail(CAKE_B200_EINVAL, "len %d > cache capacity %d", len, k->cap);
  for (int i = 0; i < n_blocks; i++) {
    const int l = block_idx[i];
    RC(cache_ensure(k, l));
    const size_t n = (size_t)k->batch * c->cfg.n_kv_heads * k->cap * c->cfg.head_dim;
    RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
      typedef typename decltype(tag_)::type T;
      fill_synth_kernel<T><<<1024, 256, 0, c->stream>>>((T *)k->k[l], n, seed + 2 * l, 1.0f);
      fill_synth_kernel<T><<<1024, 256, 0, c->stream>>>((T *)k->v[l], n, seed + 2 * l + 1, 1.0f);
      return CAKE_B200_OK;
    }));
    k->len[l] = len;
  }
  CU(cudaGetLastError());
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ decode path (batch 1, seq 1)
// x_in -> x_out through `n` blocks; position comes from cache->d_pos (device).  5 kernels per layer.
static int enqueue_qkv(cake_b200_ctx *c, const cake_b200_block *b, const void *x) {
  // rms_1 + fused qkv projection (+bias)                           transformer.rs:112, attention.rs:162-164
  GemvArgs a{};
  a.W = b->wqkv; a.x = x; a.norm_w = b->ln1; a.bias = b->bqkv; a.out = c->qkv; a.eps = c->cfg.rms_eps;
  a.N = c->nqkv; a.K = c->cfg.hidden;
  return launch_gemv<EPI_PLAIN>(c, a);
}
static int enqueue_attn(cake_b200_ctx *c, const cake_b200_block *b, cake_b200_cache *kc, int l) {
  // qk-norm, RoPE, KV append, GQA attention                        attention.rs:202-346, cache.rs:184-210
  const cake_b200_config &f = c->cfg;
  AttnDecodeArgs a{};
  a.qkv = c->qkv; a.kcache = kc->k[l]; a.vcache = kc->v[l]; a.cos_t = c->cos_t; a.sin_t = c->sin_t;
  a.q_norm = b->qn; a.k_norm = b->kn; a.y = c->y; a.ws_ml = c->ws_ml; a.ws_acc = c->ws_acc;
  a.counters = c->attn_counters; a.d_pos = kc->d_pos; a.n_heads = f.n_heads; a.n_kv = f.n_kv_heads;
  a.cap = kc->cap; a.rot = c->rot; a.nsplit = c->nsplit; a.eps = f.rms_eps;
  a.scale = (float)(1.0 / sqrt((double)f.head_dim));
  a.window = f.sliding_window;
  dim3 grid(c->nsplit, f.n_kv_heads), block(ATTN_THREADS);
  const size_t smem = attn_smem_bytes(f.head_dim, c->es);
  return DISPATCH_T(f.dtype, T_LAMBDA {
    typedef typename decltype(tag_)::type T;
    switch (f.head_dim) {
      case 16: return launch_pdl(c, attn_decode_kernel<T, 16>, grid, block, smem, a);
      case 32: return launch_pdl(c, attn_decode_kernel<T, 32>, grid, block, smem, a);
      case 64: return launch_pdl(c, attn_decode_kernel<T, 64>, grid, block, smem, a);
      case 128: return launch_pdl(c, attn_decode_kernel<T, 128>, grid, block, smem, a);
      default: return launch_pdl(c, attn_decode_kernel<T, 256>, grid, block, smem, a);
    }
  });
}
static int enqueue_oproj(cake_b200_ctx *c, const cake_b200_block *b, const void *residual, void *out) {
  // o_proj + residual                                              attention.rs:354, transformer.rs:123
  GemvArgs a{};
  a.W = b->wo; a.x = c->y; a.residual = residual; a.out = out; a.N = c->cfg.hidden;
  a.K = c->cfg.n_heads * c->cfg.head_dim;
  return launch_gemv<EPI_RESIDUAL>(c, a);
}
static int enqueue_gate_up(cake_b200_ctx *c, const cake_b200_block *b, const void *x1) {
  // rms_2 + gate_up + silu*mul                                     transformer.rs:129, mlp.rs:22-28
  GemvArgs a{};
  a.W = b->wgu; a.x = x1; a.norm_w = b->ln2; a.out = c->mm; a.eps = c->cfg.rms_eps; a.N = 2 * c->cfg.inter; a.act = c->cfg.use_gelu_mlp ? 1 : 0;
  a.K = c->cfg.hidden;
  return launch_gemv<EPI_SWIGLU>(c, a);
}
static int enqueue_down(cake_b200_ctx *c, const cake_b200_block *b, const void *residual, void *out) {
  // down + residual                                                mlp.rs:30, transformer.rs:131
  GemvArgs a{};
  a.W = b->wd; a.x = c->mm; a.residual = residual; a.out = out; a.N = c->cfg.hidden; a.K = c->cfg.inter;
  return launch_gemv<EPI_RESIDUAL>(c, a);
}

static int enqueue_decode_layers(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *block_idx, int n,
                                 cake_b200_cache *kc, const void *x_in, void *x_out) {
  const void *cur = x_in;
  for (int i = 0; i < n; i++) {
    const cake_b200_block *b = blocks[i];
    void *dst = (i == n - 1) ? x_out : c->xa;
    RC(enqueue_qkv(c, b, cur));
    RC(enqueue_attn(c, b, kc, block_idx[i]));
    RC(enqueue_oproj(c, b, cur, c->xb));
    RC(enqueue_gate_up(c, b, c->xb));
    RC(enqueue_down(c, b, c->xb, dst));
    cur = dst;
  }
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ megakernel plan + launch
static int plan_mk_geom(const cake_b200_ctx *c, int N, int K, int G, MkGeom *g, int *partial_floats, int *max_groups) {
  const int es = c->es;
  if (K % 64 != 0 || N % G != 0) return fail(CAKE_B200_EINVAL, "mega: K=%d must be a multiple of 64, N=%d of %d", K, N, G);
  int KC = K;
  while ((size_t)KC * es > MK_STAGE_BYTES && (KC / 2) % 64 == 0) KC /= 2;
  const size_t seg = (size_t)KC * es;
  if (seg > MK_STAGE_BYTES) return fail(CAKE_B200_EINVAL, "mega: K=%d cannot be staged", K);
  int RS = 1;
  while (RS < 64 && (size_t)(RS * 2) * seg <= MK_STAGE_BYTES) RS *= 2;
  int WPR, RPW;
  if (RS >= MK_CW) { WPR = 1; RPW = RS / MK_CW; }
  else { WPR = MK_CW / RS; RPW = 1; }
  while (WPR > 1 && ((KC / 8) % WPR != 0)) { WPR /= 2; }  // slices must divide the segment's 16-byte vectors
  if (RS * RPW < (MK_CW / WPR) * RPW || RS != (MK_CW / WPR) * RPW)  // every row slot must map to a staged row
    return fail(CAKE_B200_EINVAL, "mega: unsupported geometry N=%d K=%d (KC=%d RS=%d WPR=%d)", N, K, KC, RS, WPR);
  *g = MkGeom{N, K, KC, RS, WPR, RPW};
  // per-CTA cap of row groups in one phase: twice the even share (dynamic claiming, decode_mega.cuh mk_split)
  const int n_groups = (N + RS - 1) / RS;
  const int cap = 2 * ((n_groups + c->sm_count - 1) / c->sm_count) + 4;
  if (cap > *max_groups) *max_groups = cap;
  // *partial_floats collects max(RS*WPR) here and the static-path need; plan_mega multiplies by the global cap
  if (RS * WPR > *partial_floats % 65536) *partial_floats = (*partial_floats / 65536) * 65536 + RS * WPR;
  const int pf_static = (((N / G) / c->sm_count) + 1) * G * WPR;
  if (pf_static > *partial_floats / 65536) *partial_floats = pf_static * 65536 + *partial_floats % 65536;
  return CAKE_B200_OK;
}

struct MkPlan {
  MkArgs a;
  size_t smem;
};
static bool mega_supported(const cake_b200_ctx *c) {
  const int G = c->cfg.n_heads / c->cfg.n_kv_heads;
#define MK_CASE(HD_, G_) if (c->cfg.head_dim == HD_ && G == G_) return true;
  MK_FOR_ALL(MK_CASE)
#undef MK_CASE
  return false;
}
static int plan_mega(cake_b200_ctx *c, cake_b200_cache *kc, bool with_head, MkPlan *p) {
  const cake_b200_config &f = c->cfg;
  MkArgs &a = p->a;
  memset(&a, 0, sizeof(a));
  int pf = 0, mg = 0;
  RC(plan_mk_geom(c, c->nqkv, f.hidden, 1, &a.g_qkv, &pf, &mg));
  RC(plan_mk_geom(c, f.hidden, f.n_heads * f.head_dim, 1, &a.g_o, &pf, &mg));
  RC(plan_mk_geom(c, 2 * f.inter, f.hidden, 2, &a.g_gu, &pf, &mg));
  RC(plan_mk_geom(c, f.hidden, f.inter, 1, &a.g_down, &pf, &mg));
  if (with_head) RC(plan_mk_geom(c, f.vocab, f.hidden, 1, &a.g_head, &pf, &mg));
  {  // partial-sum scratch: every phase may use up to `mg` groups of its own RS*WPR floats (dynamic), or its static rows
    const int dyn_need = mg * (pf % 65536), static_need = pf / 65536;
    pf = dyn_need > static_need ? dyn_need : static_need;
  }
  a.partial_floats = pf;
  a.max_groups = mg;
  a.tickets = c->tickets;
  {
    const char *e = getenv("CAKE_B200_L2_PREFETCH");
    a.l2_prefetch = (e && e[0] >= '1' && e[0] <= '2') ? e[0] - '0' : 0;
  }
  a.hidden = f.hidden; a.inter = f.inter; a.n_heads = f.n_heads; a.n_kv = f.n_kv_heads; a.hd = f.head_dim;
  a.rot = c->rot; a.cap = kc ? kc->cap : 0; a.nsplit = c->nsplit; a.eps = f.rms_eps;
  a.scale = (float)(1.0 / sqrt((double)f.head_dim));
  a.max_k = f.inter > f.hidden ? f.inter : f.hidden;
  if (f.n_heads * f.head_dim > a.max_k) a.max_k = f.n_heads * f.head_dim;
  a.xa = c->xa; a.xb = c->xb; a.qkv = c->qkv; a.y = c->y; a.mm = c->mm;
  a.ws_ml = c->ws_ml; a.ws_acc = c->ws_acc; a.attn_counters = c->attn_counters;
  a.cos_t = c->cos_t; a.sin_t = c->sin_t; a.d_pos = kc ? kc->d_pos : nullptr; a.d_step = c->d_step; a.gbar = c->gbar;
  a.vocab = f.vocab; a.embed = c->embed; a.d_token = c->d_token;
  a.ln_f = c->ln_f; a.lm_head = c->lm_head; a.logits = c->logits; a.part_val = c->part_val; a.part_idx = c->part_idx;
  a.argmax_counter = c->argmax_counter; a.token_out = c->d_token; a.token_ring = nullptr; a.ring_cap = TOKEN_RING;
  a.trace = c->trace;
  a.step_trace = c->step_trace;
  a.trace_tag = 0;
  a.act = f.use_gelu_mlp ? 1 : 0;
  a.window = f.sliding_window;
  int ns = MK_MAX_STAGES;
  const size_t limit = 227 * 1024 - 4096;  // the kernel also has ~2.2 KB of static shared memory
  while (ns > 2 && mk_smem_bytes(a.max_k, pf, mg, ns, c->es) > limit) ns--;
  a.n_stages = ns;
  p->smem = mk_smem_bytes(a.max_k, pf, mg, ns, c->es);
  if (p->smem > limit) return fail(CAKE_B200_EINVAL, "mega: needs %zu B of shared memory", p->smem);
  return CAKE_B200_OK;
}

template <typename T> static int launch_mega_T(cake_b200_ctx *c, cudaLaunchConfig_t *cfg, const MkArgs &a) {
  const int G = c->cfg.n_heads / c->cfg.n_kv_heads;
#define MK_CASE(HD_, G_)                                                \
  if (c->cfg.head_dim == HD_ && G == G_) {                              \
    CU(cudaLaunchKernelEx(cfg, decode_mega_kernel<T, HD_, G_>, a));     \
    return CAKE_B200_OK;                                                \
  }
  MK_FOR_ALL(MK_CASE)
#undef MK_CASE
  return fail(CAKE_B200_EINVAL, "megakernel: head_dim %d with %d query heads per kv head is not instantiated", c->cfg.head_dim, G);
}

static int launch_mega(cake_b200_ctx *c, const MkPlan &p) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(c->sm_count);
  cfg.blockDim = dim3(MK_THREADS);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;  // all CTAs must be co-resident (grid barriers inside)
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  const MkArgs &a = p.a;
  CU(cudaMemsetAsync(c->tickets, 0, sizeof(unsigned) * (4 * (size_t)a.n_layers + 4), c->stream));  // a memset node in the graph
  int rc = DISPATCH_T(c->cfg.dtype, T_LAMBDA {
    typedef typename decltype(tag_)::type T;
    return launch_mega_T<T>(c, &cfg, a);
  });
  RC(rc);
  if (c->capturing) c->g_kernels++;
  else c->launches++;
  return CAKE_B200_OK;
}

static void fill_mk_table(MkLayer *tab, cake_b200_block *const *blocks, const int *block_idx, int n, cake_b200_cache *kc) {
  for (int i = 0; i < n; i++) {
    const cake_b200_block *b = blocks[i];
    tab[i] = MkLayer{b->wqkv, b->wo, b->wgu, b->wd, b->ln1, b->ln2, b->bqkv, b->qn, b->kn, kc->k[block_idx[i]], kc->v[block_idx[i]]};
  }
}

// x_in -> x_out through n blocks in ONE persistent kernel; position from cache->d_pos
static int enqueue_mega_layers(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *block_idx, int n,
                               cake_b200_cache *kc, const void *x_in, void *x_out) {
  if (n > MK_MAX_LAYERS) return fail(CAKE_B200_EINVAL, "more than %d blocks in one call", MK_MAX_LAYERS);
  std::vector<const void *> sig;
  sig.push_back(kc);
  sig.push_back(reinterpret_cast<const void *>((uintptr_t)kc->gen));
  for (int i = 0; i < n; i++) { sig.push_back(blocks[i]); sig.push_back(kc->k[block_idx[i]]); sig.push_back(kc->v[block_idx[i]]); }
  if (sig != c->mk_sig) {  // (re)build the layer table; rare
    CU(cudaStreamSynchronize(c->stream));
    fill_mk_table(c->mk_tab_host, blocks, block_idx, n, kc);
    CU(cudaMemcpyAsync(c->mk_tab_dev, c->mk_tab_host, sizeof(MkLayer) * n, cudaMemcpyHostToDevice, c->stream));
    c->mk_sig = sig;
  }
  MkPlan p;
  RC(plan_mega(c, kc, false, &p));
  p.a.layers = c->mk_tab_dev;
  p.a.n_layers = n;
  p.a.x_in = x_in;
  p.a.x_out = x_out;
  p.a.has_head = 0;
  p.a.advance = 0;
  return launch_mega(c, p);
}

// ln_f + lm_head + greedy argmax on one hidden row -> logits_out (nullable), token -> c->d_token
static int enqueue_head(cake_b200_ctx *c, const void *x_row, void *logits_out, bool ring) {
  GemvArgs a{};
  a.W = c->lm_head; a.x = x_row; a.norm_w = c->ln_f; a.out = logits_out; a.eps = c->cfg.rms_eps;
  a.N = c->cfg.vocab; a.K = c->cfg.hidden;
  a.part_val = c->part_val; a.part_idx = c->part_idx; a.counter = c->argmax_counter; a.token_out = c->d_token;
  a.token_ring = ring ? c->token_ring : nullptr; a.step = c->d_step; a.ring_cap = TOKEN_RING;
  return launch_gemv<EPI_ARGMAX>(c, a);
}

// ------------------------------------------------------------------------------------------ TMA tensor maps (driver entry point, no -lcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode_tiled = nullptr;
static int tmap_init() {
  if (g_encode_tiled) return CAKE_B200_OK;
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(CAKE_B200_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
  g_encode_tiled = (EncodeTiledFn)fn;
  return CAKE_B200_OK;
}
// [rows, K] row-major matrix of D -> 2-D map with a (64 x box_rows) box and 128-byte swizzle
static int make_tmap(CUtensorMap *m, const void *ptr, uint64_t rows, uint64_t K, int dtype, uint32_t box_rows) {
  RC(tmap_init());
  const cuuint64_t gdim[2] = {K, rows};
  const cuuint64_t gstride[1] = {K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = g_encode_tiled(m, dtype == CAKE_B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                                    const_cast<void *>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CAKE_B200_ECUDA, "cuTensorMapEncodeTiled(rows=%llu, K=%llu) -> CUresult %d", (unsigned long long)rows, (unsigned long long)K, (int)r);
  return CAKE_B200_OK;
}

// [slabs][rows][cols] tensor of D with a slab stride of slab_rows rows -> 3-D map, (64 x box_rows x 1) box, 128-byte swizzle.
// `rows` is the VISIBLE extent: the TMA unit zero-fills rows past it (query rows >= S, keys >= the attended length), so a
// tile never carries a neighbouring slab's rows or uninitialised cache memory into an MMA.
static int make_tmap3(CUtensorMap *m, const void *ptr, uint64_t cols, uint64_t row_stride_elems, uint64_t rows, uint64_t slab_rows,
                      uint64_t slabs, int dtype, uint32_t box_rows) {
  RC(tmap_init());
  const cuuint64_t gdim[3] = {cols, rows, slabs};
  const cuuint64_t gstride[2] = {row_stride_elems * 2, slab_rows * row_stride_elems * 2};
  const cuuint32_t box[3] = {64u, box_rows, 1u};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = g_encode_tiled(m, dtype == CAKE_B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3,
                                    const_cast<void *>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CAKE_B200_ECUDA, "cuTensorMapEncodeTiled(3d: cols=%llu rows=%llu slabs=%llu) -> CUresult %d", (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)slabs, (int)r);
  return CAKE_B200_OK;
}

static bool tc_gemm_ok(int M, int N, int K) { return M >= 1 && N % TC_BN == 0 && K % TC_BK == 0; }

// C = epi(A[M,K] W[N,K]^T) on the tensor cores (gemm_tc.cuh)
template <typename T, int EPI>
static int gemm_tc_T(cake_b200_ctx *c, const void *A, const void *W, const void *bias, const void *res, void *C, int M, int N, int K) {
  // 128 x 256 tiles when N allows and there are enough tiles to fill the chip; CAKE_B200_TC_BN=128 forces the small tile (A/B aid)
  static const int force_bn = []() { const char *e = getenv("CAKE_B200_TC_BN"); return e ? atoi(e) : 0; }();
  const int tiles_m = (M + TC_BM - 1) / TC_BM;
  const bool big = force_bn ? force_bn == 256 && N % 256 == 0 : (N % 256 == 0 && tiles_m * (N / 256) >= c->sm_count);
  const int bn = big ? 256 : 128;
  // pairs of CTAs sharing the W tile by TMA multicast (gemm_tc.cuh MC): needs an even number of 128-row tiles
  static const int no_mc = []() { const char *e = getenv("CAKE_B200_TC_MC"); return (e && e[0] == '0') ? 1 : 0; }();
  const bool mc = big && !no_mc && M % 256 == 0 && (tiles_m / 2) * (N / 256) >= c->sm_count / 2;
  CUtensorMap ma, mb;
  RC(make_tmap(&ma, A, (uint64_t)M, (uint64_t)K, c->cfg.dtype, TC_BM));
  RC(make_tmap(&mb, W, (uint64_t)N, (uint64_t)K, c->cfg.dtype, mc ? 128u : (uint32_t)bn));
  TcParams p{bias, res, C, M, N, K, c->cfg.use_gelu_mlp ? 1 : 0};
  const int tiles = tiles_m * (N / bn);
  dim3 grid(tiles < c->sm_count ? tiles : c->sm_count), block(TC_THREADS);
  if (mc) {
    dim3 g2((unsigned)(c->sm_count & ~1));
    return launch_cluster_pdl(c, gemm_tc_kernel<T, EPI, 256, true>, g2, block, (size_t)TcCfg<256>::SMEM_BYTES, 2u, ma, mb, p);
  }
  if (big) return launch_pdl(c, gemm_tc_kernel<T, EPI, 256>, grid, block, (size_t)TcCfg<256>::SMEM_BYTES, ma, mb, p);
  return launch_pdl(c, gemm_tc_kernel<T, EPI, 128>, grid, block, (size_t)TcCfg<128>::SMEM_BYTES, ma, mb, p);
}

// ------------------------------------------------------------------------------------------ prefill path (any batch / seq)
static int pf_reserve(cake_b200_ctx *c, size_t rows) {
  if (rows <= c->pf_rows) return CAKE_B200_OK;
  CU(cudaStreamSynchronize(c->stream));
  void **bufs[] = {&c->pf_h, &c->pf_qkv, &c->pf_y, &c->pf_x1, &c->pf_gu, &c->pf_mm};
  for (void **b : bufs)
    if (*b) { cudaFree(*b); *b = nullptr; }
  const cake_b200_config &f = c->cfg;
  const size_t es = c->es;
  CU(cudaMalloc(&c->pf_h, rows * f.hidden * es));
  CU(cudaMalloc(&c->pf_qkv, rows * c->nqkv * es));
  CU(cudaMalloc(&c->pf_y, rows * f.n_heads * f.head_dim * es));
  CU(cudaMalloc(&c->pf_x1, rows * f.hidden * es));
  CU(cudaMalloc(&c->pf_gu, rows * 2 * f.inter * es));
  CU(cudaMalloc(&c->pf_mm, rows * f.inter * es));
  c->pf_rows = rows;
  return CAKE_B200_OK;
}

// RMSNorm over M rows of n: warp-per-row vector kernel for n = 256 * {4, 8, 16, 32} and 16-byte aligned rows, else the generic one
template <typename T>
static int rmsnorm_rows_T(cake_b200_ctx *c, const void *x, const void *w, void *out, int M, int n, float eps) {
  const bool al = (((uintptr_t)x | (uintptr_t)w | (uintptr_t)out) & 15) == 0;
  const dim3 g((unsigned)((M + 7) / 8)), blk(256);
  if (al && n == 1024) return launch_pdl(c, rmsnorm_rows_vec_kernel<T, 4>, g, blk, 0, (const T *)x, (const T *)w, (T *)out, M, eps);
  if (al && n == 2048) return launch_pdl(c, rmsnorm_rows_vec_kernel<T, 8>, g, blk, 0, (const T *)x, (const T *)w, (T *)out, M, eps);
  if (al && n == 4096) return launch_pdl(c, rmsnorm_rows_vec_kernel<T, 16>, g, blk, 0, (const T *)x, (const T *)w, (T *)out, M, eps);
  if (al && n == 8192) return launch_pdl(c, rmsnorm_rows_vec_kernel<T, 32>, g, blk, 0, (const T *)x, (const T *)w, (T *)out, M, eps);
  return launch_pdl(c, rmsnorm_rows_kernel<T>, dim3(M), dim3(256), 0, (const T *)x, (const T *)w, (T *)out, n, eps);
}

template <typename T>
static int gemm_T(cake_b200_ctx *c, const void *A, const void *W, const void *bias, const void *res, void *C, int M,
                  int N, int K) {
  dim3 grid((N + GV0_BN - 1) / GV0_BN, (M + GV0_BM - 1) / GV0_BM), block(256);
  return launch_pdl(c, gemm_v0_kernel<T>, grid, block, 0, (const T *)A, (const T *)W, (const T *)bias, (const T *)res,
                    (T *)C, M, N, K);
}

static int enqueue_prefill_layers(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *block_idx, int n,
                                  cake_b200_cache *kc, const void *x_in, void *x_out, int B, int S, int pos0) {
  const cake_b200_config &f = c->cfg;
  const int H = f.hidden, I = f.inter, hd = f.head_dim, sq = f.n_heads * hd, M = B * S;
  RC(pf_reserve(c, (size_t)M));
  // cache.rs:173-205: a call on a non-empty cache sees the last `window` positions of cat(cache, new) — every query of
  // the chunk the same old rows [ws, pos0) — while the first call (pos0 == 0) attends over everything it stores.  The
  // attention kernels only see row indices relative to ws (base pointers advanced by ws rows); K/V append is unaffected.
  // per block (attention.rs load_custom gives every layer its own window): block_window() below
  auto block_window = [&](const cake_b200_block *b, int *ws_out) -> int {
    const int w = b->window >= 0 ? b->window : f.sliding_window;
    int ws = 0;
    if (w > 0 && pos0 > 0 && pos0 + S > w) {
      ws = pos0 + S - w;
      if (ws > pos0) return fail(CAKE_B200_EINVAL, "a chunk of %d tokens on a non-empty cache exceeds the sliding window %d", S, w);
    }
    *ws_out = ws;
    return CAKE_B200_OK;
  };
  for (int i = 0; i < n; i++) { int t; RC(block_window(blocks[i], &t)); }  // refuse before anything is enqueued
  const char *env_tc = getenv("CAKE_B200_NO_TC");
  const bool use_tc = !(env_tc && env_tc[0] == '1');  // CAKE_B200_NO_TC=1: CUDA-core GEMM (A/B testing aid)
  return DISPATCH_T(f.dtype, T_LAMBDA {
      typedef typename decltype(tag_)::type T;
    const void *cur = x_in;
    for (int i = 0; i < n; i++) {
      const cake_b200_block *b = blocks[i];
      const int l = block_idx[i];
      int ws = 0;
      RC(block_window(b, &ws));
      const size_t wsoff = (size_t)ws * hd;
      const int apos0 = pos0 - ws;
      // transformer.rs:112 rms_1 — or none: olmo2/block.rs:70 feeds x itself to the attention
      const void *att_in = cur;
      if (b->ln1) { RC(rmsnorm_rows_T<T>(c, cur, b->ln1, c->pf_h, M, H, f.rms_eps)); att_in = c->pf_h; }
      if (use_tc && tc_gemm_ok(M, c->nqkv, H) && ((uintptr_t)att_in & 15) == 0) RC((gemm_tc_T<T, TCE_PLAIN>(c, att_in, b->wqkv, b->bqkv, nullptr, c->pf_qkv, M, c->nqkv, H)));
      else RC(gemm_T<T>(c, att_in, b->wqkv, b->bqkv, nullptr, c->pf_qkv, M, c->nqkv, H));
      const void *qn = b->qn, *kn = b->kn;
      if (f.pre_reshape_qk_norm && qn && kn) {  // attention.rs:176-192: RmsNorm over the whole q and k projections, in place
        RC(launch_pdl(c, rmsnorm_strided_kernel<T>, dim3(M), dim3(256), 0, (T *)c->pf_qkv, (const T *)qn, c->nqkv, sq, f.rms_eps));
        RC(launch_pdl(c, rmsnorm_strided_kernel<T>, dim3(M), dim3(256), 0, (T *)c->pf_qkv + sq, (const T *)kn, c->nqkv, (int)(f.n_kv_heads * hd), f.rms_eps));
        qn = kn = nullptr;
      }
      const int rope_on = b->use_rope ? 1 : 0;
      {
        const long items = (long)M * (f.n_heads + 2 * f.n_kv_heads);
        if (hd == 128 && c->rot == 128 && ((uintptr_t)c->cos_t & 3) == 0)  // full rotary on 128-wide heads: token-per-warp vector kernel
          RC(launch_pdl(c, rope_append_vec_kernel<T>, dim3((unsigned)((M + 3) / 4)), dim3(128), 0, (T *)c->pf_qkv, (T *)kc->k[l], (T *)kc->v[l],
                        (const T *)c->cos_t, (const T *)c->sin_t, (const T *)qn, (const T *)kn, B, S, f.n_heads, f.n_kv_heads, kc->cap, pos0,
                        f.rms_eps, rope_on));
        else
        RC(launch_pdl(c, rope_append_kernel<T>, dim3((unsigned)((items + 3) / 4)), dim3(128), (size_t)4 * hd * 4,
                      (T *)c->pf_qkv, (T *)kc->k[l], (T *)kc->v[l], (const T *)c->cos_t, (const T *)c->sin_t,
                      (const T *)qn, (const T *)kn, B, S, f.n_heads, f.n_kv_heads, hd, c->rot, kc->cap, pos0, f.rms_eps, rope_on));
      }
      if (use_tc && (hd == 64 || hd == 128)) {  // tensor-core flash attention (attn_prefill.cuh)
        dim3 grid((S + FA_BM - 1) / FA_BM, f.n_heads, B), block(FA_THREADS);
        const float sc = (float)(1.0 / sqrt((double)hd));
        // head_dim 128: tcgen05 flash attention (attn_prefill_tc.cuh); CAKE_B200_FA=mma keeps the mma.sync kernel (A/B aid)
        static const bool fa_mma = []() { const char *e = getenv("CAKE_B200_FA"); return e && !strcmp(e, "mma"); }();
        static const float fa_tau = []() { const char *e = getenv("CAKE_B200_FA_TAU"); return e ? (float)atof(e) : 5.545f; }();
        if (hd == 128 && !fa_mma) {
          CUtensorMap mq, mk, mv;
          const uint64_t nqkv_cols = (uint64_t)(f.n_heads + 2 * f.n_kv_heads) * hd;
          RC(make_tmap3(&mq, c->pf_qkv, nqkv_cols, nqkv_cols, (uint64_t)S, (uint64_t)S, (uint64_t)B, f.dtype, FT_BM));
          RC(make_tmap3(&mk, (const T *)kc->k[l] + wsoff, (uint64_t)hd, (uint64_t)hd, (uint64_t)(apos0 + S), (uint64_t)kc->cap,
                        (uint64_t)B * f.n_kv_heads, f.dtype, FT_BN));
          RC(make_tmap3(&mv, (const T *)kc->v[l] + wsoff, (uint64_t)hd, (uint64_t)hd, (uint64_t)(apos0 + S), (uint64_t)kc->cap,
                        (uint64_t)B * f.n_kv_heads, f.dtype, FT_BN));
          dim3 gt((S + FT_BM - 1) / FT_BM, f.n_heads, B);
          RC(launch_pdl(c, attn_prefill_tc_kernel<T>, gt, dim3(FT_THREADS), (size_t)FT_SMEM_BYTES, mq, mk, mv, (T *)c->pf_y, S, f.n_heads,
                        f.n_kv_heads, apos0, sc * 1.4426950408889634f, fa_tau * 1.4426950408889634f));
        } else if (hd == 128)
          RC(launch_pdl(c, attn_prefill_mma_kernel<T, 128>, grid, block, (size_t)4 * FA_BN * 128 * 2, (const T *)c->pf_qkv,
                        (const T *)kc->k[l] + wsoff, (const T *)kc->v[l] + wsoff, (T *)c->pf_y, S, f.n_heads, f.n_kv_heads, kc->cap, apos0, sc));
        else
          RC(launch_pdl(c, attn_prefill_mma_kernel<T, 64>, grid, block, (size_t)4 * FA_BN * 64 * 2, (const T *)c->pf_qkv,
                        (const T *)kc->k[l] + wsoff, (const T *)kc->v[l] + wsoff, (T *)c->pf_y, S, f.n_heads, f.n_kv_heads, kc->cap, apos0, sc));
      } else {
        const long items = (long)M * f.n_heads;
        RC(launch_pdl(c, attn_prefill_v0_kernel<T>, dim3((unsigned)((items + 3) / 4)), dim3(128), 0, (const T *)c->pf_qkv,
                      (const T *)kc->k[l] + wsoff, (const T *)kc->v[l] + wsoff, (T *)c->pf_y, B, S, f.n_heads, f.n_kv_heads, hd, kc->cap,
                      apos0, (float)(1.0 / sqrt((double)hd))));
      }
      if (b->pan) {  // olmo2/block.rs:77-79, gemma3/block.rs:120-122: x1 = x + rms_norm(o_proj(y))  (pf_h is free again here)
        if (use_tc && tc_gemm_ok(M, H, sq)) RC((gemm_tc_T<T, TCE_PLAIN>(c, c->pf_y, b->wo, nullptr, nullptr, c->pf_h, M, H, sq)));
        else RC(gemm_T<T>(c, c->pf_y, b->wo, nullptr, nullptr, c->pf_h, M, H, sq));
        RC(launch_pdl(c, rmsnorm_residual_rows_kernel<T>, dim3(M), dim3(256), 0, (const T *)c->pf_h, (const T *)b->pan, (const T *)cur, (T *)c->pf_x1, H, f.rms_eps));
      } else if (use_tc && tc_gemm_ok(M, H, sq) && ((uintptr_t)cur & 15) == 0) RC((gemm_tc_T<T, TCE_RESIDUAL>(c, c->pf_y, b->wo, nullptr, cur, c->pf_x1, M, H, sq)));
      else RC(gemm_T<T>(c, c->pf_y, b->wo, nullptr, cur, c->pf_x1, M, H, sq));
      const void *mlp_in = c->pf_x1;
      if (b->ln2) { RC(rmsnorm_rows_T<T>(c, c->pf_x1, b->ln2, c->pf_h, M, H, f.rms_eps)); mlp_in = c->pf_h; }
      if (use_tc && tc_gemm_ok(M, 2 * I, H)) {
        RC((gemm_tc_T<T, TCE_SWIGLU>(c, mlp_in, b->wgu, nullptr, nullptr, c->pf_mm, M, 2 * I, H)));
      } else {
        RC(gemm_T<T>(c, mlp_in, b->wgu, nullptr, nullptr, c->pf_gu, M, 2 * I, H));
        RC(launch_pdl(c, swiglu_rows_kernel<T>, dim3(c->sm_count * 4), dim3(256), 0, (const T *)c->pf_gu, (T *)c->pf_mm, (size_t)M * I, f.use_gelu_mlp ? 1 : 0));
      }
      if (b->pfn) {  // olmo2/block.rs:84-86, gemma3/block.rs:130-132: out = x1 + rms_norm(down(...))
        if (use_tc && tc_gemm_ok(M, H, I)) RC((gemm_tc_T<T, TCE_PLAIN>(c, c->pf_mm, b->wd, nullptr, nullptr, c->pf_h, M, H, I)));
        else RC(gemm_T<T>(c, c->pf_mm, b->wd, nullptr, nullptr, c->pf_h, M, H, I));
        RC(launch_pdl(c, rmsnorm_residual_rows_kernel<T>, dim3(M), dim3(256), 0, (const T *)c->pf_h, (const T *)b->pfn, (const T *)c->pf_x1, (T *)x_out, H, f.rms_eps));
      } else if (use_tc && tc_gemm_ok(M, H, I)) RC((gemm_tc_T<T, TCE_RESIDUAL>(c, c->pf_mm, b->wd, nullptr, c->pf_x1, x_out, M, H, I)));
      else RC(gemm_T<T>(c, c->pf_mm, b->wd, nullptr, c->pf_x1, x_out, M, H, I));
      cur = x_out;
    }
    return CAKE_B200_OK;
  });
}

// ------------------------------------------------------------------------------------------ forward_batch
extern "C" int cake_b200_forward_batch(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *block_idx,
                                       int n_blocks, cake_b200_cache *kc, const void *x_dev, void *y_dev, int batch,
                                       int seq, int index_pos) {
  if (!c || !blocks || !block_idx || !kc || !x_dev || !y_dev || n_blocks < 1) return fail(CAKE_B200_EINVAL, "null/empty argument");
  if (batch != kc->batch) return fail(CAKE_B200_EINVAL, "batch %d != cache batch %d", batch, kc->batch);
  if (seq < 1 || index_pos < 0 || (long long)index_pos + (long long)seq > (long long)kc->cap)
    return fail(CAKE_B200_ESTATE, "index_pos %d + seq %d exceeds cache capacity %d", index_pos, seq, kc->cap);
  CU(cudaSetDevice(c->device));
  RC(wait_loads(c));
  for (int i = 0; i < n_blocks; i++) {
    RC(cache_ensure(kc, block_idx[i]));
    if (kc->len[block_idx[i]] != index_pos)
      return fail(CAKE_B200_ESTATE, "block %d: index_pos %d != cache length %d (clear the cache or feed positions in order)",
                  block_idx[i], index_pos, kc->len[block_idx[i]]);
  }
  bool variant = false;  // sibling block structures step through the batched path (cake_b200_block_set_variant)
  for (int i = 0; i < n_blocks; i++) {
    if (!blocks[i]) return fail(CAKE_B200_EINVAL, "blocks[%d] is null", i);
    variant = variant || blocks[i]->variant();
  }
  if (batch == 1 && seq == 1 && !variant) {
    set_int_kernel<<<1, 1, 0, c->stream>>>(kc->d_pos, index_pos);
    c->launches++;
    if (c->use_mega && mega_supported(c)) RC(enqueue_mega_layers(c, blocks, block_idx, n_blocks, kc, x_dev, y_dev));
    else RC(enqueue_decode_layers(c, blocks, block_idx, n_blocks, kc, x_dev, y_dev));
  } else {
    RC(enqueue_prefill_layers(c, blocks, block_idx, n_blocks, kc, x_dev, y_dev, batch, seq, index_pos));
  }
  for (int i = 0; i < n_blocks; i++) kc->len[block_idx[i]] = index_pos + seq;
  CU(cudaGetLastError());
  return CAKE_B200_OK;
}

static int io_reserve(cake_b200_ctx *c, size_t bytes) {
  if (bytes > c->io_x_cap) {
    CU(cudaStreamSynchronize(c->stream));
    if (c->io_x) cudaFree(c->io_x);
    CU(cudaMalloc(&c->io_x, bytes));
    c->io_x_cap = bytes;
  }
  if (bytes > c->h_pin_x_cap) {
    if (c->h_pin_x) cudaFreeHost(c->h_pin_x);
    CU(cudaMallocHost(&c->h_pin_x, bytes));
    c->h_pin_x_cap = bytes;
  }
  return CAKE_B200_OK;
}

extern "C" int cake_b200_forward_batch_host(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *block_idx,
                                            int n_blocks, cake_b200_cache *kc, const void *x_host, void *y_host,
                                            int batch, int seq, int index_pos) {
  if (!c || !x_host || !y_host || !kc) return fail(CAKE_B200_EINVAL, "null argument");
  // validated BEFORE any size is derived from them: the dims may come off the wire (cake_worker)
  if (batch < 1 || seq < 1 || batch != kc->batch) return fail(CAKE_B200_EINVAL, "batch %d (cache batch %d) / seq %d out of range", batch, kc->batch, seq);
  if (index_pos < 0 || (long long)index_pos + (long long)seq > (long long)kc->cap)
    return fail(CAKE_B200_ESTATE, "index_pos %d + seq %d exceeds cache capacity %d", index_pos, seq, kc->cap);
  CU(cudaSetDevice(c->device));
  const size_t bytes = (size_t)batch * seq * c->cfg.hidden * c->es;
  RC(io_reserve(c, bytes));
  memcpy(c->h_pin_x, x_host, bytes);
  CU(cudaMemcpyAsync(c->io_x, c->h_pin_x, bytes, cudaMemcpyHostToDevice, c->stream));
  RC(cake_b200_forward_batch(c, blocks, block_idx, n_blocks, kc, c->io_x, c->io_x, batch, seq, index_pos));
  CU(cudaMemcpyAsync(c->h_pin_x, c->io_x, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  memcpy(y_host, c->h_pin_x, bytes);
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ head / tail
extern "C" int cake_b200_head_load(cake_b200_ctx *c, const void *embed, const void *ln_f, const void *lm_head) {
  if (!c || !embed || !ln_f) return fail(CAKE_B200_EINVAL, "null argument");
  if (!lm_head && !c->cfg.tie_embeddings) return fail(CAKE_B200_EINVAL, "lm_head is NULL but tie_embeddings is 0");
  CU(cudaSetDevice(c->device));
  const size_t es = c->es, H = c->cfg.hidden, V = c->cfg.vocab;
  RC(upload(c, &c->embed, embed, V * H * es));
  RC(upload(c, &c->ln_f, ln_f, H * es));
  if (lm_head) RC(upload(c, &c->lm_head, lm_head, V * H * es));
  else c->lm_head = c->embed;  // text_model.rs:164-167
  return pipe_mark(c);
}

extern "C" int cake_b200_embed(cake_b200_ctx *c, const uint32_t *ids_host, int batch, int seq, void *x_dev) {
  if (!c || !ids_host || !x_dev || !c->embed) return fail(CAKE_B200_EINVAL, "null argument or head not loaded");
  CU(cudaSetDevice(c->device));
  RC(wait_loads(c));
  if (batch < 1 || seq < 1) return fail(CAKE_B200_EINVAL, "embed: batch and seq must be >= 1");
  const size_t n = (size_t)batch * seq;
  for (size_t i = 0; i < n; i++)  // index_select on an out-of-range id is an error in the reference (backends/mod.rs:513-528)
    if (ids_host[i] >= (uint32_t)c->cfg.vocab)
      return fail(CAKE_B200_EINVAL, "token id %u at %zu out of range (vocab %d)", ids_host[i], i, c->cfg.vocab);
  if (n > c->d_ids_cap) {
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_ids) cudaFree(c->d_ids);
    CU(cudaMalloc(&c->d_ids, n * 4));
    c->d_ids_cap = n;
  }
  CU(cudaMemcpyAsync(c->d_ids, ids_host, n * 4, cudaMemcpyHostToDevice, c->stream));
  // the ids buffer is pageable in general: the copy above returns after staging, safe to reuse
  RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
      typedef typename decltype(tag_)::type T;
    return launch_pdl(c, embed_kernel<T>, dim3((unsigned)n), dim3(128), 0, (const T *)c->embed, (const uint32_t *)c->d_ids,
                      (T *)x_dev, (int)n, c->cfg.hidden, c->cfg.vocab, c->cfg.embed_scale);
  }));
  return CAKE_B200_OK;
}

extern "C" int cake_b200_logits(cake_b200_ctx *c, const void *x_dev, int batch, int seq, void *logits_dev,
                                uint32_t *argmax_host) {
  if (!c || !x_dev || !c->lm_head) return fail(CAKE_B200_EINVAL, "null argument or head not loaded");
  if (batch > 64) return fail(CAKE_B200_EINVAL, "batch > 64");
  CU(cudaSetDevice(c->device));
  RC(wait_loads(c));
  const size_t H = c->cfg.hidden, V = c->cfg.vocab, es = c->es;
  for (int b = 0; b < batch; b++) {
    const char *row = (const char *)x_dev + ((size_t)b * seq + (seq - 1)) * H * es;  // text_model.rs:342-346
    void *lo = logits_dev ? (char *)logits_dev + (size_t)b * V * es : c->logits;
    RC(enqueue_head(c, row, lo, false));
    if (argmax_host) CU(cudaMemcpyAsync(c->h_pin + b, c->d_token, 4, cudaMemcpyDeviceToHost, c->stream));
  }
  if (argmax_host) {
    CU(cudaStreamSynchronize(c->stream));
    for (int b = 0; b < batch; b++) argmax_host[b] = c->h_pin[b];
  }
  return CAKE_B200_OK;
}

static int apply_repeat_penalty(cake_b200_ctx *c, void *logits_dev, float penalty, const uint32_t *ctx_tokens_host, int n_tokens);
extern "C" int cake_b200_repeat_penalty_argmax(cake_b200_ctx *c, void *logits_dev, float penalty,
                                               const uint32_t *ctx_tokens_host, int n_tokens, uint32_t *argmax_host) {
  if (!c || !logits_dev || !argmax_host) return fail(CAKE_B200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  RC(apply_repeat_penalty(c, logits_dev, penalty, ctx_tokens_host, n_tokens));
  RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
      typedef typename decltype(tag_)::type T;
    argmax_kernel<T><<<1, 1024, 0, c->stream>>>((const T *)logits_dev, c->cfg.vocab, c->d_token);
    return CAKE_B200_OK;
  }));
  c->launches++;
  CU(cudaMemcpyAsync(c->h_pin, c->d_token, 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *argmax_host = c->h_pin[0];
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ samplers (sample.cuh)
static int check_sampling(const cake_b200_ctx *c, const cake_b200_sampling *s) {
  if (s->kind < 0 || s->kind > 5) return fail(CAKE_B200_EINVAL, "sampling kind %d unknown (0..5)", s->kind);
  if ((s->kind == SAMPLE_TOPK || s->kind == SAMPLE_TOPK_TOPP) && (s->top_k < 1 || (s->top_k > SAMPLE_MAX_K && s->top_k < c->cfg.vocab)))
    return fail(CAKE_B200_EINVAL, "top_k %d unsupported (1..%d, or >= vocab)", s->top_k, SAMPLE_MAX_K);
  return CAKE_B200_OK;
}
static int samp_reserve(cake_b200_ctx *c) {
  if (!c->samp_p) CU(cudaMalloc(&c->samp_p, (size_t)c->cfg.vocab * 4 + 16));
  return CAKE_B200_OK;
}
static SampleArgs to_args(const cake_b200_sampling &s) { return SampleArgs{s.kind, s.top_k, s.temperature, s.top_p, (unsigned long long)s.seed}; }

static int apply_repeat_penalty(cake_b200_ctx *c, void *logits_dev, float penalty, const uint32_t *ctx_tokens_host, int n_tokens) {
  // text_model.rs:66-72: de-duplicate, keeping first occurrences
  std::vector<uint32_t> uniq;
  for (int i = 0; i < n_tokens; i++) {
    bool dup = false;
    for (uint32_t u : uniq) dup |= (u == ctx_tokens_host[i]);
    if (!dup) uniq.push_back(ctx_tokens_host[i]);
  }
  if (uniq.empty() || penalty == 1.0f) return CAKE_B200_OK;
  if (uniq.size() > c->d_pen_cap) {
    CU(cudaStreamSynchronize(c->stream));
    if (c->d_pen) cudaFree(c->d_pen);
    CU(cudaMalloc(&c->d_pen, uniq.size() * 4));
    c->d_pen_cap = uniq.size();
  }
  CU(cudaMemcpyAsync(c->d_pen, uniq.data(), uniq.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CU(cudaStreamSynchronize(c->stream));  // uniq is a stack-owned pageable buffer
  RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
    typedef typename decltype(tag_)::type T;
    repeat_penalty_kernel<T><<<(unsigned)((uniq.size() + 127) / 128), 128, 0, c->stream>>>((T *)logits_dev, c->cfg.vocab, penalty, c->d_pen, (int)uniq.size());
    return CAKE_B200_OK;
  }));
  c->launches++;
  return CAKE_B200_OK;
}

extern "C" int cake_b200_sample(cake_b200_ctx *c, void *logits_dev, const cake_b200_sampling *s, float repeat_penalty,
                                const uint32_t *ctx_tokens_host, int n_tokens, uint64_t step, const float *noise_host,
                                uint32_t *token_host) {
  if (!c || !logits_dev || !s || !token_host || n_tokens < 0 || (n_tokens > 0 && !ctx_tokens_host)) return fail(CAKE_B200_EINVAL, "null argument");
  RC(check_sampling(c, s));
  CU(cudaSetDevice(c->device));
  RC(samp_reserve(c));
  RC(apply_repeat_penalty(c, logits_dev, repeat_penalty, ctx_tokens_host, n_tokens));
  const float *noise_dev = nullptr;
  if (noise_host) {
    const size_t n = (s->kind == SAMPLE_GUMBEL) ? (size_t)c->cfg.vocab : 1;
    if (!c->samp_noise) CU(cudaMalloc(&c->samp_noise, (size_t)c->cfg.vocab * 4 + 16));
    CU(cudaMemcpyAsync(c->samp_noise, noise_host, n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    noise_dev = c->samp_noise;
  }
  set_int_kernel<<<1, 1, 0, c->stream>>>((int *)(c->samp_p + c->cfg.vocab), (int)(step & 0x7fffffff));  // the draw's counter
  RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
    typedef typename decltype(tag_)::type T;
    sample_kernel<T><<<1, SAMPLE_THREADS, 0, c->stream>>>((const T *)logits_dev, c->cfg.vocab, to_args(*s), c->samp_p, noise_dev,
                                                          (const int *)(c->samp_p + c->cfg.vocab), 0, c->d_token, nullptr, 1);
    return CAKE_B200_OK;
  }));
  c->launches += 2;
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(c->h_pin, c->d_token, 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *token_host = c->h_pin[0];
  return CAKE_B200_OK;
}

extern "C" int cake_b200_decode_set_sampling(cake_b200_ctx *c, const cake_b200_sampling *s) {
  if (!c) return fail(CAKE_B200_EINVAL, "null argument");
  if (!s) { c->sampling = cake_b200_sampling{}; return CAKE_B200_OK; }
  RC(check_sampling(c, s));
  CU(cudaSetDevice(c->device));
  RC(samp_reserve(c));  // no allocation inside the graph capture of decode_build
  c->sampling = *s;
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ comm
extern "C" int cake_b200_comm_unique_id(void *out128) {
  if (!out128) return fail(CAKE_B200_EINVAL, "null argument");
  RC(nccl_load());
  ncclUniqueId id;
  NC(g_nccl.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
  return CAKE_B200_OK;
}
extern "C" int cake_b200_comm_init(cake_b200_ctx *c, const void *unique_id128, int rank, int world) {
  if (!c || !unique_id128 || rank < 0 || rank >= world) return fail(CAKE_B200_EINVAL, "bad comm arguments");
  RC(nccl_load());
  CU(cudaSetDevice(c->device));
  ncclUniqueId id;
  memcpy(&id, unique_id128, 128);
  NC(g_nccl.CommInitRank(&c->comm, world, id, rank));
  c->rank = rank;
  c->world = world;
  // Warm up the ring's p2p connections now: NCCL sets a peer connection up lazily on first use, which must
  // not happen inside the CUDA-graph capture of the decode step.
  if (world > 1) {
    const size_t xbytes = (size_t)c->cfg.hidden * c->es;
    NC(g_nccl.GroupStart());
    NC(g_nccl.Send(c->xa, xbytes, ncclUint8, (rank + 1) % world, c->comm, c->stream));
    NC(g_nccl.Recv(c->xb, xbytes, ncclUint8, (rank + world - 1) % world, c->comm, c->stream));
    NC(g_nccl.GroupEnd());
    CU(cudaStreamSynchronize(c->stream));
  }
  return CAKE_B200_OK;
}
extern "C" int cake_b200_send(cake_b200_ctx *c, const void *x_dev, size_t bytes, int peer) {
  if (!c || !c->comm) return fail(CAKE_B200_ESTATE, "communicator not initialised");
  NC(g_nccl.Send(x_dev, bytes, ncclUint8, peer, c->comm, c->stream));
  return CAKE_B200_OK;
}
extern "C" int cake_b200_recv(cake_b200_ctx *c, void *x_dev, size_t bytes, int peer) {
  if (!c || !c->comm) return fail(CAKE_B200_ESTATE, "communicator not initialised");
  NC(g_nccl.Recv(x_dev, bytes, ncclUint8, peer, c->comm, c->stream));
  return CAKE_B200_OK;
}

// Ring hand-off over NVLink peer memory.  Each rank owns an inbox {u64 arrivals, pad to 128 B, x[hidden]} that the
// previous rank's decode kernel writes directly (st.global over NVLink) and releases with one red.release.sys per CTA.
constexpr size_t INBOX_X_OFF = 128;
extern "C" int cake_b200_ring_export(cake_b200_ctx *c, void *handle64) {
  if (!c || !handle64) return fail(CAKE_B200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  if (!c->inbox) {
    CU(cudaMalloc(&c->inbox, INBOX_X_OFF + (size_t)c->cfg.hidden * c->es + 128));
    CU(cudaMemset(c->inbox, 0, INBOX_X_OFF + (size_t)c->cfg.hidden * c->es + 128));
  }
  cudaIpcMemHandle_t h;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CU(cudaIpcGetMemHandle(&h, c->inbox));
  memcpy(handle64, &h, 64);
  return CAKE_B200_OK;
}
extern "C" int cake_b200_ring_import(cake_b200_ctx *c, const void *next_rank_handle64) {
  if (!c || !next_rank_handle64) return fail(CAKE_B200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, next_rank_handle64, 64);
  void *p = nullptr;
  CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->peer_inbox = (unsigned char *)p;
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ decode loop (one CUDA graph per shard)
extern "C" int cake_b200_decode_build(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *block_idx,
                                      int n_blocks, cake_b200_cache *kc, int rank, int world) {
  if (!c || !kc || n_blocks < 0 || world < 1 || rank < 0 || rank >= world) return fail(CAKE_B200_EINVAL, "bad decode_build arguments");
  if (kc->batch != 1) return fail(CAKE_B200_EINVAL, "decode graph is batch 1 (cake run/serve never batch, text_model.rs:418-420)");
  if (world > 1 && !c->comm) return fail(CAKE_B200_ESTATE, "world > 1 needs cake_b200_comm_init first");
  if (rank == 0 && !c->lm_head) return fail(CAKE_B200_ESTATE, "rank 0 needs cake_b200_head_load first");
  for (int i = 0; i < n_blocks; i++)
    if (blocks[i] && blocks[i]->variant())
      return fail(CAKE_B200_EINVAL, "block %d is a sibling block structure (cake_b200_block_set_variant / pre-reshape QK-norm): "
                  "the decode graph covers the standard block only; step it with cake_b200_forward_batch", block_idx[i]);
  CU(cudaSetDevice(c->device));
  RC(wait_loads(c));
  for (int i = 0; i < n_blocks; i++) RC(cache_ensure(kc, block_idx[i]));
  if (c->gexec) { cudaGraphExecDestroy(c->gexec); c->gexec = nullptr; }
  if (c->graph) { cudaGraphDestroy(c->graph); c->graph = nullptr; }
  const size_t xbytes = (size_t)c->cfg.hidden * c->es;
  CU(cudaStreamSynchronize(c->stream));
  if (n_blocks > MK_MAX_LAYERS) return fail(CAKE_B200_EINVAL, "more than %d blocks per shard", MK_MAX_LAYERS);
  if (n_blocks > 0) {
    std::vector<MkLayer> tab(n_blocks);
    fill_mk_table(tab.data(), blocks, block_idx, n_blocks, kc);
    CU(cudaMemcpy(c->g_tab_dev, tab.data(), sizeof(MkLayer) * n_blocks, cudaMemcpyHostToDevice));
  }
  CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
  c->capturing = true;
  c->g_kernels = 0;
  int rc = [&]() -> int {
    if (c->use_mega && mega_supported(c)) {
      // rank 0: [embed+layers(+head if alone)] ; world>1: layers -> send ... recv -> head.  ranks>0: recv -> layers -> send
      MkPlan p;
      const bool p2p = world > 1 && c->inbox && c->peer_inbox;  // hand-off fused into the kernels (no NCCL node)
      unsigned long long *inbox_ctr = (unsigned long long *)c->inbox, *peer_ctr = (unsigned long long *)c->peer_inbox;
      void *inbox_x = c->inbox ? c->inbox + INBOX_X_OFF : nullptr, *peer_x = c->peer_inbox ? c->peer_inbox + INBOX_X_OFF : nullptr;
      if (rank == 0) {
        RC(plan_mega(c, kc, world == 1, &p));
        p.a.layers = c->g_tab_dev; p.a.n_layers = n_blocks; p.a.x_in = nullptr; p.a.x_out = c->xa;
        if (c->cfg.embed_scale != 0.f) {  // text_model.rs:274-276: the scaled embedding row is materialised first
          RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
            typedef typename decltype(tag_)::type T;
            return launch_pdl(c, embed_kernel<T>, dim3(1), dim3(128), 0, (const T *)c->embed, (const uint32_t *)c->d_token,
                              (T *)c->x0, 1, c->cfg.hidden, c->cfg.vocab, c->cfg.embed_scale);
          }));
          p.a.x_in = c->x0;
        }
        p.a.has_head = (world == 1); p.a.advance = (world == 1); p.a.token_ring = c->token_ring;
        if (p2p) { p.a.x_out = peer_x; p.a.peer_ctr = peer_ctr; p.a.ring_seq = c->ring_seq; }
        if (n_blocks > 0 || world == 1) RC(launch_mega(c, p));
        if (world > 1) {
          if (n_blocks == 0) return fail(CAKE_B200_EINVAL, "rank 0 must own at least one layer");
          if (!p2p) {
            RC(cake_b200_send(c, c->xa, xbytes, 1));
            RC(cake_b200_recv(c, c->xa, xbytes, world - 1));
          }
          MkPlan h;
          RC(plan_mega(c, kc, true, &h));
          h.a.layers = c->g_tab_dev; h.a.n_layers = 0; h.a.x_in = p2p ? inbox_x : c->xa; h.a.x_out = c->xa;
          h.a.has_head = 1; h.a.advance = 1; h.a.token_ring = c->token_ring; h.a.trace_tag = 1;
          if (p2p) { h.a.inbox_ctr = inbox_ctr; h.a.ring_seq = c->ring_seq; }
          RC(launch_mega(c, h));
        }
        if (c->sampling.kind != SAMPLE_ARGMAX && c->sampling.temperature > 0.f) {
          // non-greedy: a sampler kernel behind the decode kernel re-draws the token from the logits it left behind
          // (256 KB, L2-hot) and overwrites d_token / the ring slot of this step; the step counter was already advanced
          RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
            typedef typename decltype(tag_)::type T;
            return launch_pdl(c, sample_kernel<T>, dim3(1), dim3(SAMPLE_THREADS), 0, (const T *)c->logits, c->cfg.vocab, to_args(c->sampling),
                              c->samp_p, (const float *)nullptr, (const int *)c->d_step, -1, c->d_token, c->token_ring, (int)TOKEN_RING);
          }));
        }
      } else {
        if (!p2p) RC(cake_b200_recv(c, c->xa, xbytes, rank - 1));
        RC(plan_mega(c, kc, false, &p));
        p.a.layers = c->g_tab_dev; p.a.n_layers = n_blocks; p.a.x_in = p2p ? inbox_x : c->xa; p.a.x_out = p2p ? peer_x : c->xa;
        p.a.has_head = 0; p.a.advance = 1;
        if (p2p) { p.a.inbox_ctr = inbox_ctr; p.a.peer_ctr = peer_ctr; p.a.ring_seq = c->ring_seq; }
        if (n_blocks > 0) RC(launch_mega(c, p));
        else return fail(CAKE_B200_EINVAL, "every rank must own at least one layer");
        if (!p2p) RC(cake_b200_send(c, c->xa, xbytes, (rank + 1) % world));
      }
      return CAKE_B200_OK;
    }
    if (rank == 0) {
      RC(DISPATCH_T(c->cfg.dtype, T_LAMBDA {
      typedef typename decltype(tag_)::type T;
        return launch_pdl(c, embed_kernel<T>, dim3(1), dim3(128), 0, (const T *)c->embed, (const uint32_t *)c->d_token,
                          (T *)c->xa, 1, c->cfg.hidden, c->cfg.vocab, c->cfg.embed_scale);
      }));
    } else {
      RC(cake_b200_recv(c, c->xa, xbytes, rank - 1));
    }
    if (n_blocks > 0) RC(enqueue_decode_layers(c, blocks, block_idx, n_blocks, kc, c->xa, c->xa));
    if (world > 1) {
      RC(cake_b200_send(c, c->xa, xbytes, (rank + 1) % world));
      if (rank == 0) RC(cake_b200_recv(c, c->xa, xbytes, world - 1));
    }
    if (rank == 0) RC(enqueue_head(c, c->xa, c->logits, true));
    RC(launch_pdl(c, advance_kernel, dim3(1), dim3(32), 0, kc->d_pos, c->d_step));
    return CAKE_B200_OK;
  }();
  c->capturing = false;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(c->stream, &g);
  if (rc != CAKE_B200_OK) {
    if (g) cudaGraphDestroy(g);
    return rc;
  }
  if (e != cudaSuccess) return fail(CAKE_B200_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
  c->graph = g;
  CU(cudaGraphInstantiate(&c->gexec, g, 0));
  c->g_cache = kc;
  c->g_block_idx.assign(block_idx, block_idx + n_blocks);
  c->g_rank = rank;
  c->g_world = world;
  return CAKE_B200_OK;
}

extern "C" int cake_b200_decode_begin(cake_b200_ctx *c, uint32_t first_token, int index_pos) {
  if (!c || !c->gexec) return fail(CAKE_B200_ESTATE, "decode_build has not been called");
  CU(cudaSetDevice(c->device));
  RC(wait_loads(c));
  cake_b200_cache *kc = c->g_cache;
  for (int l : c->g_block_idx)
    if (kc->len[l] != index_pos)
      return fail(CAKE_B200_ESTATE, "block %d: index_pos %d != cache length %d", l, index_pos, kc->len[l]);
  if (index_pos >= kc->cap) return fail(CAKE_B200_ESTATE, "cache full");
  if (c->g_rank == 0 && first_token >= (uint32_t)c->cfg.vocab)  // candle's index_select fails on an out-of-range id (text_model.rs:271)
    return fail(CAKE_B200_EINVAL, "token id %u out of range (vocab %d)", first_token, c->cfg.vocab);
  set_int_kernel<<<1, 1, 0, c->stream>>>(kc->d_pos, index_pos);
  set_int_kernel<<<1, 1, 0, c->stream>>>(c->d_step, 0);
  set_u32_kernel<<<1, 1, 0, c->stream>>>(c->d_token, first_token);
  c->launches += 3;
  c->steps_done = 0;
  CU(cudaGetLastError());
  return CAKE_B200_OK;
}

extern "C" int cake_b200_decode_run(cake_b200_ctx *c, int n_steps) {
  if (!c || !c->gexec) return fail(CAKE_B200_ESTATE, "decode_build has not been called");
  CU(cudaSetDevice(c->device));
  cake_b200_cache *kc = c->g_cache;
  int cur = c->g_block_idx.empty() ? 0 : kc->len[c->g_block_idx[0]];
  if (!c->g_block_idx.empty() && cur + n_steps > kc->cap)
    return fail(CAKE_B200_ESTATE, "decode_run: %d steps from position %d exceed cache capacity %d", n_steps, cur, kc->cap);
  for (int i = 0; i < n_steps; i++) CU(cudaGraphLaunch(c->gexec, c->stream));
  for (int l : c->g_block_idx) kc->len[l] += n_steps;
  c->steps_done += n_steps;
  c->launches += c->g_kernels * (uint64_t)n_steps;
  return CAKE_B200_OK;
}

extern "C" int cake_b200_decode_tokens(cake_b200_ctx *c, uint32_t *out_host, int n) {
  if (!c || !out_host || n < 0 || n > TOKEN_RING || n > c->steps_done) return fail(CAKE_B200_EINVAL, "bad decode_tokens arguments");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<uint32_t> ring(TOKEN_RING);
  CU(cudaMemcpy(ring.data(), c->token_ring, (size_t)TOKEN_RING * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) out_host[i] = ring[(size_t)(c->steps_done - n + i) % TOKEN_RING];
  return CAKE_B200_OK;
}

extern "C" int cake_b200_decode_step_host(cake_b200_ctx *c, uint32_t token_in, uint32_t *token_out) {
  if (!c || !c->gexec || !token_out) return fail(CAKE_B200_ESTATE, "decode_build has not been called");
  CU(cudaSetDevice(c->device));
  if (token_in >= (uint32_t)c->cfg.vocab) return fail(CAKE_B200_EINVAL, "token id %u out of range (vocab %d)", token_in, c->cfg.vocab);
  c->h_pin[32] = token_in;
  CU(cudaMemcpyAsync(c->d_token, c->h_pin + 32, 4, cudaMemcpyHostToDevice, c->stream));
  RC(cake_b200_decode_run(c, 1));
  CU(cudaMemcpyAsync(c->h_pin + 33, c->d_token, 4, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  *token_out = c->h_pin[33];
  return CAKE_B200_OK;
}

extern "C" int cake_b200_decode_logits(cake_b200_ctx *c, void *logits_host, size_t bytes) {
  if (!c || !logits_host) return fail(CAKE_B200_EINVAL, "null argument");
  const size_t need = (size_t)c->cfg.vocab * c->es;
  if (bytes < need) return fail(CAKE_B200_EINVAL, "decode_logits needs %zu bytes", need);
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(logits_host, c->logits, need, cudaMemcpyDeviceToHost));
  return CAKE_B200_OK;
}

// ------------------------------------------------------------------------------------------ measurement aid
extern "C" int cake_b200_bench_kernel(cake_b200_ctx *c, cake_b200_block *const *blocks, const int *block_idx,
                                      int n_blocks, cake_b200_cache *kc, int which, int reps, float *ms_per_launch) {
  if (!c || !blocks || !kc || !ms_per_launch || n_blocks < 1 || reps < 1 || which < 0 || which > 4)
    return fail(CAKE_B200_EINVAL, "bad bench_kernel arguments");
  for (int i = 0; i < n_blocks; i++)
    if (!blocks[i] || blocks[i]->variant()) return fail(CAKE_B200_EINVAL, "bench_kernel times the per-op kernels of standard blocks only");
  CU(cudaSetDevice(c->device));
  RC(wait_loads(c));
  for (int i = 0; i < n_blocks; i++) RC(cache_ensure(kc, block_idx[i]));
  const int len = kc->len[block_idx[0]];
  if (which == 4 && len < 1) return fail(CAKE_B200_ESTATE, "attention bench needs a non-empty cache");
  if (which == 4) {  // scores the token at position len-1 again: reads len rows, rewrites row len-1
    set_int_kernel<<<1, 1, 0, c->stream>>>(kc->d_pos, len - 1);
  }
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  auto round = [&]() -> int {
    for (int i = 0; i < n_blocks; i++) {
      const cake_b200_block *b = blocks[i];
      switch (which) {
        case 0: RC(enqueue_qkv(c, b, c->xa)); break;
        case 1: RC(enqueue_oproj(c, b, c->xa, c->xb)); break;
        case 2: RC(enqueue_gate_up(c, b, c->xb)); break;
        case 3: RC(enqueue_down(c, b, c->xb, c->xb)); break;
        default: RC(enqueue_attn(c, b, kc, block_idx[i])); break;
      }
    }
    return CAKE_B200_OK;
  };
  RC(round());  // warm-up
  CU(cudaEventRecord(e0, c->stream));
  for (int r = 0; r < reps; r++) RC(round());
  CU(cudaEventRecord(e1, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_per_launch = ms / (float)(reps * n_blocks);
  return CAKE_B200_OK;
}

extern "C" int cake_b200_load_stats(cake_b200_ctx *c, double *bytes, double *seconds) {
  if (!c || !bytes || !seconds) return fail(CAKE_B200_EINVAL, "null argument");
  CU(cudaSetDevice(c->device));
  if (c->load_stream) {  // include the tail of the last DMA in the wall time of the caller, not in `seconds` (host staging time)
    CU(cudaStreamSynchronize(c->load_stream));
  }
  *bytes = c->load_bytes;
  *seconds = c->load_seconds;
  return CAKE_B200_OK;
}

extern "C" int cake_b200_decode_trace(cake_b200_ctx *c, uint64_t *out_host, int n_steps) {
  if (!c || !out_host || n_steps < 0 || n_steps > MK_STEP_RING || n_steps > c->steps_done)
    return fail(CAKE_B200_EINVAL, "bad decode_trace arguments (at most %d of the steps since decode_begin)", MK_STEP_RING);
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  std::vector<unsigned long long> ring((size_t)MK_STEP_RING * 8);
  CU(cudaMemcpy(ring.data(), c->step_trace, ring.size() * 8, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n_steps; i++) {
    const size_t slot = (size_t)((c->steps_done - n_steps + i) % MK_STEP_RING) * 8;
    for (int j = 0; j < 8; j++) out_host[(size_t)i * 8 + j] = ring[slot + j];
  }
  return CAKE_B200_OK;
}

/* profiling aid (not in the public header): copies the megakernel's phase-boundary %globaltimer stamps of
 * the last launch (CTA 0) to `out`; needs CAKE_B200_MEGA_TRACE=1 at ctx creation. */
extern "C" int cake_b200_debug_trace(cake_b200_ctx *c, unsigned long long *out, int n) {
  if (!c || !c->trace || n > 4096) return fail(CAKE_B200_ESTATE, "tracing is off (CAKE_B200_MEGA_TRACE=1)");
  CU(cudaSetDevice(c->device));
  CU(cudaStreamSynchronize(c->stream));
  CU(cudaMemcpy(out, c->trace, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return CAKE_B200_OK;
}
