// decode_mega.cuh — the decode hot loop as ONE persistent kernel per shard-step.
//
// Why: with one kernel per op (5 per layer) every dependency between ops is a kernel boundary that
// costs ~3 us of HBM idle time even with programmatic dependent launch (profiles/gemv_sweep_r01.txt:
// o_proj 7.2 us streamed vs 5.1 us ideal) — 130 boundaries per token = ~0.4 ms of a 2.3 ms budget.
// Here the whole walk  [rms_1+qkv -> rope/append/attention -> o_proj+res -> rms_2+gate_up+silu*mul ->
// down+res] x layers (+ ln_f/lm_head/argmax on the master)  runs inside one grid of one CTA per SM:
//   * warp 16 is the PRODUCER: it walks the same static phase list and issues TMA bulk copies of the
//     next weight rows (32 KB stages, evict_first) and of this CTA's K/V cache tile into a 5-6 stage
//     shared-memory ring.  It never takes part in a grid barrier, so HBM keeps streaming while the
//     consumers synchronise — the ring (~160 KB/SM, ~24 MB chip-wide) absorbs the bubbles;
//   * warps 0-15 are CONSUMERS: per phase they wait for the grid barrier that publishes the previous
//     phase's output vector, stage x (RMS-normalised in-kernel), multiply-accumulate the staged rows in
//     fp32 and run the fused epilogue;
//   * grid barriers are a 64-bit monotonic counter in global memory (release: bar.sync + threadfence +
//     atomicAdd; acquire: ld.acquire poll + threadfence); activations are read with ld.global.cg.
// Arithmetic and rounding points are identical to the per-op kernels (gemv.cuh / attn_decode.cuh); only the
// fp32 summation order within a dot product differs (tests/test_gpu_parity.py::test_megakernel_equals_per_op_kernels).
#pragma once
#include "common.cuh"
#include "gemv.cuh"
#include "attn_decode.cuh"

namespace cake {

constexpr int MK_CW = 16;                       // consumer warps
constexpr int MK_CT = MK_CW * 32;               // consumer threads
constexpr int MK_THREADS = (MK_CW + 1) * 32;    // + producer warp
constexpr int MK_STAGE_BYTES = 32768;
constexpr int MK_MAX_STAGES = 6;
constexpr int MK_MAX_LAYERS = 96;

struct MkLayer {
  const void *wqkv, *wo, *wgu, *wd, *ln1, *ln2, *bqkv, *qn, *kn;
  void *kc, *vc;
};
struct MkGeom {  // one decode GEMV type; a stage holds RS row segments of KC columns
  int N, K, KC, RS, WPR, RPW;
};
struct MkArgs {
  const MkLayer *layers;
  int n_layers;
  MkGeom g_qkv, g_o, g_gu, g_down, g_head;
  int hidden, inter, n_heads, n_kv, hd, rot, cap, nsplit, max_k;
  float eps, scale;
  const void *x_in;          // input hidden state (nullptr on the master: embed row of *d_token)
  const void *embed;
  const uint32_t *d_token;
  void *x_out;               // output hidden state of this shard
  void *xa, *xb, *qkv, *y, *mm;
  float *ws_ml, *ws_acc;
  unsigned *attn_counters;
  const void *cos_t, *sin_t;
  int *d_pos, *d_step;
  unsigned long long *gbar;      // [0] monotonic barrier arrival counter, [1] its value at launch start
  int has_head, advance, n_stages, vocab, partial_floats;
  const void *ln_f, *lm_head;
  void *logits;
  float *part_val;
  int *part_idx;
  unsigned *argmax_counter;
  uint32_t *token_out, *token_ring;
  int ring_cap;
  // shard-to-shard hand-off over NVLink peer memory (fused into this kernel; no NCCL kernel in the step):
  //   inbox_ctr != nullptr: wait until the previous shard has delivered step *ring_seq into our inbox (x_in)
  //   peer_ctr  != nullptr: x_out is the next shard's inbox; signal it when every CTA has written its rows
  const unsigned long long *inbox_ctr;
  unsigned long long *peer_ctr;
  unsigned long long *ring_seq;   // hand-offs received so far (device-resident, never reset)
  unsigned *tickets;          // [phase] work-claim counters, zeroed before every launch
  int max_groups;             // per-CTA cap of row groups in one phase (sizes the partial-sum scratch)
  int l2_prefetch;            // experimental (CAKE_B200_L2_PREFETCH=1): producer prefetches its future rows into L2
  unsigned long long *trace;  // optional: %globaltimer stamps of CTA 0 at phase boundaries (profiling aid)
  // always on (3 stores per launch): %globaltimer of CTA 0 at kernel entry / input acquired / exit, ring of
  // MK_STEP_RING steps x 2 launches (tag 0: the layer launch, tag 1: rank 0's head-only launch when sharded) x 4 u64
  unsigned long long *step_trace;
  int trace_tag;
  int act;     // mlp.rs:25-26: 0 silu, 1 gelu_tanh
  int window;  // cache.rs:173-205 sliding window (0 = full context): attention covers the last `window` positions
};
constexpr int MK_STEP_RING = 2048;
__device__ __forceinline__ unsigned long long mk_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__host__ __device__ inline size_t mk_xs_bytes(int max_k, int es) {
  size_t b = (size_t)max_k * es;
  const size_t attn = (size_t)(ATTN_MAX_G * 256 * 4) + (size_t)ATTN_MAX_G * ATTN_TILE * 4 + (size_t)2 * 256 * 4;
  return b > attn ? b : attn;  // the attention phase overlays its scratch on the x buffer
}
__host__ __device__ inline size_t mk_smem_bytes(int max_k, int partial_floats, int max_groups, int n_stages, int es) {
  size_t off = (size_t)n_stages * MK_STAGE_BYTES;
  off += mk_xs_bytes(max_k, es);
  off = (off + 15) & ~(size_t)15;
  off += (size_t)partial_floats * 4;     // partial sums [row][slice]
  off += (size_t)max_groups * 4;         // first row of each locally processed group
  off += 128 * 4;                        // scratch
  off = (off + 7) & ~(size_t)7;
  off += (size_t)2 * MK_MAX_STAGES * 8;  // mbarriers
  off += (size_t)MK_MAX_STAGES * 4;      // stage_row
  return off + 256;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ldcg_v4(const void *p) { return __ldcg(reinterpret_cast<const uint4 *>(p)); }

struct MkRing {
  unsigned char *ring;
  uint64_t *full, *empty;
  int *stage_row;  // first weight row held by a stage, or -1 = end-of-phase marker (dynamic phases)
  int n_stages;
  int s;
  uint32_t ph;
  __device__ __forceinline__ void advance() {
    if (++s == n_stages) { s = 0; ph ^= 1u; }
  }
};

// rows [r0,r1) of this CTA for a GEMV over N rows with row granularity G
__device__ __forceinline__ void mk_rows(int N, int G, int &r0, int &r1) {
  const long units = N / G;
  r0 = (int)(units * blockIdx.x / gridDim.x) * G;
  r1 = (int)(units * (blockIdx.x + 1) / gridDim.x) * G;
}

// Dynamic phases: the N/RS row groups of a GEMV are split into a static part (75 %: CTA c owns groups
// [c*sg, (c+1)*sg)) and a pool claimed one group at a time through an atomic ticket.  Per-SM HBM bandwidth is
// not uniform (measured: the slowest CTA of the gate_up phase finished 5.5 us after the fastest with static
// rows), so fast SMs take more pool groups and all CTAs reach the grid barrier together.  A row's dot product
// is computed by one CTA with a fixed lane/warp mapping whichever CTA it is -> results stay bit-deterministic.
struct MkSplit {
  int n_groups, sg, pool_start;
};
__device__ __forceinline__ MkSplit mk_split(const MkGeom &g) {
  MkSplit sp;
  sp.n_groups = (g.N + g.RS - 1) / g.RS;
  sp.sg = (int)(((long)sp.n_groups * 3 / 4) / gridDim.x);
  sp.pool_start = sp.sg * gridDim.x;
  return sp;
}
__device__ __forceinline__ bool mk_is_dynamic(const MkGeom &g) { return g.K == g.KC && (g.KC / 8) / g.WPR <= 128; }

// L2 prefetch cursor.  HBM idles whenever the consumers sit in a grid barrier or in the attention phase and the
// shared-memory ring is full (~20 us per layer).  The producer therefore also walks its OWN future static row
// groups — across phase and layer boundaries — MK_PF_AHEAD groups ahead of the load cursor and issues
// cp.async.bulk.prefetch.L2 for them: the bytes stream into the 126 MB L2 during those windows and the later
// TMA loads hit L2.  ~14 x 32 KB x 148 CTAs = 66 MB in flight.
constexpr int MK_PF_AHEAD = 14;
struct MkPrefetch {
  const MkArgs *a;
  int l, ph, grp;  // layer, phase within the layer (0 qkv, 1 o, 2 gate_up, 3 down; l == n_layers: head), next group
  int es;
  __device__ __forceinline__ bool phase(const MkGeom *&g, const void *&W) const {
    if (l < a->n_layers) {
      const MkLayer &L = a->layers[l];
      g = ph == 0 ? &a->g_qkv : ph == 1 ? &a->g_o : ph == 2 ? &a->g_gu : &a->g_down;
      W = ph == 0 ? L.wqkv : ph == 1 ? L.wo : ph == 2 ? L.wgu : L.wd;
      return true;
    }
    if (l == a->n_layers && a->has_head && ph == 0) { g = &a->g_head; W = a->lm_head; return true; }
    return false;
  }
  __device__ __forceinline__ void step() {  // prefetch one more static group of this CTA, if any is left
    while (true) {
      const MkGeom *g;
      const void *W;
      if (!phase(g, W)) return;
      if (mk_is_dynamic(*g)) {
        const MkSplit sp = mk_split(*g);
        if (grp < sp.sg) {
          const int row = ((int)blockIdx.x * sp.sg + grp) * g->RS;
          const size_t rowb = (size_t)g->K * es;
          const unsigned bytes = (unsigned)(min(g->RS, g->N - row) * rowb);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const unsigned char *>(W) + (size_t)row * rowb), "r"(bytes) : "memory");
          grp++;
          return;
        }
      }
      grp = 0;
      if (l < a->n_layers && ph < 3) ph++;
      else { l++; ph = 0; }
    }
  }
};

// Targeted variant (l2_prefetch == 2): only while HBM is otherwise idle — right after the attention K/V tiles were
// issued — pull this CTA's static rows of the following GEMV phases into L2.
__device__ __forceinline__ void mk_prefetch_static(const MkGeom &g, const void *W, int es, int first, int count) {
  if (!mk_is_dynamic(g)) return;
  const MkSplit sp = mk_split(g);
  const size_t rowb = (size_t)g.K * es;
  for (int i = first; i < sp.sg && i < first + count; i++) {
    const int row = ((int)blockIdx.x * sp.sg + i) * g.RS;
    const unsigned bytes = (unsigned)(min(g.RS, g.N - row) * rowb);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const unsigned char *>(W) + (size_t)row * rowb), "r"(bytes) : "memory");
  }
}

template <typename T>
__device__ __forceinline__ void mk_produce_gemv_dyn(MkRing &rg, const MkGeom &g, const void *W, unsigned *ticket, int max_groups,
                                                    uint64_t pol, MkPrefetch &pf) {
  constexpr int es = sizeof(T);
  const MkSplit sp = mk_split(g);
  const size_t rowb = (size_t)g.K * es;
  const unsigned char *Wb = reinterpret_cast<const unsigned char *>(W);
  int issued = 0;
  auto issue = [&](int grp) {
    const int row = grp * g.RS, nr = min(g.RS, g.N - row);
    mbar_wait(&rg.empty[rg.s], rg.ph ^ 1u);
    rg.stage_row[rg.s] = row;
    mbar_arrive_expect_tx(&rg.full[rg.s], (uint32_t)(nr * rowb));
    bulk_g2s(rg.ring + (size_t)rg.s * MK_STAGE_BYTES, Wb + (size_t)row * rowb, (uint32_t)(nr * rowb), &rg.full[rg.s], pol);
    rg.advance();
    issued++;
  };
  for (int grp = blockIdx.x * sp.sg; grp < (int)(blockIdx.x + 1) * sp.sg; grp++) {
    issue(grp);
    if (pf.a->l2_prefetch == 1) pf.step();  // keep the L2 prefetch MK_PF_AHEAD static groups ahead of the load cursor
  }
  if (sp.pool_start < sp.n_groups && issued < max_groups) {
    unsigned next = atomicAdd(ticket, 1u);
    while (sp.pool_start + (int)next < sp.n_groups) {
      const int cur = sp.pool_start + (int)next;
      const bool more = issued + 1 < max_groups;  // a claimed group is never dropped: claim only what fits
      if (more) next = atomicAdd(ticket, 1u);     // claim ahead: the round trip overlaps the wait for a free stage
      issue(cur);
      if (!more) break;
    }
  }
  mbar_wait(&rg.empty[rg.s], rg.ph ^ 1u);  // end-of-phase marker
  rg.stage_row[rg.s] = -1;
  mbar_arrive(&rg.full[rg.s]);
  rg.advance();
}

template <typename T>
__device__ __forceinline__ void mk_produce_gemv(MkRing &rg, const MkGeom &g, const void *W, int G, uint64_t pol) {
  constexpr int es = sizeof(T);
  int r0, r1;
  mk_rows(g.N, G, r0, r1);
  const int nchunk = g.K / g.KC;
  const size_t seg = (size_t)g.KC * es;
  const unsigned char *Wb = reinterpret_cast<const unsigned char *>(W);
  for (int row = r0; row < r1; row += g.RS) {
    const int nr = min(g.RS, r1 - row);
    for (int j = 0; j < nchunk; j++) {
      mbar_wait(&rg.empty[rg.s], rg.ph ^ 1u);
      unsigned char *dst = rg.ring + (size_t)rg.s * MK_STAGE_BYTES;
      mbar_arrive_expect_tx(&rg.full[rg.s], (uint32_t)(nr * seg));
      if (nchunk == 1) {
        bulk_g2s(dst, Wb + (size_t)row * g.K * es, (uint32_t)(nr * seg), &rg.full[rg.s], pol);
      } else {
        for (int r = 0; r < nr; r++)
          bulk_g2s(dst + r * seg, Wb + ((size_t)(row + r) * g.K + (size_t)j * g.KC) * es, (uint32_t)seg, &rg.full[rg.s], pol);
      }
      rg.advance();
    }
  }
}

struct MkAttnItem {
  int active, kvh, split, s0, s1, tile;
};
template <typename T>
__device__ __forceinline__ MkAttnItem mk_attn_item(const MkArgs &a, int pos) {
  MkAttnItem it;
  const int items = a.n_kv * a.nsplit;
  it.active = (int)blockIdx.x < items;
  it.kvh = blockIdx.x / a.nsplit;
  it.split = blockIdx.x % a.nsplit;
  const int Tn = pos + 1;
  int per = (Tn + a.nsplit - 1) / a.nsplit;
  per = (per + 7) & ~7;
  it.s0 = min(Tn, it.split * per);
  it.s1 = min(Tn, it.s0 + per);
  const int rows = MK_STAGE_BYTES / (a.hd * (int)sizeof(T));
  it.tile = rows < ATTN_TILE ? rows : ATTN_TILE;
  if (!it.active) it.s0 = it.s1 = 0;
  return it;
}

template <typename T>
__device__ __forceinline__ void mk_produce_attn(MkRing &rg, const MkArgs &a, const MkLayer &L, int pos, uint64_t pol) {
  // sliding window: the visible rows start at ws = pos + 1 - window; everything below works on row indices relative to ws
  const int ws = (a.window > 0 && pos + 1 > a.window) ? pos + 1 - a.window : 0;
  pos -= ws;
  const MkAttnItem it = mk_attn_item<T>(a, pos);
  const int HD = a.hd;
  const T *kc = reinterpret_cast<const T *>(L.kc) + ((size_t)it.kvh * a.cap + ws) * HD;
  const T *vc = reinterpret_cast<const T *>(L.vc) + ((size_t)it.kvh * a.cap + ws) * HD;
  for (int t0 = it.s0; t0 < it.s1; t0 += it.tile) {
    const int t1 = min(it.s1, t0 + it.tile);
    const int nold = min(t1, pos) - t0;  // rows already in the cache (the appended row is handled by the consumers)
    for (int kv = 0; kv < 2; kv++) {
      mbar_wait(&rg.empty[rg.s], rg.ph ^ 1u);
      if (nold > 0) {
        const uint32_t bytes = (uint32_t)nold * HD * sizeof(T);
        mbar_arrive_expect_tx(&rg.full[rg.s], bytes);
        bulk_g2s(rg.ring + (size_t)rg.s * MK_STAGE_BYTES, (kv ? vc : kc) + (size_t)t0 * HD, bytes, &rg.full[rg.s], pol);
      } else {
        mbar_arrive(&rg.full[rg.s]);
      }
      rg.advance();
    }
  }
}

// ---------------------------------------------------------------------------------------- consumers
__device__ __forceinline__ void mk_grid_sync(unsigned long long *ctr, unsigned long long target, int ct) {
  named_bar_sync(1, MK_CT);  // every consumer thread of this CTA has issued its global writes
  if (ct == 0) {
    // release: publishes this CTA's writes (ordered before us by the bar.sync above); fire-and-forget
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ULL) : "memory");
    while (ld_acquire_u64(ctr) < target) {
    }
  }
  named_bar_sync(1, MK_CT);  // data written by other SMs is read with ld.global.cg below (L1 is bypassed)
}

// x -> shared (D), optionally RMS-normalised.  Activations were written by other SMs: ld.global.cg.
template <typename T>
__device__ __forceinline__ void mk_stage_x(T *xs, const void *x, const void *norm_w, int K, float eps, float *scratch,
                                           int ct, int warp, int lane) {
  constexpr int es = sizeof(T);
  uint4 *xsv = reinterpret_cast<uint4 *>(xs);
  const int nv = K * es / 16;
  const char *xb = reinterpret_cast<const char *>(x);
  if (norm_w == nullptr) {
    for (int v = ct; v < nv; v += MK_CT) xsv[v] = ldcg_v4(xb + (size_t)v * 16);
  } else {
    float ss = 0.f;
    for (int v = ct; v < nv; v += MK_CT) {
      const uint4 u = ldcg_v4(xb + (size_t)v * 16);
      xsv[v] = u;
      float f[8];
      unpack8<T>(u, f);
#pragma unroll
      for (int i = 0; i < 8; i++) ss += f[i] * f[i];
    }
    ss = warp_sum(ss);
    if (lane == 0) scratch[warp] = ss;
    named_bar_sync(1, MK_CT);
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < MK_CW; w++) tot += scratch[w];
    const float inv = 1.0f / sqrtf(tot / (float)K + eps);
    const uint4 *wg = reinterpret_cast<const uint4 *>(norm_w);
    for (int v = ct; v < nv; v += MK_CT) {
      float f[8], w8[8];
      unpack8<T>(xsv[v], f);
      unpack8<T>(wg[v], w8);
      uint4 o;
      o.x = pack2<T>(f[0] * inv * w8[0], f[1] * inv * w8[1]);
      o.y = pack2<T>(f[2] * inv * w8[2], f[3] * inv * w8[3]);
      o.z = pack2<T>(f[4] * inv * w8[4], f[5] * inv * w8[5]);
      o.w = pack2<T>(f[6] * inv * w8[6], f[7] * inv * w8[7]);
      xsv[v] = o;
    }
  }
  named_bar_sync(1, MK_CT);
}

struct MkEpi {
  const void *bias, *residual;
  void *out;
  int act;  // EPI_SWIGLU: 0 silu, 1 gelu_tanh
  // argmax
  float *part_val;
  int *part_idx;
  unsigned *counter;
  uint32_t *token_out, *token_ring;
  const int *step;
  int ring_cap;
  // end-of-step bookkeeping done by the argmax finaliser
  int advance;
  int *d_pos, *d_step;
  unsigned long long *gbar;
  unsigned long long next_start;
  unsigned long long *ring_seq;  // non-null when this launch consumed an inbox hand-off
};

template <typename T, int EPI>
__device__ __forceinline__ void mk_consume_gemv(MkRing &rg, const MkGeom &g, const T *xs, float *partial, int *loc_row,
                                                float *scratch, const MkEpi &e, int ct, int warp, int lane) {
  constexpr int G = (EPI == EPI_SWIGLU) ? 2 : 1;
  const int RS = g.RS, WPR = g.WPR, RPW = g.RPW, N = g.N;
  const int nchunk = g.K / g.KC;
  const int slots = MK_CW / WPR, slot = warp / WPR, ks = warp % WPR;
  const int segv = g.KC / 8, nvec = segv / WPR;
  const uint4 *xsv = reinterpret_cast<const uint4 *>(xs);
  const bool dyn = mk_is_dynamic(g);
  int r0 = 0, nrows = 0, nloc = 0;
  if (dyn) {
    // A lane's columns are the same for every row of the phase, so its x slice is converted to fp32 ONCE into
    // registers; per 16-byte weight vector the loop is 1 LDS.128 + 8 converts + 8 FMA (ncu r01: re-reading and
    // re-converting x per vector made the loop issue-bound).  Stages arrive with their row in stage_row[].
    float xr[4][8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int v = lane + 32 * i;
      uint4 u = make_uint4(0u, 0u, 0u, 0u);
      if (v < nvec) u = xsv[ks * nvec + v];
      unpack8<T>(u, xr[i]);
    }
    const bool v3 = lane + 96 < nvec, v2 = lane + 64 < nvec, v1 = lane + 32 < nvec, v0 = lane < nvec;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    while (true) {
      mbar_wait(&rg.full[rg.s], rg.ph);
      const int row = rg.stage_row[rg.s];
      if (row >= 0) {
        const uint4 *st = reinterpret_cast<const uint4 *>(rg.ring + (size_t)rg.s * MK_STAGE_BYTES) + ks * nvec + lane;
        for (int r = 0; r < RPW; r++) {
          const uint4 *rowp = st + (size_t)(slot + r * slots) * segv;
          const uint4 w0 = v0 ? rowp[0] : z, w1 = v1 ? rowp[32] : z, w2 = v2 ? rowp[64] : z, w3 = v3 ? rowp[96] : z;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, wf[8];
          unpack8<T>(w0, wf);
          a0 = fmaf(wf[0], xr[0][0], a0); a1 = fmaf(wf[1], xr[0][1], a1); a2 = fmaf(wf[2], xr[0][2], a2); a3 = fmaf(wf[3], xr[0][3], a3);
          a0 = fmaf(wf[4], xr[0][4], a0); a1 = fmaf(wf[5], xr[0][5], a1); a2 = fmaf(wf[6], xr[0][6], a2); a3 = fmaf(wf[7], xr[0][7], a3);
          unpack8<T>(w1, wf);
          a0 = fmaf(wf[0], xr[1][0], a0); a1 = fmaf(wf[1], xr[1][1], a1); a2 = fmaf(wf[2], xr[1][2], a2); a3 = fmaf(wf[3], xr[1][3], a3);
          a0 = fmaf(wf[4], xr[1][4], a0); a1 = fmaf(wf[5], xr[1][5], a1); a2 = fmaf(wf[6], xr[1][6], a2); a3 = fmaf(wf[7], xr[1][7], a3);
          unpack8<T>(w2, wf);
          a0 = fmaf(wf[0], xr[2][0], a0); a1 = fmaf(wf[1], xr[2][1], a1); a2 = fmaf(wf[2], xr[2][2], a2); a3 = fmaf(wf[3], xr[2][3], a3);
          a0 = fmaf(wf[4], xr[2][4], a0); a1 = fmaf(wf[5], xr[2][5], a1); a2 = fmaf(wf[6], xr[2][6], a2); a3 = fmaf(wf[7], xr[2][7], a3);
          unpack8<T>(w3, wf);
          a0 = fmaf(wf[0], xr[3][0], a0); a1 = fmaf(wf[1], xr[3][1], a1); a2 = fmaf(wf[2], xr[3][2], a2); a3 = fmaf(wf[3], xr[3][3], a3);
          a0 = fmaf(wf[4], xr[3][4], a0); a1 = fmaf(wf[5], xr[3][5], a1); a2 = fmaf(wf[6], xr[3][6], a2); a3 = fmaf(wf[7], xr[3][7], a3);
          const float v = warp_sum((a0 + a1) + (a2 + a3));
          if (lane == 0) partial[((size_t)nloc * RS + slot + r * slots) * WPR + ks] = v;  // rows past N: never read
        }
        if (ct == 0) loc_row[nloc] = row;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&rg.empty[rg.s]);
      rg.advance();
      if (row < 0) break;
      nloc++;
    }
  } else {
    int r1;
    mk_rows(N, G, r0, r1);
    nrows = r1 - r0;
    for (int g0 = 0; g0 < nrows; g0 += RS) {
      float acc[4][2];
#pragma unroll
      for (int r = 0; r < 4; r++) acc[r][0] = acc[r][1] = 0.f;
      for (int j = 0; j < nchunk; j++) {
        mbar_wait(&rg.full[rg.s], rg.ph);
        const uint4 *st = reinterpret_cast<const uint4 *>(rg.ring + (size_t)rg.s * MK_STAGE_BYTES);
        const uint4 *xc = xsv + (size_t)j * segv + ks * nvec;
#pragma unroll 2
        for (int v = lane; v < nvec; v += 32) {
          float xf[8];
          unpack8<T>(xc[v], xf);
#pragma unroll
          for (int r = 0; r < 4; r++) {
            if (r < RPW) {
              float wf[8];
              unpack8<T>(st[(size_t)(slot + r * slots) * segv + ks * nvec + v], wf);
#pragma unroll
              for (int i = 0; i < 8; i++) acc[r][i & 1] = fmaf(wf[i], xf[i], acc[r][i & 1]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&rg.empty[rg.s]);
        rg.advance();
      }
#pragma unroll
      for (int r = 0; r < 4; r++) {
        if (r < RPW) {
          const float v = warp_sum(acc[r][0] + acc[r][1]);
          const int rl = g0 + slot + r * slots;
          if (lane == 0 && rl < nrows) partial[rl * WPR + ks] = v;
        }
      }
    }
  }
  named_bar_sync(1, MK_CT);
  // local row index li -> global row (dynamic: group list; static: contiguous range)
  const int n_local = dyn ? nloc * RS : nrows;
  auto row_of = [&](int li) { return dyn ? loc_row[li / RS] + li % RS : r0 + li; };
  auto row_sum = [&](int li) {
    float sum = 0.f;
    for (int w = 0; w < WPR; w++) sum += partial[(size_t)li * WPR + w];
    return rnd<T>(sum);
  };
  T *out = reinterpret_cast<T *>(e.out);
  if (EPI == EPI_PLAIN) {
    const T *bias = reinterpret_cast<const T *>(e.bias);
    for (int li = ct; li < n_local; li += MK_CT) {
      const int row = row_of(li);
      if (row >= N) continue;
      float v = row_sum(li);
      if (bias) v = rnd<T>(v + DT<T>::to_f(bias[row]));
      out[row] = DT<T>::from_f(v);
    }
  } else if (EPI == EPI_RESIDUAL) {
    const T *res = reinterpret_cast<const T *>(e.residual);
    for (int li = ct; li < n_local; li += MK_CT) {
      const int row = row_of(li);
      if (row >= N) continue;
      const float v = row_sum(li);
      out[row] = DT<T>::from_f(v + DT<T>::to_f(ldcg_T<T>(res + row)));
    }
  } else if (EPI == EPI_SWIGLU) {
    for (int p = ct; p < n_local / 2; p += MK_CT) {
      const int row = row_of(2 * p);
      if (row >= N) continue;
      const float gte = row_sum(2 * p), up = row_sum(2 * p + 1);
      const float sl = gate_act<T>(gte, e.act);
      out[row / 2] = DT<T>::from_f(sl * up);
    }
  } else {
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int li = ct; li < n_local; li += MK_CT) {
      const int row = row_of(li);
      if (row >= N) continue;
      const float v = row_sum(li);
      if (out) out[row] = DT<T>::from_f(v);
      if (v > best || (v == best && row < bidx)) { best = v; bidx = row; }  // first maximum wins
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    int *scratch_i = reinterpret_cast<int *>(scratch + 32);
    if (lane == 0) { scratch[warp] = best; scratch_i[warp] = bidx; }
    named_bar_sync(1, MK_CT);
    if (ct == 0) {
      for (int w = 1; w < MK_CW; w++)
        if (scratch[w] > best || (scratch[w] == best && scratch_i[w] < bidx)) { best = scratch[w]; bidx = scratch_i[w]; }
      e.part_val[blockIdx.x] = best;
      e.part_idx[blockIdx.x] = bidx;
      __threadfence();
      const unsigned ticket = atomicAdd(e.counter, 1u);
      if (ticket == gridDim.x - 1) {
        __threadfence();
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (unsigned c = 0; c < gridDim.x; c++) {
          const float v = __ldcg(e.part_val + c);
          const int i = __ldcg(e.part_idx + c);
          if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        *e.token_out = (uint32_t)bi;
        if (e.token_ring) e.token_ring[*e.step % e.ring_cap] = (uint32_t)bi;
        *e.counter = 0;
        // end-of-step bookkeeping: this is the last CTA of the grid to finish
        if (e.advance) { *e.d_pos += 1; *e.d_step += 1; }
        if (e.ring_seq) *e.ring_seq += 1ULL;
        e.gbar[1] = e.next_start;
      }
    }
  }
}

// Sum G per-lane values over the 16 lanes of a half-warp with a halving butterfly: after log2(G) exchange
// steps every lane owns ONE head's partial, then the remaining xor steps finish it.  G + log2(16/G) - 1
// shuffles instead of 4*G.  Returns the sum for head `g_out` (valid on every lane of the half-warp).
template <int G>
__device__ __forceinline__ float reduce16_heads(float (&s)[G], int lane, int &g_out) {
  int g = 0;
  int bit = 8;
#pragma unroll
  for (int n = G; n > 1; n >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; i++) {
      const float keep = upper ? s[i + n / 2] : s[i];
      const float send = upper ? s[i] : s[i + n / 2];
      s[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
    if (upper) g += n / 2;
    bit >>= 1;
  }
  float v = s[0];
#pragma unroll
  for (int b2 = 8; b2 >= 1; b2 >>= 1)
    if (b2 <= bit) v += __shfl_xor_sync(0xffffffffu, v, b2);
  g_out = g;
  return v;
}

// norm_rope_inplace with the position's cos/sin row already staged in shared memory as fp32 (D values)
template <typename T, int HD>
__device__ __forceinline__ void mk_norm_rope(float *buf, const T *norm_w, float eps, const float *cs, const float *sn, int rot, int lane) {
  if (norm_w) {
    float ss = 0.f;
    for (int d = lane; d < HD; d += 32) ss += buf[d] * buf[d];
    ss = warp_sum(ss);
    const float inv = 1.0f / sqrtf(ss / (float)HD + eps);
    for (int d = lane; d < HD; d += 32) buf[d] = rnd<T>(buf[d] * inv * DT<T>::to_f(norm_w[d]));
    __syncwarp();
  }
  const int half = rot / 2;
  for (int i = lane; i < half; i += 32) {
    const float c = cs[i], s = sn[i];
    const float x1 = buf[i], x2 = buf[i + half];
    buf[i] = rnd<T>(rnd<T>(x1 * c) - rnd<T>(x2 * s));
    buf[i + half] = rnd<T>(rnd<T>(x2 * c) + rnd<T>(x1 * s));
  }
  __syncwarp();
}

// Attention phase for this CTA's (kv head, split) item.  Same math as attn_decode_kernel (f32 scores,
// softmax and PV, one rounding of the result); K/V tiles come from the ring, prefetched by the producer while
// the qkv GEMV was still running.  G = n_heads / n_kv_heads is a template parameter so that all per-head
// state lives in registers.
template <typename T, int HD, int G>
__device__ __forceinline__ void mk_consume_attn(MkRing &rg, const MkArgs &a, const MkLayer &L, int pos, unsigned char *scr,
                                                const float *rope_cs, int ct, int warp, int lane, unsigned long long *tr) {
  auto stamp = [&](int i) {
    if (tr && ct == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      tr[i] = t;
    }
  };
  stamp(0);
  constexpr int LPR = HD / 8;            // lanes per cached row (16 B each); HD in {16,64,128} -> 2, 8, 16
  constexpr int RPWI = 32 / LPR;         // rows per warp per iteration
  constexpr int NW = MK_CW;
  static_assert(HD == 16 || HD == 64 || HD == 128, "head_dim");
  // sliding window (cache.rs:173-205): rows below ws = pos + 1 - window are invisible; indices below are relative to ws
  const int ws = (a.window > 0 && pos + 1 > a.window) ? pos + 1 - a.window : 0;
  pos -= ws;
  const MkAttnItem it = mk_attn_item<T>(a, pos);
  float *q_s = reinterpret_cast<float *>(scr);             // [G][HD]
  float *sc = q_s + ATTN_MAX_G * 256;                      // [ATTN_TILE][G]  (position-major: one vector per position)
  float *knew = sc + ATTN_MAX_G * ATTN_TILE;               // [2][HD] appended k / v row
  __shared__ float m_run[ATTN_MAX_G], l_run[ATTN_MAX_G], fac[ATTN_MAX_G];
  __shared__ float wgt[ATTN_MAX_G][ATTN_MAX_SPLIT];
  __shared__ int is_last;
  if (!it.active) return;  // CTA-uniform; inactive CTAs issue no stages either

  const T *qkv = reinterpret_cast<const T *>(a.qkv);
  const float *cosr = rope_cs, *sinr = rope_cs + 128;  // this step's cos/sin row, staged once per launch
  T *kc = reinterpret_cast<T *>(L.kc) + ((size_t)it.kvh * a.cap + ws) * HD;
  T *vc = reinterpret_cast<T *>(L.vc) + ((size_t)it.kvh * a.cap + ws) * HD;
  const int kvh = it.kvh, split = it.split, s0 = it.s0, s1 = it.s1, TILE = it.tile;
  const bool owner = (pos >= s0 && pos < s1);

  // q for the G heads of this kv head (qkv was written by other SMs: ld.global.cg)
  for (int g = warp; g < G; g += NW) {
    const T *src = qkv + (size_t)(kvh * G + g) * HD;
    float *dst = q_s + g * HD;
    for (int d = lane; d < HD; d += 32) dst[d] = DT<T>::to_f(ldcg_T<T>(src + d));
    __syncwarp();
    mk_norm_rope<T, HD>(dst, reinterpret_cast<const T *>(L.qn), a.eps, cosr, sinr, a.rot, lane);
  }
  if (owner) {
    if (warp == NW - 1) {
      const T *src = qkv + (size_t)(a.n_heads + kvh) * HD;
      for (int d = lane; d < HD; d += 32) knew[d] = DT<T>::to_f(ldcg_T<T>(src + d));
      __syncwarp();
      mk_norm_rope<T, HD>(knew, reinterpret_cast<const T *>(L.kn), a.eps, cosr, sinr, a.rot, lane);
      for (int d = lane; d < HD; d += 32) kc[(size_t)pos * HD + d] = DT<T>::from_f(knew[d]);
    } else if (warp == NW - 2) {
      const T *vsrc = qkv + (size_t)(a.n_heads + a.n_kv + kvh) * HD;
      for (int d = lane; d < HD; d += 32) {
        const T vv = ldcg_T<T>(vsrc + d);
        vc[(size_t)pos * HD + d] = vv;
        knew[HD + d] = DT<T>::to_f(vv);
      }
    }
  }
  if (ct < ATTN_MAX_G) { m_run[ct] = -INFINITY; l_run[ct] = 0.f; }
  named_bar_sync(1, MK_CT);
  stamp(1);

  const int grp = lane / LPR, gl = lane % LPR;  // row group within the warp / lane within the row
  float qreg[G][8];
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int i = 0; i < 8; i++) qreg[g][i] = q_s[g * HD + gl * 8 + i];
  float acc[G][8];  // PV accumulator: dims gl*8..+8 of every head, over this half-warp's positions
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int i = 0; i < 8; i++) acc[g][i] = 0.f;

  T *Ks = nullptr, *Vs = nullptr;
  int sk = 0, sv = 0;
  for (int t0 = s0; t0 < s1; t0 += TILE) {
    const int tn = min(TILE, s1 - t0);
    if (t0 > s0) {  // release the previous tile's stages (the last tile's are kept for the final reduction)
      __syncwarp();
      if (lane == 0) { mbar_arrive(&rg.empty[sk]); mbar_arrive(&rg.empty[sv]); }
    }
    sk = rg.s;
    const uint32_t phk = rg.ph;
    rg.advance();
    sv = rg.s;
    const uint32_t phv = rg.ph;
    rg.advance();
    Ks = reinterpret_cast<T *>(rg.ring + (size_t)sk * MK_STAGE_BYTES);
    Vs = reinterpret_cast<T *>(rg.ring + (size_t)sv * MK_STAGE_BYTES);
    mbar_wait(&rg.full[sk], phk);
    mbar_wait(&rg.full[sv], phv);
    stamp(2);
    if (owner && pos >= t0 && pos < t0 + tn) {  // drop the appended row into its slot of the staged tiles
      const int slot = pos - t0;
      for (int d = ct; d < HD; d += MK_CT) {
        Ks[(size_t)slot * HD + d] = DT<T>::from_f(knew[d]);
        Vs[(size_t)slot * HD + d] = DT<T>::from_f(knew[HD + d]);
      }
    }
    named_bar_sync(1, MK_CT);
    // ---- scores: s[p][g] = (q_g . k_p) * scale, f32 ----------------------------------------------
    for (int pb = warp * RPWI; pb < tn; pb += NW * RPWI) {  // warp-uniform trip count (shuffles below)
      const int p = pb + grp;
      const bool valid = p < tn;
      float kf[8];
      uint4 kraw = make_uint4(0u, 0u, 0u, 0u);
      if (valid) kraw = *reinterpret_cast<const uint4 *>(Ks + (size_t)p * HD + gl * 8);
      unpack8<T>(kraw, kf);
      float sg[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) sacc = fmaf(qreg[g][i], kf[i], sacc);
        sg[g] = sacc;
      }
      if (LPR == 16) {
        int gh;
        const float v = reduce16_heads<G>(sg, lane, gh);
        constexpr int FIN = (G >= 8) ? 1 : (G == 4 ? 3 : (G == 2 ? 7 : 15));  // lanes that hold a finished head
        if (valid && (gl & FIN) == 0) sc[p * G + gh] = v * a.scale;
      } else {
#pragma unroll
        for (int g = 0; g < G; g++) {
          float v = sg[g];
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (gl == 0 && valid) sc[p * G + g] = v * a.scale;
        }
      }
    }
    named_bar_sync(1, MK_CT);
    stamp(3);
    // ---- online softmax bookkeeping, one warp per head ---------------------------------------------
    for (int g = warp; g < G; g += NW) {
      float mx = -INFINITY;
      for (int p = lane; p < tn; p += 32) mx = fmaxf(mx, sc[p * G + g]);
      mx = warp_max(mx);
      const float m_new = fmaxf(m_run[g], mx);
      float sum = 0.f;
      for (int p = lane; p < tn; p += 32) {
        const float ev = expf(sc[p * G + g] - m_new);
        sc[p * G + g] = ev;
        sum += ev;
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        const float f = (m_run[g] == -INFINITY) ? 0.f : expf(m_run[g] - m_new);
        fac[g] = f;
        l_run[g] = l_run[g] * f + sum;
        m_run[g] = m_new;
      }
    }
    named_bar_sync(1, MK_CT);
    stamp(4);
    // ---- PV: a lane owns dims gl*8..+8 of every head; one 16-byte V load feeds 8*G FMAs ------------
#pragma unroll
    for (int g = 0; g < G; g++) {
      const float f = fac[g];
#pragma unroll
      for (int i = 0; i < 8; i++) acc[g][i] *= f;
    }
    for (int p = warp * RPWI + grp; p < tn; p += NW * RPWI) {
      float vf[8];
      unpack8<T>(*reinterpret_cast<const uint4 *>(Vs + (size_t)p * HD + gl * 8), vf);
#pragma unroll
      for (int g = 0; g < G; g++) {
        const float ev = sc[p * G + g];
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] = fmaf(ev, vf[i], acc[g][i]);
      }
    }
    named_bar_sync(1, MK_CT);  // every warp is done with this tile's K, V and probabilities
    stamp(5);
  }

  // ---- combine the row groups: lanes of a warp first (shuffles), then the 16 warps through the (dead)
  //      K/V stage buffers of the last tile -----------------------------------------------------------
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
    for (int g = 0; g < G; g++)
#pragma unroll
      for (int i = 0; i < 8; i++) acc[g][i] += __shfl_xor_sync(0xffffffffu, acc[g][i], o);
  const bool have_tile = (s1 > s0);
  float *red0 = reinterpret_cast<float *>(Ks), *red1 = reinterpret_cast<float *>(Vs);  // 8 warps each: [8][G][HD] <= 32 KB
  if (have_tile) {
    if (lane < LPR) {
      float *dst = (warp < 8 ? red0 : red1) + (size_t)(warp & 7) * G * HD;
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int i = 0; i < 8; i++) dst[g * HD + lane * 8 + i] = acc[g][i];
    }
    named_bar_sync(1, MK_CT);
    for (int i = ct; i < G * HD; i += MK_CT) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) sum += red0[(size_t)w * G * HD + i];
#pragma unroll
      for (int w = 0; w < 8; w++) sum += red1[(size_t)w * G * HD + i];
      const int g = i / HD, d = i % HD;
      a.ws_acc[((size_t)(kvh * G + g) * a.nsplit + split) * HD + d] = sum;
    }
    __syncwarp();
    named_bar_sync(1, MK_CT);
    if (lane == 0) { mbar_arrive(&rg.empty[sk]); mbar_arrive(&rg.empty[sv]); }
  } else {
    for (int i = ct; i < G * HD; i += MK_CT) {
      const int g = i / HD, d = i % HD;
      a.ws_acc[((size_t)(kvh * G + g) * a.nsplit + split) * HD + d] = 0.f;
    }
  }
  if (ct < G) {
    a.ws_ml[((size_t)(kvh * G + ct) * a.nsplit + split) * 2 + 0] = m_run[ct];
    a.ws_ml[((size_t)(kvh * G + ct) * a.nsplit + split) * 2 + 1] = l_run[ct];
  }
  stamp(6);
  named_bar_sync(1, MK_CT);
  if (ct == 0) {
    // acq_rel at gpu scope: releases the partials every thread of this CTA wrote before the bar.sync (cumulative)
    // and, for the last arriver, acquires the other splits' partials
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&a.attn_counters[kvh]) : "memory");
    is_last = (ticket == (unsigned)a.nsplit - 1);
  }
  named_bar_sync(1, MK_CT);
  stamp(7);
  if (!is_last) return;
  for (int g = warp; g < G; g += NW) {
    const int h = kvh * G + g;
    float m = -INFINITY, l = 0.f;
    if (lane < a.nsplit) {
      m = __ldcg(a.ws_ml + ((size_t)h * a.nsplit + lane) * 2);
      l = __ldcg(a.ws_ml + ((size_t)h * a.nsplit + lane) * 2 + 1);
    }
    const float M = warp_max(m);
    const float w = (m == -INFINITY) ? 0.f : expf(m - M);
    const float Lsum = warp_sum(w * l);
    wgt[g][lane] = w / Lsum;
  }
  named_bar_sync(1, MK_CT);
  T *y = reinterpret_cast<T *>(a.y);
  for (int i = ct; i < G * HD; i += MK_CT) {
    const int g = i / HD, d = i % HD, h = kvh * G + g;
    const float *src = a.ws_acc + (size_t)h * a.nsplit * HD + d;
    float o = 0.f;
#pragma unroll 8
    for (int s2 = 0; s2 < a.nsplit; s2++) o = fmaf(wgt[g][s2], __ldcg(src + (size_t)s2 * HD), o);
    y[(size_t)h * HD + d] = DT<T>::from_f(o);
  }
  if (ct == 0) a.attn_counters[kvh] = 0;
}

// ---------------------------------------------------------------------------------------- the kernel
template <typename T, int HD, int G>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(const MkArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int es = sizeof(T);
  unsigned char *ring = smem_raw;
  T *xs = reinterpret_cast<T *>(ring + (size_t)a.n_stages * MK_STAGE_BYTES);
  size_t off = (size_t)a.n_stages * MK_STAGE_BYTES + mk_xs_bytes(a.max_k, es);
  off = (off + 15) & ~(size_t)15;
  float *partial = reinterpret_cast<float *>(smem_raw + off);
  off += (size_t)a.partial_floats * 4;  // sized by the host: max rows x slices over all GEMV types
  int *loc_row = reinterpret_cast<int *>(smem_raw + off);
  off += (size_t)a.max_groups * 4;
  float *scratch = reinterpret_cast<float *>(smem_raw + off);
  off += 128 * 4;
  off = (off + 7) & ~(size_t)7;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + off);
  uint64_t *empty = full + MK_MAX_STAGES;
  int *stage_row = reinterpret_cast<int *>(empty + MK_MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.n_stages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], MK_CW);
    }
    mbar_fence_init();
  }
  __syncthreads();
  const int pos = *a.d_pos;
  MkRing rg{ring, full, empty, stage_row, a.n_stages, 0, 0u};

  if (warp == MK_CW) {
    // ================= producer: stream weights and K/V tiles through the whole phase list ==========
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first(), pol_kv = policy_evict_last();
      MkPrefetch pf{&a, 0, 0, 0, (int)sizeof(T)};
      if (a.l2_prefetch == 1) for (int i = 0; i < MK_PF_AHEAD; i++) pf.step();
      auto produce = [&](const MkGeom &g, const void *W, int gran, int phase) {
        if (mk_is_dynamic(g)) mk_produce_gemv_dyn<T>(rg, g, W, a.tickets + phase, a.max_groups, pol_w, pf);
        else mk_produce_gemv<T>(rg, g, W, gran, pol_w);
      };
      for (int l = 0; l < a.n_layers; l++) {
        const MkLayer L = a.layers[l];
        produce(a.g_qkv, L.wqkv, 1, 4 * l + 0);
        mk_produce_attn<T>(rg, a, L, pos, pol_kv);
        if (a.l2_prefetch == 2) {
          mk_prefetch_static(a.g_o, L.wo, (int)sizeof(T), 0, 1 << 20);
          mk_prefetch_static(a.g_gu, L.wgu, (int)sizeof(T), 0, 10);
        }
        produce(a.g_o, L.wo, 1, 4 * l + 1);
        produce(a.g_gu, L.wgu, 2, 4 * l + 2);
        produce(a.g_down, L.wd, 1, 4 * l + 3);
      }
      if (a.has_head) produce(a.g_head, a.lm_head, 1, 4 * a.n_layers);
    }
    return;
  }

  // ================= consumers ========================================================================
  const int ct = threadIdx.x;
  __shared__ float rope_cs[256];  // cos (first 128) and sin (last 128) row of this step's position, as fp32
  {
    const int half = a.rot / 2;
    const T *cr = reinterpret_cast<const T *>(a.cos_t) + (size_t)pos * half, *sr = reinterpret_cast<const T *>(a.sin_t) + (size_t)pos * half;
    if (ct < half) rope_cs[ct] = DT<T>::to_f(cr[ct]);
    else if (ct >= 128 && ct - 128 < half) rope_cs[ct] = DT<T>::to_f(sr[ct - 128]);
    // visibility: the first named barrier of the first phase orders these writes before any use
  }
  // gbar[0] is a monotonic arrival counter; gbar[1] holds its value at the start of this launch (written
  // by the previous launch's last thread), so barrier k of this launch completes at start + k*grid.
  const unsigned long long start = *reinterpret_cast<volatile unsigned long long *>(a.gbar + 1);
  unsigned nbar = 0;
  int ntrace = 0;
  auto stamp = [&]() {
    if (a.trace && blockIdx.x == 0 && ct == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[ntrace++] = t;
    }
  };
  auto gsync = [&]() {
    nbar++;
    stamp();
    mk_grid_sync(a.gbar, start + (unsigned long long)nbar * gridDim.x, ct);
    stamp();
  };
  stamp();
  unsigned long long *const st_rec = (a.step_trace && blockIdx.x == 0 && ct == 0)
      ? a.step_trace + ((size_t)((unsigned)*a.d_step % MK_STEP_RING) * 2 + a.trace_tag) * 4 : nullptr;
  if (st_rec) st_rec[0] = mk_gtime();
  if (a.inbox_ctr) {  // ring hand-off: the previous shard pushes our input over NVLink and bumps the counter once per CTA
    if (ct == 0) {
      const unsigned long long want = (*reinterpret_cast<volatile unsigned long long *>(a.ring_seq) + 1ULL) * gridDim.x;
      unsigned long long v;
      do {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a.inbox_ctr) : "memory");
      } while (v < want);
    }
    named_bar_sync(1, MK_CT);
  }
  if (st_rec) st_rec[1] = mk_gtime();
  const T *cur = reinterpret_cast<const T *>(a.x_in);
  if (cur == nullptr) {  // master: the block input is the embedding row of the current token (text_model.rs:271)
    uint32_t tok = *a.d_token;
    if (tok >= (uint32_t)a.vocab) tok = 0;  // memory safety only: host entry points reject out-of-range ids (EINVAL); in-kernel argmax ids are < vocab
    cur = reinterpret_cast<const T *>(a.embed) + (size_t)tok * a.hidden;
  }
  MkEpi e{};
  for (int l = 0; l < a.n_layers; l++) {
    const MkLayer L = a.layers[l];
    T *dst = reinterpret_cast<T *>((l == a.n_layers - 1) ? a.x_out : a.xa);
    // rms_1 + qkv (+bias)
    mk_stage_x<T>(xs, cur, L.ln1, a.hidden, a.eps, scratch, ct, warp, lane);
    e = MkEpi{};
    e.bias = L.bqkv;
    e.out = a.qkv;
    mk_consume_gemv<T, EPI_PLAIN>(rg, a.g_qkv, xs, partial, loc_row, scratch, e, ct, warp, lane);
    gsync();
    // qk-norm, RoPE, KV append, attention
    mk_consume_attn<T, HD, G>(rg, a, L, pos, reinterpret_cast<unsigned char *>(xs), rope_cs, ct, warp, lane,
                           (a.trace && blockIdx.x == 0 && l == 1) ? a.trace + 2048 : nullptr);
    gsync();
    // o_proj + residual
    mk_stage_x<T>(xs, a.y, nullptr, a.n_heads * a.hd, a.eps, scratch, ct, warp, lane);
    e = MkEpi{};
    e.residual = cur;
    e.out = a.xb;
    mk_consume_gemv<T, EPI_RESIDUAL>(rg, a.g_o, xs, partial, loc_row, scratch, e, ct, warp, lane);
    gsync();
    // rms_2 + gate_up + silu*mul
    mk_stage_x<T>(xs, a.xb, L.ln2, a.hidden, a.eps, scratch, ct, warp, lane);
    e = MkEpi{};
    e.out = a.mm;
    e.act = a.act;
    mk_consume_gemv<T, EPI_SWIGLU>(rg, a.g_gu, xs, partial, loc_row, scratch, e, ct, warp, lane);
    if (a.trace && l == 1 && ct == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[2304 + blockIdx.x] = t;
    }
    gsync();
    if (a.trace && l == 1 && ct == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      a.trace[2560 + blockIdx.x] = t;
    }
    // down + residual
    mk_stage_x<T>(xs, a.mm, nullptr, a.inter, a.eps, scratch, ct, warp, lane);
    e = MkEpi{};
    e.residual = a.xb;
    e.out = dst;
    mk_consume_gemv<T, EPI_RESIDUAL>(rg, a.g_down, xs, partial, loc_row, scratch, e, ct, warp, lane);
    if (l < a.n_layers - 1 || a.has_head) gsync();
    cur = dst;
  }
  if (a.has_head) {
    mk_stage_x<T>(xs, cur, a.ln_f, a.hidden, a.eps, scratch, ct, warp, lane);
    e = MkEpi{};
    e.out = a.logits;
    e.part_val = a.part_val;
    e.part_idx = a.part_idx;
    e.counter = a.argmax_counter;
    e.token_out = a.token_out;
    e.token_ring = a.token_ring;
    e.step = a.d_step;
    e.ring_cap = a.ring_cap;
    e.advance = a.advance;
    e.d_pos = a.d_pos;
    e.d_step = a.d_step;
    e.gbar = a.gbar;
    e.next_start = start + (unsigned long long)nbar * gridDim.x;
    e.ring_seq = a.inbox_ctr ? a.ring_seq : nullptr;
    mk_consume_gemv<T, EPI_ARGMAX>(rg, a.g_head, xs, partial, loc_row, scratch, e, ct, warp, lane);
  }
  if (a.peer_ctr && !a.has_head) {  // our rows of x_out are in the next shard's inbox: release them, one arrival per CTA
    named_bar_sync(1, MK_CT);
    if (ct == 0) {
      __threadfence_system();
      asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(a.peer_ctr), "l"(1ULL) : "memory");
    }
  }
  if (st_rec) st_rec[2] = mk_gtime();
  // bookkeeping without a head: CTA 0 only gets here after passing barriers that every CTA arrived at,
  // and every CTA read *d_pos, *ring_seq and the barrier start value before its first barrier.
  if (!a.has_head && blockIdx.x == 0 && ct == 0) {
    if (a.advance) { *a.d_pos += 1; *a.d_step += 1; }
    if (a.inbox_ctr) *a.ring_seq += 1ULL;
    a.gbar[1] = start + (unsigned long long)nbar * gridDim.x;
  }
}

}  // namespace cake
