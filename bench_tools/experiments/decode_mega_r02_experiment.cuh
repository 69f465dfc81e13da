// EXPERIMENT RECORD — NOT COMPILED INTO THE PRODUCT (kept for the evidence trail of round 2; see profiles/README.md
// "Round 2: what was tried on decode_mega_kernel" and DESIGN.md §4.7).  This is the decode megakernel as rewritten in
// round 2: software-pipelined GEMV consumer loop with a batched 4-stage butterfly reduction, setmaxnreg register split
// (20 warps: 4 consumer warpgroups at 104-112 registers, producer warpgroup at 32-64), one phase loop instead of five
// inlined GEMV copies (20 k -> 10.5 k instructions), an attention phase with one CTA barrier per tile, residual staging,
// an L2 prefetch helper warp (removed again).  Measured on B200 against the round-1 kernel that ships
// (cake_b200/csrc/decode_mega.cuh), same box, same run:
//     Llama-3-8B  bs=1 2k ctx:  341.5 tok/s (round-1 kernel)  vs  336.4 (this file, best configuration)
//     Llama-3-70B bs=1 2k ctx:   38.3 tok/s (round-1 kernel)  vs   25.5 (this file)
// The faster consumer loop is real (+7 % against the generic loop inside this file) but the phases are bounded by data
// arrival, not by the consumers, and the restructuring around it cost more than the loop gained.
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

// ---- build switches (defaults = the product; the others exist for profiling / A-B builds, bench_tools/) ----
#ifndef MK_TRACE
#define MK_TRACE 0
#endif
#ifndef MK_RES_STAGE
#define MK_RES_STAGE 0   // stage the epilogue's residual vector with x + preload norm weights before the barrier (A/B r02: -3 %)
#endif
#ifndef MK_SETMAXNREG
#define MK_SETMAXNREG 1
#endif
#ifndef MK_LS_CACHE
#define MK_LS_CACHE 1
#endif
#ifndef MK_ATTN_V1
#define MK_ATTN_V1 1  // A/B r02 (profiles/): the round-1 attention phase is still 4 % faster end to end than the rewrite below
#endif

namespace cake {

constexpr int MK_CW = 16;                       // consumer warps
constexpr int MK_CT = MK_CW * 32;               // consumer threads
#if MK_SETMAXNREG
constexpr int MK_THREADS = (MK_CW + 4) * 32;    // + one warpgroup: producer warp 16 (warps 17-19 only take part in the register hand-over)
#else
constexpr int MK_THREADS = (MK_CW + 1) * 32;    // + producer warp
#endif
// Register split (setmaxnreg, per warpgroup): 20 warps launch at 96 registers (5 per SM sub-partition x 96 x 32 <= 16 K).
// The CTA's pool is what it was launched with — 640 x 96 = 61 440 registers: the producer warpgroup shrinks to
// MK_REGS_PRODUCER and the four consumer warpgroups grow to MK_REGS_CONSUMER, 128 x 64 + 512 x 104 = 61 440 exactly
// (a request beyond the pool blocks forever), which keeps the pipelined GEMV loop free of spills.
#ifndef MK_REGS_P
#define MK_REGS_P 64
#define MK_REGS_C 104
#endif
static_assert(!MK_SETMAXNREG || 128 * MK_REGS_P + 512 * MK_REGS_C <= MK_THREADS * 96, "setmaxnreg targets exceed the CTA's register pool");
constexpr int MK_REGS_PRODUCER = MK_REGS_P;
constexpr int MK_REGS_CONSUMER = MK_REGS_C;
constexpr int MK_STAGE_BYTES = 32768;
constexpr int MK_MAX_STAGES = 6;
constexpr int MK_MAX_LAYERS = 96;

struct MkLayer {
  const void *wqkv, *wo, *wgu, *wd, *ln1, *ln2, *bqkv, *qn, *kn;
  void *kc, *vc;
};
struct MkGeom {  // one decode GEMV type; a stage holds RS row segments of KC columns
  int N, K, KC, RS, WPR, RPW;
  int no_fast;  // A/B aid (CAKE_B200_NO_FAST=1): take the generic consumer loop
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
  unsigned long long *trace;  // optional: %globaltimer stamps of CTA 0 at phase boundaries (profiling aid)
  // always on (3 stores per launch): %globaltimer of CTA 0 at kernel entry / input acquired / exit, ring of
  // MK_STEP_RING steps x 2 launches (tag 0: the layer launch, tag 1: rank 0's head-only launch when sharded) x 4 u64
  unsigned long long *step_trace;
  int trace_tag;
};
constexpr int MK_STEP_RING = 2048;

__host__ __device__ inline size_t mk_xs_bytes(int max_k, int hidden, int es) {
  size_t b = (size_t)max_k * es + (size_t)hidden * es;  // x of the widest phase + the residual vector of the phase's epilogue
  // attention scratch: two score buffers [ATTN_TILE][G], q [G][HD <= 128] and the appended k row, all fp32
  const size_t attn = (size_t)2 * ATTN_MAX_G * ATTN_TILE * 4 + (size_t)ATTN_MAX_G * 128 * 4 + (size_t)128 * 4 + 1024;
  return b > attn ? b : attn;  // the attention phase overlays its scratch on the x buffer
}
__host__ __device__ inline size_t mk_smem_bytes(int max_k, int hidden, int partial_floats, int max_groups, int n_stages, int es) {
  size_t off = (size_t)n_stages * MK_STAGE_BYTES;
  off += mk_xs_bytes(max_k, hidden, es);
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

// ---- fine-grained phase trace (profiling builds only: nvcc -DMK_TRACE=1 -> libcake_b200_trace.so, bench_tools/mega_trace.py).
// One 16-u64 record per (CTA, phase 0..4 = qkv/attn/o/gate_up/down) of layer MK_TRACE_LAYER at a.trace[4096 + ...]:
//   0 consumer enters the phase   1 x staged   2 last stage consumed   3 arrives at the grid barrier   4 passes it
//   5 clocks warp 0 spent waiting for full stages (starved)   6 stages consumed
//   7 producer's first issue   8 static groups issued   9 pool groups issued (end marker)
//   10 clocks the producer spent waiting for an empty stage (ring full)   11 pool groups claimed
//   12..15 attention: q/k/v prepared, tiles done, partials written, ticket taken
constexpr int MK_TRACE_LAYER = 1;
constexpr int MK_TRACE_BASE = 4096;
constexpr int MK_TRACE_U64 = MK_TRACE_BASE + 160 * 5 * 16;
__device__ __forceinline__ unsigned long long mk_gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#if MK_TRACE
#define MKT(stmt) do { stmt; } while (0)
#else
#define MKT(stmt) do { } while (0)
#endif
__device__ __forceinline__ unsigned long long *mk_trec(unsigned long long *trace, int layer, int phase) {
  return (trace && layer == MK_TRACE_LAYER) ? trace + MK_TRACE_BASE + ((size_t)blockIdx.x * 5 + phase) * 16 : nullptr;
}

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
  sp.sg = (int)(((long)sp.n_groups * 3 / 4) / gridDim.x);  // (7/8 static for the short phases: measured worse, r02 call 13)
  sp.pool_start = sp.sg * gridDim.x;
  return sp;
}
__device__ __forceinline__ bool mk_is_dynamic(const MkGeom &g) { return g.K == g.KC && (g.KC / 8) / g.WPR <= 128; }

template <typename T>
__device__ __forceinline__ void mk_produce_gemv_dyn(MkRing &rg, const MkGeom &g, const void *W, unsigned *ticket, int max_groups,
                                                    uint64_t pol, unsigned long long *tr) {
  constexpr int es = sizeof(T);
  const MkSplit sp = mk_split(g);
  const size_t rowb = (size_t)g.K * es;
  const unsigned char *Wb = reinterpret_cast<const unsigned char *>(W);
  int issued = 0;
  (void)tr;
  MKT(if (tr) tr[7] = mk_gtime());
#if MK_TRACE
  long long blocked = 0;
#endif
  auto issue = [&](int grp) {
    const int row = grp * g.RS, nr = min(g.RS, g.N - row);
#if MK_TRACE
    const long long c0 = clock64();
#endif
    mbar_wait(&rg.empty[rg.s], rg.ph ^ 1u);
#if MK_TRACE
    blocked += clock64() - c0;
#endif
    rg.stage_row[rg.s] = row;
    mbar_arrive_expect_tx(&rg.full[rg.s], (uint32_t)(nr * rowb));
    bulk_g2s(rg.ring + (size_t)rg.s * MK_STAGE_BYTES, Wb + (size_t)row * rowb, (uint32_t)(nr * rowb), &rg.full[rg.s], pol);
    rg.advance();
    issued++;
  };
  for (int grp = blockIdx.x * sp.sg; grp < (int)(blockIdx.x + 1) * sp.sg; grp++) {
    issue(grp);
  }
  MKT(if (tr) tr[8] = mk_gtime());
#if MK_TRACE
  const int n_static = issued;
#endif
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
  MKT(if (tr) { tr[9] = mk_gtime(); tr[10] = (unsigned long long)blocked; tr[11] = (unsigned long long)(issued - n_static); });
}

template <typename T>
__device__ __forceinline__ void mk_produce_gemv(MkRing &rg, const MkGeom &g, const void *W, int G, uint64_t pol, unsigned long long *tr) {
  constexpr int es = sizeof(T);
  int r0, r1;
  mk_rows(g.N, G, r0, r1);
  (void)tr;
  MKT(if (tr) tr[7] = mk_gtime());
#if MK_TRACE
  long long blocked = 0;
#endif
  const int nchunk = g.K / g.KC;
  const size_t seg = (size_t)g.KC * es;
  const unsigned char *Wb = reinterpret_cast<const unsigned char *>(W);
  for (int row = r0; row < r1; row += g.RS) {
    const int nr = min(g.RS, r1 - row);
    for (int j = 0; j < nchunk; j++) {
#if MK_TRACE
      const long long c0 = clock64();
#endif
      mbar_wait(&rg.empty[rg.s], rg.ph ^ 1u);
#if MK_TRACE
      blocked += clock64() - c0;
#endif
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
  MKT(if (tr) { tr[8] = tr[9] = mk_gtime(); tr[10] = (unsigned long long)blocked; tr[11] = 0; });
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
  const MkAttnItem it = mk_attn_item<T>(a, pos);
  const int HD = a.hd;
  const T *kc = reinterpret_cast<const T *>(L.kc) + (size_t)it.kvh * a.cap * HD;
  const T *vc = reinterpret_cast<const T *>(L.vc) + (size_t)it.kvh * a.cap * HD;
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
// `res` (nullable): the residual vector of this phase's epilogue, H elements, copied to res_s in the same L2 round trip
// (the epilogue's own ld.cg of the residual cost 1.5-2 us under the weight stream, trace r02).
// `pre_w` (nullable): this thread's norm-weight vectors (v = ct, ct + MK_CT), loaded BEFORE the preceding grid barrier.
template <typename T>
__device__ __forceinline__ void mk_stage_x(T *xs, const void *x, const void *norm_w, int K, float eps, float *scratch,
                                           int ct, int warp, int lane, T *res_s, const void *res, int H, const uint4 *pre_w) {
  constexpr int es = sizeof(T);
  uint4 *xsv = reinterpret_cast<uint4 *>(xs);
  const int nv = K * es / 16;
  const char *xb = reinterpret_cast<const char *>(x);
  if (res) {
    const int nr = H * es / 16;
    uint4 *rsv = reinterpret_cast<uint4 *>(res_s);
    const char *rb = reinterpret_cast<const char *>(res);
    for (int v = ct; v < nr; v += MK_CT) rsv[v] = ldcg_v4(rb + (size_t)v * 16);
  }
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
    int j = 0;
    for (int v = ct; v < nv; v += MK_CT, j++) {
      float f[8], w8[8];
      unpack8<T>(xsv[v], f);
      unpack8<T>((pre_w && j < 2) ? pre_w[j] : wg[v], w8);
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

struct MkRows {  // which rows this CTA reduced in a GEMV phase (their slice sums are in `partial`)
  int dyn, r0, nrows, nloc;
};
// The phase-independent part of a decode GEMV: slice sums of the staged rows against x (shared), into `partial`.
// ONE copy of this code serves every phase (the kernel walks a phase table): with one inlined copy per phase the
// kernel was 20 k instructions and 11 % of all stall samples were instruction-fetch misses at phase starts (ncu r02a).
template <typename T>
__device__ __forceinline__ MkRows mk_gemv_core(MkRing &rg, const MkGeom &g, int gran, const T *xs, float *partial, int *loc_row,
                                              int ct, int warp, int lane, unsigned long long *tr) {
  (void)tr;
#if MK_TRACE
  long long starved = 0;
  int nst = 0;
#endif
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
    // one 16-byte weight vector against the lane's x slice i: 8 converts + 8 FMA into 4 independent accumulators
#define MK_FMA8(w_, i_)                                                                                             \
    do {                                                                                                            \
      float wf[8];                                                                                                  \
      unpack8<T>(w_, wf);                                                                                           \
      a0 = fmaf(wf[0], xr[i_][0], a0); a1 = fmaf(wf[1], xr[i_][1], a1); a2 = fmaf(wf[2], xr[i_][2], a2); a3 = fmaf(wf[3], xr[i_][3], a3); \
      a0 = fmaf(wf[4], xr[i_][4], a0); a1 = fmaf(wf[5], xr[i_][5], a1); a2 = fmaf(wf[6], xr[i_][6], a2); a3 = fmaf(wf[7], xr[i_][7], a3); \
    } while (0)
    if (RPW == 1 && nvec > 96 && !g.no_fast) {
      // ---- fast path (every production shape: one row slice of 97..128 vectors per warp per stage) -------------
      // The round-1 loop was latency-bound, not HBM-bound (trace r02: 0.5-0.66 us of consumer time per 32 KB stage
      // with < 10 % of it spent waiting for data; SASS: the four LDS.128 of a row went through ONE register quad, and
      // every stage ended in a dependent 5-step shuffle tree).  Here (a) the next stage's four vectors are loaded
      // BEFORE the current stage's FMAs are issued (two register sets, software pipelining across stages), (b) a
      // stage is released as soon as its data is in registers, and (c) the lane partials of four consecutive
      // stages are reduced together by a halving butterfly: 6 shuffles per 4 row slices instead of 20, off the
      // per-stage critical path.  The summation order is fixed by (lane, stage-in-batch): deterministic.
      const uint4 *const ring_v = reinterpret_cast<const uint4 *>(rg.ring) + (slot * segv + ks * nvec + lane);
      constexpr int STAGE_V = MK_STAGE_BYTES / 16;
      auto wait_full = [&](int s_, uint32_t ph_) {
#if MK_TRACE
        const long long c0 = clock64();
#endif
        mbar_wait(&rg.full[s_], ph_);
#if MK_TRACE
        starved += clock64() - c0;
        nst++;
#endif
      };
      uint4 w0 = z, w1 = z, w2 = z, w3 = z;  // the lane's four weight vectors of the stage being reduced
      int cur_s = rg.s;
      wait_full(cur_s, rg.ph);
      int row = rg.stage_row[cur_s];
      if (row >= 0) {
        const uint4 *rowp = ring_v + (size_t)cur_s * STAGE_V;
        w0 = rowp[0]; w1 = rowp[32]; w2 = rowp[64]; w3 = rowp[96];  // [96] may lie past the slice: masked by v3 below
      }
      const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
      const int my_j = (up16 ? 2 : 0) + (up8 ? 1 : 0);  // which stage of a batch this lane ends up holding
      while (row >= 0) {
        float vb[4] = {0.f, 0.f, 0.f, 0.f};
        int cnt = 0;
        const int base = nloc;
#pragma unroll
        for (int j = 0; j < 4; j++) {
          if (row >= 0) {
            rg.advance();
            const int nxt_s = rg.s;
            wait_full(nxt_s, rg.ph);
            const int row_n = rg.stage_row[nxt_s];
            // rotating prefetch: vector i of the NEXT stage is loaded right after the last use of vector i of this one,
            // so its latency hides behind the FMAs of the other three vectors and the stage hand-over
            const uint4 *rowp = ring_v + (size_t)nxt_s * STAGE_V;
            const bool more = row_n >= 0;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            MK_FMA8(w0, 0);
            if (more) w0 = rowp[0];
            MK_FMA8(w1, 1);
            if (more) w1 = rowp[32];
            MK_FMA8(w2, 2);
            if (more) w2 = rowp[64];
            if (!v3) w3 = z;
            MK_FMA8(w3, 3);
            if (more) w3 = rowp[96];
            vb[j] = (a0 + a1) + (a2 + a3);
            __syncwarp();                          // every lane holds its share of stage cur_s in registers / is done with it
            if (lane == 0) mbar_arrive(&rg.empty[cur_s]);
            if (ct == 0) loc_row[nloc] = row;
            nloc++;
            cnt++;
            row = row_n;
            cur_s = nxt_s;
          }
        }
        // batched reduction: after the xor-16 and xor-8 exchanges each lane owns ONE stage's partial
        const float k0 = (up16 ? vb[2] : vb[0]) + __shfl_xor_sync(0xffffffffu, up16 ? vb[0] : vb[2], 16);
        const float k1 = (up16 ? vb[3] : vb[1]) + __shfl_xor_sync(0xffffffffu, up16 ? vb[1] : vb[3], 16);
        float k = (up8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        if ((lane & 7) == 0 && my_j < cnt) partial[((size_t)(base + my_j) * RS + slot) * WPR + ks] = k;
      }
      // cur_s is the end-of-phase marker stage
      __syncwarp();
      if (lane == 0) mbar_arrive(&rg.empty[cur_s]);
      rg.advance();
    } else {
    while (true) {
#if MK_TRACE
      const long long c0 = clock64();
#endif
      mbar_wait(&rg.full[rg.s], rg.ph);
#if MK_TRACE
      starved += clock64() - c0;
      nst++;
#endif
      const int row = rg.stage_row[rg.s];
      if (row >= 0) {
        const uint4 *st = reinterpret_cast<const uint4 *>(rg.ring + (size_t)rg.s * MK_STAGE_BYTES) + ks * nvec + lane;
        for (int r = 0; r < RPW; r++) {
          const uint4 *rowp = st + (size_t)(slot + r * slots) * segv;
          const uint4 w0 = v0 ? rowp[0] : z, w1 = v1 ? rowp[32] : z, w2 = v2 ? rowp[64] : z, w3 = v3 ? rowp[96] : z;
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          MK_FMA8(w0, 0);
          MK_FMA8(w1, 1);
          MK_FMA8(w2, 2);
          MK_FMA8(w3, 3);
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
    }
#undef MK_FMA8
  } else {
    int r1;
    mk_rows(N, gran, r0, r1);
    nrows = r1 - r0;
    for (int g0 = 0; g0 < nrows; g0 += RS) {
      float acc[4][2];
#pragma unroll
      for (int r = 0; r < 4; r++) acc[r][0] = acc[r][1] = 0.f;
      for (int j = 0; j < nchunk; j++) {
#if MK_TRACE
        const long long c0 = clock64();
#endif
        mbar_wait(&rg.full[rg.s], rg.ph);
#if MK_TRACE
        starved += clock64() - c0;
        nst++;
#endif
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
  MKT(if (tr && ct == 0) { tr[2] = mk_gtime(); tr[5] = (unsigned long long)starved; tr[6] = (unsigned long long)nst; });
  named_bar_sync(1, MK_CT);
  return MkRows{dyn ? 1 : 0, r0, nrows, nloc};
}

template <typename T, int EPI>
__device__ __forceinline__ void mk_gemv_epilogue(const MkGeom &g, const MkRows &rw, const float *partial, const int *loc_row,
                                                 float *scratch, const MkEpi &e, int ct, int warp, int lane) {
  const int RS = g.RS, WPR = g.WPR, N = g.N;
  const bool dyn = rw.dyn != 0;
  const int r0 = rw.r0, nrows = rw.nrows, nloc = rw.nloc;
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
    const T *res = reinterpret_cast<const T *>(e.residual);  // staged in shared memory by mk_stage_x
    for (int li = ct; li < n_local; li += MK_CT) {
      const int row = row_of(li);
      if (row >= N) continue;
      const float v = row_sum(li);
      out[row] = DT<T>::from_f(v + DT<T>::to_f(MK_RES_STAGE ? res[row] : ldcg_T<T>(res + row)));
    }
  } else if (EPI == EPI_SWIGLU) {
    for (int p = ct; p < n_local / 2; p += MK_CT) {
      const int row = row_of(2 * p);
      if (row >= N) continue;
      const float gte = row_sum(2 * p), up = row_sum(2 * p + 1);
      const float sl = rnd<T>(gte / (1.0f + expf(-gte)));
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

// One head vector's 8 consecutive dims [d0, d0+8) for this lane, straight from the qkv row in global memory (written by
// other SMs in the previous phase: ld.global.cg), with the reference's per-op D roundings (attention.rs:202-253):
// optional per-head RMS norm over HD (the 16/8/2 lanes of a row group reduce the sum of squares by shuffles), then
// rotate-half RoPE.  RoPE pairs dim d with d +/- rot/2, which another lane owns: the partner's 8 values are simply
// loaded as well (one more 16-byte L2 hit) instead of being exchanged.  cs/sn: this step's cos/sin row (fp32 copies of D).
template <typename T, int HD>
__device__ __forceinline__ void mk_head_dims(const T *vec, const T *norm_w, float eps, const float *cs, const float *sn, int rot,
                                             int d0, bool rope, float (&out)[8]) {
  // called by the first LPR lanes of a warp (one 16-byte slice each); the norm's shuffles stay within those lanes
  constexpr int LPR = HD / 8;
  const int half = rot / 2;
  float own[8], part[8];
  unpack8<T>(ldcg_v4(vec + d0), own);
  // partner block: dims d0 + half (d0 < half), d0 - half (half <= d0 < rot); contiguous iff half % 8 == 0
  const bool lo = d0 < half, in_rot = rope && d0 < rot;
  const int p0 = lo ? d0 + half : d0 - half;
  if (in_rot) {
    if ((half & 7) == 0) {
      unpack8<T>(ldcg_v4(vec + p0), part);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) {  // rot/2 not a multiple of 8 (tiny test shapes): element-wise partner
        const int d = d0 + i;
        const int pd = d < half ? d + half : (d < rot ? d - half : d);
        part[i] = DT<T>::to_f(ldcg_T<T>(vec + pd));
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) part[i] = 0.f;
  }
  if (norm_w) {
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) ss += own[i] * own[i];
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ss += __shfl_xor_sync((LPR == 32) ? 0xffffffffu : ((1u << LPR) - 1u), ss, o);
    const float inv = 1.0f / sqrtf(ss / (float)HD + eps);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      own[i] = rnd<T>(own[i] * inv * DT<T>::to_f(norm_w[d0 + i]));
      if (in_rot) {
        const int d = d0 + i;
        const int pd = ((half & 7) == 0) ? p0 + i : (d < half ? d + half : (d < rot ? d - half : d));
        part[i] = rnd<T>(part[i] * inv * DT<T>::to_f(norm_w[pd]));
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int d = d0 + i;
    if (rope && d < rot) {
      const bool first = d < half;
      const int ci = first ? d : d - half;
      const float c = cs[ci], sv = sn[ci];
      // i < half: o = rnd(rnd(x1 c) - rnd(x2 s)) with x1 = own, x2 = partner;  i >= half: o = rnd(rnd(x2 c) + rnd(x1 s)) with x2 = own
      out[i] = first ? rnd<T>(rnd<T>(own[i] * c) - rnd<T>(part[i] * sv)) : rnd<T>(rnd<T>(own[i] * c) + rnd<T>(part[i] * sv));
    } else {
      out[i] = own[i];
    }
  }
}

// Sum G per-lane values over the LPR lanes of a row group; every lane of the group gets all G sums.
template <int G, int LPR>
__device__ __forceinline__ void group_allreduce(float (&s)[G]) {
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
}

// Attention phase for this CTA's (kv head, split) item — same math as attn_decode_kernel (f32 scores, softmax and PV,
// one rounding of the result); K/V tiles come from the ring, prefetched by the producer while the qkv GEMV was still
// running.  G = n_heads / n_kv_heads is a template parameter so that all per-head state lives in registers.
//
// Round-2 structure (trace r02: the round-1 phase spent 7 CTA-wide barriers and a one-warp-per-head softmax around
// ~1 us of arithmetic; phase span 12.5 us for 8.4 MB):
//   * no q staging: every lane loads its 8 dims of the G query heads (and its RoPE partner dims) straight from L2 and
//     applies QK-norm + RoPE in registers; the appended K/V row lives in the registers of the row group that owns
//     its slot (which also writes it to the cache) — no shared-memory round trip, no barrier;
//   * per tile ONE barrier: scores -> sc (shared) | barrier | every warp derives the tile maximum itself (4 LDS.128 +
//     shuffles), probabilities are exp'ed on the fly by the PV lanes, the denominators stay lane-local;
//   * two score buffers alternate, so a following tile needs no second barrier;
//   * end: warp-local shuffle merge, one cross-warp reduction through the (dead) K/V stages, ticket, last split merges.
template <typename T, int HD, int G>
__device__ __forceinline__ void mk_consume_attn(MkRing &rg, const MkArgs &a, const MkLayer &L, int pos, unsigned char *scr,
                                                const float *rope_cs, int ct, int warp, int lane, unsigned long long *tr) {
  auto stamp = [&](int i) {  // profiling builds: tr[12..15] = q/k/v prepared, tiles done, partials written, ticket taken
    (void)i;
    MKT(if (tr && ct == 0 && i >= 12) tr[i] = mk_gtime());
  };
  constexpr int LPR = HD / 8;            // lanes per cached row (16 B each); HD in {16,64,128} -> 2, 8, 16
  constexpr int RPWI = 32 / LPR;         // rows per warp per iteration
  // More than 4 query heads per kv head (Llama-3-70B: 8) would need 2 x 8 x 8 accumulator / q registers per lane: the 16
  // warps split into G/4 HEAD GROUPS instead, each walking all positions of the tile for its HG = 4 heads.
  constexpr int HG = (G > 4) ? 4 : G;    // heads per lane
  constexpr int NHG = G / HG;            // head groups
  constexpr int NW = MK_CW / NHG;        // warps per head group
  static_assert(HD == 16 || HD == 64 || HD == 128, "head_dim");
  static_assert(G % HG == 0 && MK_CW % NHG == 0, "query heads per kv head");
  const int hg = warp / NW, wg = warp % NW;  // this warp's head group / its index within the group
  const int h0 = hg * HG;                    // first head (within the kv head's G) of this warp
  const MkAttnItem it = mk_attn_item<T>(a, pos);
  float *sc0 = reinterpret_cast<float *>(scr);             // [ATTN_TILE][G] scores, two buffers (position-major)
  float *sc1 = sc0 + ATTN_MAX_G * ATTN_TILE;
  __shared__ float l_warp[MK_CW][4], m_warp[MK_CW][4];
  __shared__ float wgt[ATTN_MAX_G][ATTN_MAX_SPLIT];
  __shared__ int is_last;
  if (!it.active) return;  // CTA-uniform; inactive CTAs issue no stages either

  const T *qkv = reinterpret_cast<const T *>(a.qkv);
  const float *cosr = rope_cs, *sinr = rope_cs + 128;  // this step's cos/sin row, staged once per launch
  T *kc = reinterpret_cast<T *>(L.kc) + (size_t)it.kvh * a.cap * HD;
  T *vc = reinterpret_cast<T *>(L.vc) + (size_t)it.kvh * a.cap * HD;
  const int kvh = it.kvh, split = it.split, s0 = it.s0, s1 = it.s1, TILE = it.tile;
  const bool owner = (pos >= s0 && pos < s1);
  const int grp = lane / LPR, gl = lane % LPR;  // row group within the warp / lane within the row
  const int d0 = gl * 8;

  // ---- q (G heads) and, on the owning split, the appended k row: one warp per vector loads it ONCE per CTA (every lane
  //      fetching its own copy from L2 made 590 k requests hammer the same ~100 cache lines: 8 us, trace r02b), applies
  //      QK-norm + RoPE in registers and parks the result in shared memory as fp32
  float *q_s = sc1 + ATTN_MAX_G * ATTN_TILE;   // [G][HD]
  float *knew_s = q_s + ATTN_MAX_G * HD;       // [HD]
  if (warp < G) {
    if (lane < LPR) {
      float o[8];
      mk_head_dims<T, HD>(qkv + (size_t)(kvh * G + warp) * HD, reinterpret_cast<const T *>(L.qn), a.eps, cosr, sinr, a.rot, lane * 8, true, o);
      *reinterpret_cast<float4 *>(q_s + warp * HD + lane * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4 *>(q_s + warp * HD + lane * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  } else if (warp == G && owner) {
    if (lane < LPR) {
      float o[8];
      mk_head_dims<T, HD>(qkv + (size_t)(a.n_heads + kvh) * HD, reinterpret_cast<const T *>(L.kn), a.eps, cosr, sinr, a.rot, lane * 8, true, o);
      *reinterpret_cast<float4 *>(knew_s + lane * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4 *>(knew_s + lane * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
  named_bar_sync(1, MK_CT);
  float qreg[HG][8];
#pragma unroll
  for (int g = 0; g < HG; g++) {
    const float4 lo4 = *reinterpret_cast<const float4 *>(q_s + (h0 + g) * HD + d0), hi4 = *reinterpret_cast<const float4 *>(q_s + (h0 + g) * HD + d0 + 4);
    qreg[g][0] = lo4.x; qreg[g][1] = lo4.y; qreg[g][2] = lo4.z; qreg[g][3] = lo4.w;
    qreg[g][4] = hi4.x; qreg[g][5] = hi4.y; qreg[g][6] = hi4.z; qreg[g][7] = hi4.w;
  }
  stamp(12);

  float acc[HG][8];  // PV accumulator: dims d0..d0+8 of this warp's heads, over this row group's positions
  float m_run[HG], l_loc[HG];
#pragma unroll
  for (int g = 0; g < HG; g++) {
    m_run[g] = -INFINITY;
    l_loc[g] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) acc[g][i] = 0.f;
  }

  T *Ks = nullptr, *Vs = nullptr;
  int sk = 0, sv = 0, tile_no = 0;
  for (int t0 = s0; t0 < s1; t0 += TILE, tile_no++) {
    const int tn = min(TILE, s1 - t0);
    float *sc = (tile_no & 1) ? sc1 : sc0;
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
    const int new_slot = (owner && pos >= t0 && pos < t0 + tn) ? pos - t0 : -1;  // the appended row's slot in this tile
    // ---- scores: s[p][g] = (q_g . k_p) * scale, f32 ----------------------------------------------
    for (int pb = wg * RPWI; pb < tn; pb += NW * RPWI) {  // warp-uniform trip count (shuffles below)
      const int p = pb + grp;
      const bool valid = p < tn;
      float kf[8];
      {
        uint4 kraw = make_uint4(0u, 0u, 0u, 0u);
        if (valid && p != new_slot) kraw = *reinterpret_cast<const uint4 *>(Ks + (size_t)p * HD + d0);
        unpack8<T>(kraw, kf);
      }
      if (p == new_slot) {  // the appended row: k from the prep step (shared), appended to the cache by this row group
        const float4 lo4 = *reinterpret_cast<const float4 *>(knew_s + d0), hi4 = *reinterpret_cast<const float4 *>(knew_s + d0 + 4);
        kf[0] = lo4.x; kf[1] = lo4.y; kf[2] = lo4.z; kf[3] = lo4.w; kf[4] = hi4.x; kf[5] = hi4.y; kf[6] = hi4.z; kf[7] = hi4.w;
        if (hg == 0) {  // cache.rs:184-210: 16 bytes per lane; v is appended unchanged
          uint4 o;
          o.x = pack2<T>(kf[0], kf[1]); o.y = pack2<T>(kf[2], kf[3]); o.z = pack2<T>(kf[4], kf[5]); o.w = pack2<T>(kf[6], kf[7]);
          *reinterpret_cast<uint4 *>(kc + (size_t)pos * HD + d0) = o;
          *reinterpret_cast<uint4 *>(vc + (size_t)pos * HD + d0) = ldcg_v4(qkv + (size_t)(a.n_heads + a.n_kv + kvh) * HD + d0);
        }
      }
      float sg[HG];
#pragma unroll
      for (int g = 0; g < HG; g++) {
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) sacc = fmaf(qreg[g][i], kf[i], sacc);
        sg[g] = sacc;
      }
      if (LPR == 16) {
        int gh;
        const float v = reduce16_heads<HG>(sg, lane, gh);
        constexpr int FIN = (HG == 4 ? 3 : (HG == 2 ? 7 : 15));  // lanes that hold a finished head
        if (valid && (gl & FIN) == 0) sc[p * G + h0 + gh] = v * a.scale;
      } else {
#pragma unroll
        for (int g = 0; g < HG; g++) {
          float v = sg[g];
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (gl == 0 && valid) sc[p * G + h0 + g] = v * a.scale;
        }
      }
    }
    named_bar_sync(1, MK_CT);  // the tile's scores are complete
    // ---- tile maximum per head, derived by every warp on its own (no serial softmax step) ----------
    {
      float mx[HG];
#pragma unroll
      for (int g = 0; g < HG; g++) mx[g] = -INFINITY;
      for (int p = lane; p < tn; p += 32) {
#pragma unroll
        for (int g = 0; g < HG; g++) mx[g] = fmaxf(mx[g], sc[p * G + h0 + g]);
      }
#pragma unroll
      for (int g = 0; g < HG; g++) {
        const float m_new = fmaxf(m_run[g], warp_max(mx[g]));
        const float fac = (m_run[g] == -INFINITY) ? 0.f : expf(m_run[g] - m_new);
        m_run[g] = m_new;
        l_loc[g] *= fac;
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] *= fac;
      }
    }
    mbar_wait(&rg.full[sv], phv);
    // ---- PV: a lane owns dims d0..d0+8 of its heads; one 16-byte V load feeds 8*HG FMAs.  The tile's probabilities of
    //      a row group (PPG positions x HG heads) are exp'ed ONCE, one per lane, and handed round by shuffles.
    constexpr int PPG = (ATTN_TILE + MK_CW / NHG * RPWI - 1) / (MK_CW / NHG * RPWI);  // positions per row group per tile
    constexpr bool SHARE_EXP = (PPG * HG <= LPR);
    float ev_mine = 0.f;
    if (SHARE_EXP) {
      const int pi = gl / HG, g = gl % HG;          // this lane's (position index, head) pair
      const int p = wg * RPWI + grp + pi * NW * RPWI;
      float m_g = m_run[0];
#pragma unroll
      for (int q = 1; q < HG; q++) m_g = (q == g) ? m_run[q] : m_g;
      if (pi < PPG && p < tn) ev_mine = expf(sc[p * G + h0 + g] - m_g);
    }
    int pi = 0;
    for (int pb = wg * RPWI; pb < tn; pb += NW * RPWI, pi++) {  // warp-uniform trip count (shuffles below)
      const int p = pb + grp;
      const bool valid = p < tn;
      float vf[8];
      uint4 vraw = make_uint4(0u, 0u, 0u, 0u);
      if (valid) vraw = (p == new_slot) ? ldcg_v4(qkv + (size_t)(a.n_heads + a.n_kv + kvh) * HD + d0)   // the appended v row
                                        : *reinterpret_cast<const uint4 *>(Vs + (size_t)p * HD + d0);
      unpack8<T>(vraw, vf);
#pragma unroll
      for (int g = 0; g < HG; g++) {
        float ev;
        if (SHARE_EXP) ev = __shfl_sync(0xffffffffu, ev_mine, grp * LPR + pi * HG + g);  // 0 for rows past the tile
        else ev = valid ? expf(sc[p * G + h0 + g] - m_run[g]) : 0.f;
        l_loc[g] += ev;  // identical on the LPR lanes of the row group
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] = fmaf(ev, vf[i], acc[g][i]);
      }
    }
  }
  stamp(13);

  // ---- combine the row groups: lanes of a warp first (shuffles), then the 16 warps through the (dead)
  //      K/V stage buffers of the last tile -----------------------------------------------------------
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) {
#pragma unroll
    for (int g = 0; g < HG; g++) {
      l_loc[g] += __shfl_xor_sync(0xffffffffu, l_loc[g], o);
#pragma unroll
      for (int i = 0; i < 8; i++) acc[g][i] += __shfl_xor_sync(0xffffffffu, acc[g][i], o);
    }
  }
  const bool have_tile = (s1 > s0);
  // reduction scratch: warps 0-7 in the K stage, 8-15 in the V stage, [8][HG][HD] floats each (<= 16 KB of the 32 KB stage)
  float *red0 = reinterpret_cast<float *>(Ks), *red1 = reinterpret_cast<float *>(Vs);
  if (have_tile) {
    named_bar_sync(1, MK_CT);  // every warp is done reading the last tile's K / V rows
    if (lane < LPR) {
      float *dst = (warp < 8 ? red0 : red1) + (size_t)(warp & 7) * HG * HD;
#pragma unroll
      for (int g = 0; g < HG; g++)
#pragma unroll
        for (int i = 0; i < 8; i++) dst[g * HD + lane * 8 + i] = acc[g][i];
    }
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < HG; g++) { l_warp[warp][g] = l_loc[g]; m_warp[warp][g] = m_run[g]; }
    }
    named_bar_sync(1, MK_CT);
    for (int i = ct; i < G * HD; i += MK_CT) {
      const int g = i / HD, d = i % HD, gg = g / HG, gi = g % HG;  // head g belongs to head group gg (warps gg*NW .. +NW)
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < NW; w++) {
        const int ww = gg * NW + w;
        sum += (ww < 8 ? red0 : red1)[(size_t)(ww & 7) * HG * HD + gi * HD + d];
      }
      a.ws_acc[((size_t)(kvh * G + g) * a.nsplit + split) * HD + d] = sum;
    }
    if (ct < G) {
      const int gg = ct / HG, gi = ct % HG;
      float l = 0.f;
#pragma unroll
      for (int w = 0; w < NW; w++) l += l_warp[gg * NW + w][gi];
      // the running maximum is identical in every warp of a head group (each derived it from the same scores)
      a.ws_ml[((size_t)(kvh * G + ct) * a.nsplit + split) * 2 + 0] = m_warp[gg * NW][gi];
      a.ws_ml[((size_t)(kvh * G + ct) * a.nsplit + split) * 2 + 1] = l;
    }
    __syncwarp();
    named_bar_sync(1, MK_CT);
    if (lane == 0) { mbar_arrive(&rg.empty[sk]); mbar_arrive(&rg.empty[sv]); }
  } else {
    for (int i = ct; i < G * HD; i += MK_CT) {
      const int g = i / HD, d = i % HD;
      a.ws_acc[((size_t)(kvh * G + g) * a.nsplit + split) * HD + d] = 0.f;
    }
    if (ct < G) {
      a.ws_ml[((size_t)(kvh * G + ct) * a.nsplit + split) * 2 + 0] = -INFINITY;
      a.ws_ml[((size_t)(kvh * G + ct) * a.nsplit + split) * 2 + 1] = 0.f;
    }
    named_bar_sync(1, MK_CT);
  }
  stamp(14);
  if (ct == 0) {
    // acq_rel at gpu scope: releases the partials every thread of this CTA wrote before the bar.sync (cumulative)
    // and, for the last arriver, acquires the other splits' partials
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&a.attn_counters[kvh]) : "memory");
    is_last = (ticket == (unsigned)a.nsplit - 1);
  }
  named_bar_sync(1, MK_CT);
  stamp(15);
  if (!is_last) return;
  for (int g = warp; g < G; g += MK_CW) {
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

#if MK_ATTN_V1
// ---- round-1 attention phase, kept for A/B measurements only (-DMK_ATTN_V1=1) ----
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
__device__ __forceinline__ void mk_consume_attn_v1(MkRing &rg, const MkArgs &a, const MkLayer &L, int pos, unsigned char *scr,
                                                const float *rope_cs, int ct, int warp, int lane, unsigned long long *tr) {
  auto stamp = [&](int i) {  // profiling builds: tr[12..15] = q/k/v prepared, tiles done, partials written, ticket taken
    (void)i;
    MKT(if (tr && ct == 0 && i >= 12) tr[i] = mk_gtime());
  };
  constexpr int LPR = HD / 8;            // lanes per cached row (16 B each); HD in {16,64,128} -> 2, 8, 16
  constexpr int RPWI = 32 / LPR;         // rows per warp per iteration
  constexpr int NW = MK_CW;
  static_assert(HD == 16 || HD == 64 || HD == 128, "head_dim");
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
  T *kc = reinterpret_cast<T *>(L.kc) + (size_t)it.kvh * a.cap * HD;
  T *vc = reinterpret_cast<T *>(L.vc) + (size_t)it.kvh * a.cap * HD;
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
  stamp(12);

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
  }
  stamp(13);

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
  stamp(14);
  named_bar_sync(1, MK_CT);
  if (ct == 0) {
    // acq_rel at gpu scope: releases the partials every thread of this CTA wrote before the bar.sync (cumulative)
    // and, for the last arriver, acquires the other splits' partials
    unsigned ticket;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(ticket) : "l"(&a.attn_counters[kvh]) : "memory");
    is_last = (ticket == (unsigned)a.nsplit - 1);
  }
  named_bar_sync(1, MK_CT);
  stamp(15);
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

#endif

// ---------------------------------------------------------------------------------------- the kernel
template <typename T, int HD, int G>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mega_kernel(const MkArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int es = sizeof(T);
  unsigned char *ring = smem_raw;
  T *xs = reinterpret_cast<T *>(ring + (size_t)a.n_stages * MK_STAGE_BYTES);
  T *res_s = xs + a.max_k;  // the epilogue's residual vector (hidden elements), staged together with x
  size_t off = (size_t)a.n_stages * MK_STAGE_BYTES + mk_xs_bytes(a.max_k, a.hidden, es);
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

  // GEMV phases are numbered gp = 4*layer + {0 qkv, 1 o_proj, 2 gate_up, 3 down}; gp = 4*n_layers is the lm_head.
  const int n_gp = 4 * a.n_layers + (a.has_head ? 1 : 0);
  auto gp_geom = [&](int gp) -> const MkGeom & {
    return gp >= 4 * a.n_layers ? a.g_head : ((gp & 3) == 0 ? a.g_qkv : (gp & 3) == 1 ? a.g_o : (gp & 3) == 2 ? a.g_gu : a.g_down);
  };
  auto gp_weights = [&](int gp) -> const void * {
    if (gp >= 4 * a.n_layers) return a.lm_head;
    const MkLayer *L = a.layers + (gp >> 2);
    return (gp & 3) == 0 ? L->wqkv : (gp & 3) == 1 ? L->wo : (gp & 3) == 2 ? L->wgu : L->wd;
  };
  if (warp >= MK_CW) {
#if MK_SETMAXNREG
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MK_REGS_PRODUCER));
#endif
    if (warp == MK_CW) {
      // ================= producer: stream weights and K/V tiles through the whole phase list ==========
      if (lane == 0) {
        const uint64_t pol_w = policy_evict_first(), pol_kv = policy_evict_last();
        // The layer table is in global memory.  Reading a weight pointer from it when a phase STARTS stalls the issue
        // loop for an L2 round trip (1-2 us under the weight stream) at every phase boundary — exactly when ring slots
        // are free and HBM should be fed (A/B r02: -5 %).  The pointers of layer l+1 are therefore fetched while the
        // long gate_up phase of layer l is being issued.
        MkLayer Lc = a.n_layers > 0 ? a.layers[0] : MkLayer{}, Ln = Lc;
        for (int gp = 0; gp < n_gp; gp++) {
          const MkGeom &g = gp_geom(gp);
          const int ph = gp & 3;
          const bool head = gp >= 4 * a.n_layers;
          if (!head && ph == 0 && gp > 0) Lc = Ln;
          if (!head && ph == 2 && (gp >> 2) + 1 < a.n_layers) Ln = a.layers[(gp >> 2) + 1];  // in flight during gate_up
          const void *W = head ? a.lm_head : (ph == 0 ? Lc.wqkv : ph == 1 ? Lc.wo : ph == 2 ? Lc.wgu : Lc.wd);
          unsigned long long *tr = !head ? mk_trec(a.trace, gp >> 2, ph == 0 ? 0 : ph + 1) : nullptr;
          if (mk_is_dynamic(g)) mk_produce_gemv_dyn<T>(rg, g, W, a.tickets + gp, a.max_groups, pol_w, tr);
          else mk_produce_gemv<T>(rg, g, W, (!head && ph == 2) ? 2 : 1, pol_w, tr);
          if (ph == 0 && !head) {
            MKT(if (mk_trec(a.trace, gp >> 2, 1)) mk_trec(a.trace, gp >> 2, 1)[7] = mk_gtime());
            mk_produce_attn<T>(rg, a, Lc, pos, pol_kv);
            MKT(if (mk_trec(a.trace, gp >> 2, 1)) mk_trec(a.trace, gp >> 2, 1)[9] = mk_gtime());
          }
        }
      }
    }
    return;
  }

  // ================= consumers ========================================================================
#if MK_SETMAXNREG
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MK_REGS_CONSUMER));
#endif
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
  // One loop over the GEMV phases (gp = 4*layer + {qkv, o_proj, gate_up, down}; 4*n_layers = lm_head): the staging, the
  // GEMV core and the attention phase exist ONCE in the instruction stream.
  // The layer table lives in global memory: reading a pointer from it at the start of a phase puts an L2 round trip
  // (~1 us under the weight stream) on the critical path.  Two layers' entries are kept in shared memory instead; the
  // next layer's entry is copied in during the current layer's gate_up phase.
  __shared__ MkLayer Ls[2];
  auto cache_layer = [&](int l_) {
    constexpr int NW64 = (int)(sizeof(MkLayer) / 8);
    if (l_ < a.n_layers && ct < NW64)
      reinterpret_cast<unsigned long long *>(&Ls[l_ & 1])[ct] = reinterpret_cast<const unsigned long long *>(a.layers + l_)[ct];
  };
  if (MK_LS_CACHE) {
    cache_layer(0);
    named_bar_sync(1, MK_CT);
  }
  MkEpi e{};
  // this thread's slice of the NEXT normalising phase's weight vector (ln1 / ln2 / ln_f): independent of the activations,
  // so it is fetched before the grid barrier in front of that phase instead of after the x load
  uint4 pre_w[2];
  auto preload_norm = [&](int gp_next) {
    const void *w = nullptr;
    if (gp_next >= n_gp) return;
    if (gp_next >= 4 * a.n_layers) w = a.ln_f;
    else if ((gp_next & 3) == 0) w = Ls[(gp_next >> 2) & 1].ln1;
    else if ((gp_next & 3) == 2) w = Ls[(gp_next >> 2) & 1].ln2;
    if (!w) return;
    const int nvw = a.hidden * (int)sizeof(T) / 16;
    const uint4 *wg = reinterpret_cast<const uint4 *>(w);
    pre_w[0] = ct < nvw ? wg[ct] : make_uint4(0u, 0u, 0u, 0u);
    pre_w[1] = ct + MK_CT < nvw ? wg[ct + MK_CT] : make_uint4(0u, 0u, 0u, 0u);
  };
  pre_w[0] = pre_w[1] = make_uint4(0u, 0u, 0u, 0u);
  if (MK_RES_STAGE) preload_norm(0);
  for (int gp = 0; gp < n_gp; gp++) {
    const int l = gp >> 2, ph = gp & 3;
    const bool head = gp >= 4 * a.n_layers;
    const MkLayer *Lp = MK_LS_CACHE ? &Ls[l & 1] : a.layers + (head ? 0 : l);
    if (MK_LS_CACHE && ph == 2) cache_layer(l + 1);  // visible to every consumer after this phase's barriers
    const bool last_layer = (l == a.n_layers - 1);
    T *layer_out = reinterpret_cast<T *>(last_layer ? a.x_out : a.xa);   // this layer's output / next layer's input
    const MkGeom &g = gp_geom(gp);
    // per phase: x source, fused RMS-norm weight, epilogue kind and its operands
    const void *x_src, *norm_w = nullptr, *res_src = nullptr;
    int K, epi, trp = ph == 0 ? 0 : ph + 1;
    e = MkEpi{};
    if (head) {         // ln_f + lm_head + greedy argmax (text_model.rs:336-352,102-118)
      x_src = cur; norm_w = a.ln_f; K = a.hidden; epi = EPI_ARGMAX;
      e.out = a.logits; e.part_val = a.part_val; e.part_idx = a.part_idx; e.counter = a.argmax_counter;
      e.token_out = a.token_out; e.token_ring = a.token_ring; e.step = a.d_step; e.ring_cap = a.ring_cap;
      e.advance = a.advance; e.d_pos = a.d_pos; e.d_step = a.d_step; e.gbar = a.gbar;
      e.next_start = start + (unsigned long long)nbar * gridDim.x;
      e.ring_seq = a.inbox_ctr ? a.ring_seq : nullptr;
    } else if (ph == 0) {  // rms_1 + qkv (+bias)
      x_src = cur; norm_w = Lp->ln1; K = a.hidden; epi = EPI_PLAIN; e.bias = Lp->bqkv; e.out = a.qkv;
    } else if (ph == 1) {  // o_proj + residual
      x_src = a.y; K = a.n_heads * a.hd; epi = EPI_RESIDUAL; e.out = a.xb;
      if (MK_RES_STAGE) { res_src = cur; e.residual = res_s; } else { e.residual = cur; }
    } else if (ph == 2) {  // rms_2 + gate_up + silu*mul
      x_src = a.xb; norm_w = Lp->ln2; K = a.hidden; epi = EPI_SWIGLU; e.out = a.mm;
    } else {               // down + residual
      x_src = a.mm; K = a.inter; epi = EPI_RESIDUAL; e.out = layer_out;
      if (MK_RES_STAGE) { res_src = a.xb; e.residual = res_s; } else { e.residual = a.xb; }
    }
#if MK_TRACE
    unsigned long long *tp = (!head && ct == 0) ? mk_trec(a.trace, l, trp) : nullptr;
#define TS(rec, i) do { if (rec) (rec)[i] = mk_gtime(); } while (0)
#else
    unsigned long long *const tp = nullptr;
    (void)trp;
#define TS(rec, i) do { } while (0)
#endif
    TS(tp, 0);
    mk_stage_x<T>(xs, x_src, norm_w, K, a.eps, scratch, ct, warp, lane, res_s, res_src, a.hidden, (MK_RES_STAGE && norm_w) ? pre_w : nullptr);
    TS(tp, 1);
    const MkRows rw = mk_gemv_core<T>(rg, g, epi == EPI_SWIGLU ? 2 : 1, xs, partial, loc_row, ct, warp, lane, tp);
    if (epi == EPI_PLAIN) mk_gemv_epilogue<T, EPI_PLAIN>(g, rw, partial, loc_row, scratch, e, ct, warp, lane);
    else if (epi == EPI_RESIDUAL) mk_gemv_epilogue<T, EPI_RESIDUAL>(g, rw, partial, loc_row, scratch, e, ct, warp, lane);
    else if (epi == EPI_SWIGLU) mk_gemv_epilogue<T, EPI_SWIGLU>(g, rw, partial, loc_row, scratch, e, ct, warp, lane);
    else mk_gemv_epilogue<T, EPI_ARGMAX>(g, rw, partial, loc_row, scratch, e, ct, warp, lane);
    TS(tp, 3);
    if (head) break;
    if (MK_RES_STAGE && ph != 0) preload_norm(gp + 1);   // (after qkv the attention phase comes first; o_proj has no norm)
    if (ph < 3 || !last_layer || a.has_head) gsync();
    TS(tp, 4);
    if (ph == 0) {  // qk-norm, RoPE, KV append, attention
#if MK_TRACE
      unsigned long long *ta = (ct == 0) ? mk_trec(a.trace, l, 1) : nullptr;
#else
      unsigned long long *const ta = nullptr;
#endif
      TS(ta, 0);
#if MK_ATTN_V1
      mk_consume_attn_v1<T, HD, G>(rg, a, *Lp, pos, reinterpret_cast<unsigned char *>(xs), rope_cs, ct, warp, lane, ta);
#else
      mk_consume_attn<T, HD, G>(rg, a, *Lp, pos, reinterpret_cast<unsigned char *>(xs), rope_cs, ct, warp, lane, ta);
#endif
      TS(ta, 3);
      gsync();
      TS(ta, 4);
    }
    if (ph == 3) cur = layer_out;
#undef TS
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
