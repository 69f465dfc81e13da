// gemv.cuh — the decode-time linear layer: out[N] = epilogue( W[N,K] · prologue(x[K]) ), batch 1.
//
// This kernel moves >97% of the bytes of a decoded token, so it is built around
// HBM streaming, not math:
//   * persistent grid, one CTA per SM (148); each CTA owns a contiguous row range [r0,r1) so the
//     weight bytes of a CTA are ONE contiguous stream (row-major W) — perfectly sequential DRAM pages;
//   * a dedicated producer warp issues TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP) of 14-32 KB
//     row blocks into a shared-memory ring guarded by full/empty mbarriers; ~80-100 KB in flight
//     per SM with zero register cost, L2 policy evict_first (each weight byte is used once per token);
//   * the producer starts streaming BEFORE griddepcontrol.wait: weights do not depend on the previous
//     kernel, so with programmatic dependent launch the ring is already full when the activation
//     arrives — the kernel boundary costs no HBM idle time;
//   * 8 consumer warps as (row slot x column slice); x (optionally RMS-normalised in-kernel, f32 sum
//     of squares) lives in shared memory in dtype D; fp32 accumulation in 2 independent chains per
//     row per lane, one shuffle tree per row per warp, slices summed in a fixed order — results are
//     bit-deterministic run to run;
//   * epilogues fuse what the reference does as separate tensor ops: +bias, +residual (transformer.rs:123,131),
//     silu(gate)*up on row-interleaved Wgu (mlp.rs:22-28), logits + greedy argmax (text_model.rs:348-352,104-105).
// Rounding points follow SURVEY.md Appendix A: the matmul result is rounded to D before any epilogue op.
#pragma once
#include "common.cuh"

namespace cake {

enum { EPI_PLAIN = 0, EPI_RESIDUAL = 1, EPI_SWIGLU = 2, EPI_ARGMAX = 3 };

constexpr int GEMV_CONSUMER_WARPS = 8;
constexpr int GEMV_THREADS = (GEMV_CONSUMER_WARPS + 1) * 32;  // + 1 producer warp
constexpr int GEMV_MAX_STAGES = 16;

struct GemvArgs {
  const void *W;         // [N,K] row-major, D
  const void *x;         // [K] D
  const void *norm_w;    // [K] D or nullptr: fused RMSNorm prologue
  const void *bias;      // [N] D or nullptr
  const void *residual;  // [N] D (EPI_RESIDUAL)
  void *out;             // [N] D  ([N/2] for EPI_SWIGLU)
  float eps;
  int N, K;
  int act;  // EPI_SWIGLU: 0 silu, 1 gelu_tanh
  // stage geometry (host-planned, see plan_gemv): a stage holds RS row segments of KC columns; the 8
  // consumer warps form (8/WPR) row slots x WPR column slices; a warp covers RS/(8/WPR) rows per stage.
  int KC, RS, WPR;
  int n_stages;
  int max_rows;  // upper bound of rows per CTA (sizes the partial-sum scratch)
  // EPI_ARGMAX
  float *part_val;
  int *part_idx;
  unsigned *counter;
  uint32_t *token_out;    // greedy token
  uint32_t *token_ring;   // optional history ring (decode loop), indexed by *step
  const int *step;
  int ring_cap;
};

__host__ __device__ inline size_t gemv_smem_bytes(int K, int KC, int RS, int WPR, int n_stages, int max_rows, int es) {
  size_t stage = (size_t)RS * KC * es;
  size_t off = (size_t)n_stages * stage;                 // ring
  off += (size_t)K * es;                                 // xs
  off = (off + 15) & ~(size_t)15;
  off += (size_t)max_rows * WPR * 4;                     // partial sums [row][slice]
  off += 64 * 4;                                         // reduction scratch
  off = (off + 7) & ~(size_t)7;
  off += (size_t)2 * GEMV_MAX_STAGES * 8;                // mbarriers
  return off + 128;                                      // alignment slack
}

// Geometry chosen from measurements (bench_tools/gemv_sweep.cu, profiles/gemv_sweep_r01.txt): TMA bulk
// copies must be large (>= 16 KB; 2 KB copies cap at 4.3 TB/s even with no math) and stages ~32 KB;
// 2 row slots x 4 column slices keeps every lane busy with one shuffle tree per row per warp.
template <typename T, int EPI, int RPW>
__global__ void __launch_bounds__(GEMV_THREADS, 1) gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int es = sizeof(T);
  const int K = a.K, N = a.N, KC = a.KC, RS = a.RS, WPR = a.WPR;
  const int nchunk = K / KC;
  const size_t seg_bytes = (size_t)KC * es;
  const size_t stage_bytes = (size_t)RS * seg_bytes;
  unsigned char *ring = smem_raw;
  T *xs = reinterpret_cast<T *>(ring + (size_t)a.n_stages * stage_bytes);
  size_t off = (size_t)a.n_stages * stage_bytes + (size_t)K * es;
  off = (off + 15) & ~(size_t)15;
  float *partial = reinterpret_cast<float *>(smem_raw + off);
  off += (size_t)a.max_rows * WPR * 4;
  float *scratch = reinterpret_cast<float *>(smem_raw + off);
  off += 64 * 4;
  off = (off + 7) & ~(size_t)7;
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + off);
  uint64_t *empty = full + GEMV_MAX_STAGES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int G = (EPI == EPI_SWIGLU) ? 2 : 1;  // row granularity of the partition
  const long units = N / G;
  const int r0 = (int)(units * blockIdx.x / gridDim.x) * G;
  const int r1 = (int)(units * (blockIdx.x + 1) / gridDim.x) * G;
  const int nrows = r1 - r0;
  const int ngroups = (nrows + RS - 1) / RS;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.n_stages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], GEMV_CONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();

  if (warp == GEMV_CONSUMER_WARPS) {
    // ===================== producer warp: stream this CTA's weight rows, independent of x ==========
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      const unsigned char *Wb = reinterpret_cast<const unsigned char *>(a.W);
      int s = 0;
      uint32_t ph = 0;
      for (int g = 0; g < ngroups; g++) {
        const int row = r0 + g * RS;
        const int nr = min(RS, r1 - row);
        for (int j = 0; j < nchunk; j++) {
          mbar_wait(&empty[s], ph ^ 1u);
          unsigned char *dst = ring + (size_t)s * stage_bytes;
          mbar_arrive_expect_tx(&full[s], (uint32_t)(nr * seg_bytes));
          if (nchunk == 1) {  // whole rows: the nr rows are one contiguous block -> ONE bulk copy
            bulk_g2s(dst, Wb + (size_t)row * K * es, (uint32_t)(nr * seg_bytes), &full[s], pol);
          } else {
            for (int r = 0; r < nr; r++)
              bulk_g2s(dst + r * seg_bytes, Wb + ((size_t)(row + r) * K + (size_t)j * KC) * es, (uint32_t)seg_bytes,
                       &full[s], pol);
          }
          if (++s == a.n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    return;
  }

  // ======================= consumer warps ========================================================
  constexpr int CT = GEMV_CONSUMER_WARPS * 32;
  const int ct = threadIdx.x;  // 0..255
  pdl_wait();                  // x (and residual) come from the previous kernel

  // ---- prologue: x -> shared (D), optionally RMS-normalised (backends/mod.rs:244-246) -------------
  {
    const uint4 *xg = reinterpret_cast<const uint4 *>(a.x);
    uint4 *xsv = reinterpret_cast<uint4 *>(xs);
    const int nv = K * es / 16;
    if (a.norm_w == nullptr) {
      for (int v = ct; v < nv; v += CT) xsv[v] = xg[v];
    } else {
      float ss = 0.f;
      for (int v = ct; v < nv; v += CT) {
        uint4 u = xg[v];
        xsv[v] = u;
        float f[8];
        unpack8<T>(u, f);
#pragma unroll
        for (int i = 0; i < 8; i++) ss += f[i] * f[i];
      }
      ss = warp_sum(ss);
      if (lane == 0) scratch[warp] = ss;
      named_bar_sync(1, CT);
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < GEMV_CONSUMER_WARPS; w++) tot += scratch[w];
      const float inv = 1.0f / sqrtf(tot / (float)K + a.eps);
      const uint4 *wg = reinterpret_cast<const uint4 *>(a.norm_w);
      for (int v = ct; v < nv; v += CT) {  // each thread rewrites exactly the vectors it staged
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
    named_bar_sync(1, CT);
  }

  // ---- main loop: warp (slot, ks) covers rows slot, slot+slots, .. of each stage over column slice ks
  {
    const int slots = GEMV_CONSUMER_WARPS / WPR;
    const int slot = warp / WPR, ks = warp % WPR;
    const int segv = KC / 8;           // 16-byte vectors per row segment
    const int nvec = segv / WPR;       // ... per warp
    const uint4 *xsv = reinterpret_cast<const uint4 *>(xs);
    int s = 0;
    uint32_t ph = 0;
    for (int g = 0; g < ngroups; g++) {
      float acc[RPW][2];
#pragma unroll
      for (int r = 0; r < RPW; r++) acc[r][0] = acc[r][1] = 0.f;
      for (int j = 0; j < nchunk; j++) {
        mbar_wait(&full[s], ph);
        const uint4 *st = reinterpret_cast<const uint4 *>(ring + (size_t)s * stage_bytes);
        const uint4 *xc = xsv + (size_t)j * segv + ks * nvec;
#pragma unroll 2
        for (int v = lane; v < nvec; v += 32) {
          float xf[8];
          unpack8<T>(xc[v], xf);
#pragma unroll
          for (int r = 0; r < RPW; r++) {
            float wf[8];
            unpack8<T>(st[(size_t)(slot + r * slots) * segv + ks * nvec + v], wf);
#pragma unroll
            for (int i = 0; i < 8; i++) acc[r][i & 1] = fmaf(wf[i], xf[i], acc[r][i & 1]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == a.n_stages) { s = 0; ph ^= 1u; }
      }
#pragma unroll
      for (int r = 0; r < RPW; r++) {
        const float v = warp_sum(acc[r][0] + acc[r][1]);
        const int rl = g * RS + slot + r * slots;
        if (lane == 0 && rl < nrows) partial[rl * WPR + ks] = v;
      }
    }
  }
  named_bar_sync(1, CT);

  // ---- epilogue -----------------------------------------------------------------------------------
  auto row_sum = [&](int rl) {
    float s = 0.f;
    for (int w = 0; w < WPR; w++) s += partial[rl * WPR + w];  // fixed order: deterministic
    return rnd<T>(s);                                          // matmul result ->D
  };
  T *out = reinterpret_cast<T *>(a.out);
  if (EPI == EPI_PLAIN) {
    const T *bias = reinterpret_cast<const T *>(a.bias);
    for (int rl = ct; rl < nrows; rl += CT) {
      float v = row_sum(rl);
      if (bias) v = rnd<T>(v + DT<T>::to_f(bias[r0 + rl]));  // broadcast_add in D (backends/mod.rs:237-240)
      out[r0 + rl] = DT<T>::from_f(v);
    }
  } else if (EPI == EPI_RESIDUAL) {
    const T *res = reinterpret_cast<const T *>(a.residual);
    for (int rl = ct; rl < nrows; rl += CT) {
      float v = row_sum(rl);
      out[r0 + rl] = DT<T>::from_f(v + DT<T>::to_f(res[r0 + rl]));  // residual add in D
    }
  } else if (EPI == EPI_SWIGLU) {
    for (int p = ct; p < nrows / 2; p += CT) {
      float gte = row_sum(2 * p), up = row_sum(2 * p + 1);
      float sl = gate_act<T>(gte, a.act);  // activation ->D (cpu/mod.rs:87-89 / mlp.rs:25-26)
      out[r0 / 2 + p] = DT<T>::from_f(sl * up);       // * up ->D
    }
  } else {  // EPI_ARGMAX: logits in D + greedy token, first maximum wins (text_model.rs:104-105)
    float best = -INFINITY;
    int bidx = 0x7fffffff;
    for (int rl = ct; rl < nrows; rl += CT) {
      float v = row_sum(rl);
      if (out) out[r0 + rl] = DT<T>::from_f(v);
      if (v > best) { best = v; bidx = r0 + rl; }  // rows ascend per thread: strict > keeps the first
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > best || (ov == best && oi < bidx)) { best = ov; bidx = oi; }
    }
    int *scratch_i = reinterpret_cast<int *>(scratch + 16);
    if (lane == 0) { scratch[warp] = best; scratch_i[warp] = bidx; }
    named_bar_sync(1, CT);
    if (ct == 0) {
      for (int w = 1; w < GEMV_CONSUMER_WARPS; w++)
        if (scratch[w] > best || (scratch[w] == best && scratch_i[w] < bidx)) { best = scratch[w]; bidx = scratch_i[w]; }
      a.part_val[blockIdx.x] = best;
      a.part_idx[blockIdx.x] = bidx;
      __threadfence();
      unsigned ticket = atomicAdd(a.counter, 1u);
      if (ticket == gridDim.x - 1) {  // last CTA: fold the per-CTA winners in CTA (= row) order
        __threadfence();
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (unsigned c = 0; c < gridDim.x; c++) {
          float v = ((volatile float *)a.part_val)[c];
          int i = ((volatile int *)a.part_idx)[c];
          if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
        *a.token_out = (uint32_t)bi;
        if (a.token_ring) a.token_ring[*a.step % a.ring_cap] = (uint32_t)bi;
        *a.counter = 0;  // re-arm for the next launch / graph replay
      }
    }
  }
}

}  // namespace cake
