// gemv_sweep.cu — experiment harness (not product code): times variants of the TMA-streamed decode
// GEMV on Llama-3-8B shapes to choose the stage geometry / consumer mapping used in csrc/gemv.cuh.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cake_b200/csrc -o bench_tools/gemv_sweep bench_tools/gemv_sweep.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
using namespace cake;
typedef __nv_bfloat16 T;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e = (x);                                                                   \
    if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } \
  } while (0)

constexpr int CW = 8;  // consumer warps
struct Args {
  const T *W; const T *x; const T *nw; T *out;
  int N, K, KC, RS, WPR, n_stages, compute, norm;
};

// stage = RS row segments of KC columns.  Consumer warps form (CW/WPR) row slots x WPR column slices;
// a warp handles RPW = RS/(CW/WPR) rows of each stage over its column slice of KC/WPR columns.
template <int RPW>
__global__ void __launch_bounds__((CW + 1) * 32, 1) k_gemv(const Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int K = a.K, KC = a.KC, RS = a.RS, WPR = a.WPR, nchunk = K / KC;
  const size_t seg = (size_t)KC * 2, stage = (size_t)RS * seg;
  unsigned char *ring = smem;
  T *xs = (T *)(ring + a.n_stages * stage);
  size_t off = a.n_stages * stage + (size_t)K * 2;
  off = (off + 15) & ~(size_t)15;
  float *partial = (float *)(smem + off);  // [rows][WPR]
  const int max_rows = a.N / gridDim.x + 2;
  off += (size_t)max_rows * WPR * 4;
  float *scratch = (float *)(smem + off);
  off += 64 * 4;
  off = (off + 7) & ~(size_t)7;
  uint64_t *full = (uint64_t *)(smem + off), *empty = full + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = (int)((long)a.N * blockIdx.x / gridDim.x), r1 = (int)((long)a.N * (blockIdx.x + 1) / gridDim.x);
  const int nrows = r1 - r0, ngroups = (nrows + RS - 1) / RS;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.n_stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
    mbar_fence_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  if (warp == CW) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      const unsigned char *Wb = (const unsigned char *)a.W;
      int s = 0; uint32_t ph = 0;
      for (int g = 0; g < ngroups; g++) {
        const int row = r0 + g * RS, nr = min(RS, r1 - row);
        for (int j = 0; j < nchunk; j++) {
          mbar_wait(&empty[s], ph ^ 1u);
          unsigned char *dst = ring + s * stage;
          mbar_arrive_expect_tx(&full[s], (uint32_t)(nr * seg));
          if (nchunk == 1) bulk_g2s(dst, Wb + (size_t)row * K * 2, (uint32_t)(nr * seg), &full[s], pol);
          else for (int r = 0; r < nr; r++) bulk_g2s(dst + r * seg, Wb + ((size_t)(row + r) * K + (size_t)j * KC) * 2, (uint32_t)seg, &full[s], pol);
          if (++s == a.n_stages) { s = 0; ph ^= 1u; }
        }
      }
    }
    return;
  }
  const int ct = threadIdx.x, CT = CW * 32;
  pdl_wait();
  {
    const uint4 *xg = (const uint4 *)a.x; uint4 *xsv = (uint4 *)xs; const int nv = K / 8;
    if (!a.norm) { for (int v = ct; v < nv; v += CT) xsv[v] = xg[v]; }
    else {
      float ss = 0.f;
      for (int v = ct; v < nv; v += CT) { uint4 u = xg[v]; xsv[v] = u; float f[8]; unpack8<T>(u, f);
#pragma unroll
        for (int i = 0; i < 8; i++) ss += f[i] * f[i]; }
      ss = warp_sum(ss);
      if (lane == 0) scratch[warp] = ss;
      named_bar_sync(1, CT);
      float tot = 0.f;
      for (int w = 0; w < CW; w++) tot += scratch[w];
      const float inv = 1.0f / sqrtf(tot / (float)K + 1e-5f);
      const uint4 *wg = (const uint4 *)a.nw;
      for (int v = ct; v < nv; v += CT) { float f[8], w8[8]; unpack8<T>(xsv[v], f); unpack8<T>(wg[v], w8); uint4 o;
        o.x = pack2<T>(f[0]*inv*w8[0], f[1]*inv*w8[1]); o.y = pack2<T>(f[2]*inv*w8[2], f[3]*inv*w8[3]);
        o.z = pack2<T>(f[4]*inv*w8[4], f[5]*inv*w8[5]); o.w = pack2<T>(f[6]*inv*w8[6], f[7]*inv*w8[7]); xsv[v] = o; }
    }
    named_bar_sync(1, CT);
  }
  {
    const int slots = CW / WPR;            // row slots
    const int slot = warp / WPR, ks = warp % WPR;
    const int nvec = KC / 8 / WPR;         // vectors per warp per row segment
    const uint4 *xsv = (const uint4 *)xs;
    int s = 0; uint32_t ph = 0;
    for (int g = 0; g < ngroups; g++) {
      float acc[RPW][2];
#pragma unroll
      for (int r = 0; r < RPW; r++) acc[r][0] = acc[r][1] = 0.f;
      for (int j = 0; j < nchunk; j++) {
        mbar_wait(&full[s], ph);
        if (a.compute) {
          const uint4 *st = (const uint4 *)(ring + s * stage);
          const uint4 *xc = xsv + (size_t)j * (KC / 8) + ks * nvec;
#pragma unroll 2
          for (int v = lane; v < nvec; v += 32) {
            float xf[8]; unpack8<T>(xc[v], xf);
#pragma unroll
            for (int r = 0; r < RPW; r++) {
              float wf[8]; unpack8<T>(st[(size_t)(slot + r * slots) * (KC / 8) + ks * nvec + v], wf);
#pragma unroll
              for (int i = 0; i < 8; i++) acc[r][i & 1] = fmaf(wf[i], xf[i], acc[r][i & 1]);
            }
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
  for (int rl = ct; rl < nrows; rl += CT) {
    float s = 0.f;
    for (int w = 0; w < WPR; w++) s += partial[rl * WPR + w];
    a.out[r0 + rl] = DT<T>::from_f(s);
  }
}

// plain-LDG comparison: warp owns rows, 16B loads, unroll U, no shared-memory staging of W
template <int U>
__global__ void __launch_bounds__(512, 2) k_ldg(const Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  T *xs = (T *)smem;
  pdl_launch_dependents();
  pdl_wait();
  const int K = a.K;
  for (int v = threadIdx.x; v < K / 8; v += blockDim.x) ((uint4 *)xs)[v] = ((const uint4 *)a.x)[v];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int r0 = (int)((long)a.N * blockIdx.x / gridDim.x), r1 = (int)((long)a.N * (blockIdx.x + 1) / gridDim.x);
  const int nvec = K / 8;
  for (int r = r0 + warp; r < r1; r += nw) {
    const uint4 *wr = (const uint4 *)(a.W + (size_t)r * K);
    float acc0 = 0.f, acc1 = 0.f;
    for (int v0 = lane; v0 < nvec; v0 += 32 * U) {
      uint4 wv[U];
#pragma unroll
      for (int u = 0; u < U; u++) { const int v = v0 + 32 * u; wv[u] = (v < nvec) ? ld_nc_v4(wr + v) : make_uint4(0, 0, 0, 0); }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int v = v0 + 32 * u;
        if (v < nvec) { float wf[8], xf[8]; unpack8<T>(wv[u], wf); unpack8<T>(((const uint4 *)xs)[v], xf);
#pragma unroll
          for (int i = 0; i < 8; i++) { if (i & 1) acc1 = fmaf(wf[i], xf[i], acc1); else acc0 = fmaf(wf[i], xf[i], acc0); } }
      }
    }
    const float s = warp_sum(acc0 + acc1);
    if (lane == 0) a.out[r] = DT<T>::from_f(s);
  }
}

template <typename Kern>
static void launch(Kern k, dim3 grid, dim3 block, size_t smem, cudaStream_t st, const Args &a, bool pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, k, a));
}

int main(int argc, char **argv) {
  int sms = 148;
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0)); sms = prop.multiProcessorCount;
  cudaStream_t st; CK(cudaStreamCreate(&st));
  struct Shape { const char *name; int N, K; } shapes[] = {{"gate_up", 28672, 4096}, {"down", 4096, 14336}, {"qkv", 6144, 4096}, {"o", 4096, 4096}};
  const size_t maxW = (size_t)28672 * 4096;
  const int NCOPY = 12;  // 12 x 235 MB: consecutive launches never hit L2
  T *W; CK(cudaMalloc(&W, maxW * 2 * NCOPY)); CK(cudaMemset(W, 0x11, maxW * 2 * NCOPY));
  T *x, *nw, *out; CK(cudaMalloc(&x, 65536)); CK(cudaMalloc(&nw, 65536)); CK(cudaMalloc(&out, 1 << 20));
  CK(cudaMemset(x, 0x11, 65536)); CK(cudaMemset(nw, 0x11, 65536));
  auto set_attr = [&](auto k) {
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  };
  set_attr(k_gemv<1>); set_attr(k_gemv<2>); set_attr(k_gemv<4>); set_attr(k_ldg<4>); set_attr(k_ldg<8>);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  struct Cfg { const char *name; int KCdiv; int RS; int WPR; int smem_kb; int compute; int pdl; };
  // KCdiv: KC = K / KCdiv (0 -> KC = 1024)
  Cfg cfgs[] = {
      {"A ks8 R2 full-row 16K (orig)", 1, 2, 8, 108, 1, 1},
      {"A' same, no PDL", 1, 2, 8, 108, 1, 0},
      {"A'' same, stream-only", 1, 2, 8, 108, 0, 1},
      {"B own-row 8x2K (current)", 0, 8, 1, 108, 1, 1},
      {"B'' same, stream-only", 0, 8, 1, 108, 0, 1},
      {"C ks4 R2: 2 slots x 4 slices", 1, 2, 4, 108, 1, 1},
      {"D ks2 R4 32K stage", 1, 4, 2, 108, 1, 1},
      {"E ks8 R4 32K stage", 1, 4, 8, 108, 1, 1},
      {"F ks8 R2 16K, 200KB ring", 1, 2, 8, 200, 1, 1},
      {"F'' same stream-only", 1, 2, 8, 200, 0, 1},
      {"G own-row 8 full rows 64K, 200KB", 1, 8, 1, 200, 1, 1},
      {"H ks8 R1 8K stage", 1, 1, 8, 108, 1, 1},
      {"I ks4 R4: 2 slots x 4 slices 32K", 1, 4, 4, 108, 1, 1},
  };
  for (auto &sh : shapes) {
    printf("== %s N=%d K=%d  (%.1f MB, ideal %.1f us @6570 GB/s)\n", sh.name, sh.N, sh.K, sh.N * (double)sh.K * 2 / 1e6, sh.N * (double)sh.K * 2 / 6570e3);
    const size_t wsz = (size_t)sh.N * sh.K;
    for (auto &c : cfgs) {
      int KC = c.KCdiv ? sh.K / c.KCdiv : 1024;
      int RS = c.RS;
      if (sh.K == 14336 && c.KCdiv == 1) { KC = sh.K / 2; RS = c.RS > 1 ? c.RS / 2 : 1; }  // 28 KB rows: halve
      if (KC / 8 / c.WPR < 1) continue;
      const int rpw = RS / (CW / c.WPR);
      if (rpw < 1 || RS % (CW / c.WPR)) { printf("  %-36s skipped (RS=%d WPR=%d)\n", c.name, RS, c.WPR); continue; }
      const size_t stage = (size_t)RS * KC * 2;
      const size_t fixed = (size_t)sh.K * 2 + (sh.N / sms + 2) * c.WPR * 4 + 64 * 4 + 32 * 8 + 256;
      int ns = (int)(((size_t)c.smem_kb * 1024 - fixed) / stage);
      if (ns > 16) ns = 16;
      if (ns < 2) { printf("  %-36s skipped (stage %zu too big)\n", c.name, stage); continue; }
      const size_t smem = ns * stage + fixed;
      Args a{nullptr, x, nw, out, sh.N, sh.K, KC, RS, c.WPR, ns, c.compute, 1};
      auto go = [&](int copy) {
        a.W = W + (size_t)copy * wsz;
        if (rpw == 1) launch(k_gemv<1>, dim3(sms), dim3((CW + 1) * 32), smem, st, a, c.pdl);
        else if (rpw == 2) launch(k_gemv<2>, dim3(sms), dim3((CW + 1) * 32), smem, st, a, c.pdl);
        else launch(k_gemv<4>, dim3(sms), dim3((CW + 1) * 32), smem, st, a, c.pdl);
      };
      const int ncopy = (int)((maxW * NCOPY) / wsz) > 64 ? 64 : (int)((maxW * NCOPY) / wsz);
      for (int i = 0; i < ncopy; i++) go(i);
      CK(cudaStreamSynchronize(st));
      const int reps = 3;
      CK(cudaEventRecord(e0, st));
      for (int r = 0; r < reps; r++) for (int i = 0; i < ncopy; i++) go(i);
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double us = ms * 1e3 / (reps * ncopy);
      printf("  %-36s KC=%5d RS=%d WPR=%d stages=%2d smem=%3zuKB : %7.2f us  %7.1f GB/s\n", c.name, KC, RS, c.WPR, ns, smem / 1024, us, wsz * 2 / us / 1e3);
    }
    for (int U : {4, 8}) {
      Args a{nullptr, x, nw, out, sh.N, sh.K, 0, 0, 0, 0, 1, 0};
      const int ncopy = (int)((maxW * NCOPY) / wsz) > 64 ? 64 : (int)((maxW * NCOPY) / wsz);
      auto go = [&](int copy) { a.W = W + (size_t)copy * wsz;
        if (U == 4) launch(k_ldg<4>, dim3(sms * 2), dim3(512), (size_t)sh.K * 2, st, a, true);
        else launch(k_ldg<8>, dim3(sms * 2), dim3(512), (size_t)sh.K * 2, st, a, true); };
      for (int i = 0; i < ncopy; i++) go(i);
      CK(cudaStreamSynchronize(st));
      CK(cudaEventRecord(e0, st));
      for (int r = 0; r < 3; r++) for (int i = 0; i < ncopy; i++) go(i);
      CK(cudaEventRecord(e1, st));
      CK(cudaStreamSynchronize(st));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      const double us = ms * 1e3 / (3 * ncopy);
      printf("  LDG.128 warp-owns-row unroll %d, 2 CTA/SM x 512 thr          : %7.2f us  %7.1f GB/s\n", U, us, wsz * 2 / us / 1e3);
    }
  }
  return 0;
}
