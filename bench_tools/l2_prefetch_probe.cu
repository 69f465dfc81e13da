// l2_prefetch_probe.cu — experiment harness (not product code): can a helper warp park the NEXT phase's weights in L2
// (cp.async.bulk.prefetch.L2) while the CTA is busy with something that leaves HBM idle, so that the next phase
// starts from L2 instead of HBM?  (candidate 2.)  The two variants tried inside the megakernel were
// issued by the producer lane itself and showed no gain; this isolates the mechanism.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cake_b200/csrc -o bench_tools/l2_prefetch_probe bench_tools/l2_prefetch_probe.cu
//   ./bench_tools/l2_prefetch_probe [region_MB=32] [idle_us=8]
// Second use: HBM wake-up.  Run mode 0 with idle_us = 0 / 100 / 1000 / 3000: if the time to the first stage (or the whole
// phase B) grows with the idle time, an idle HBM drops into a state with a visible exit latency — the suspected part of
// the 45-70 us per shard boundary in the multi-GPU ring, where a shard's HBM is idle for milliseconds between its turns.
// One CTA per SM.  Each round uses a fresh 1/Nth of a large buffer (no reuse across rounds), then:
//   phase A ("attention"): every CTA spins for idle_us — HBM idle; in mode 1/2 a helper lane prefetches this CTA's slice
//                          of region B into L2 during the spin (mode 1: 32 KB requests, mode 2: 4 KB requests)
//   phase B ("o_proj"):    the CTA streams its slice of region B through a TMA ring (evict_first) and we time it.
// Prints the B-phase duration and effective GB/s per mode; mode 0 = no prefetch (HBM-cold baseline).
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
using namespace cake;

#define CK(x)                                                                                                      \
  do {                                                                                                             \
    cudaError_t e = (x);                                                                                           \
    if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } \
  } while (0)

constexpr int STAGE = 32768, NSTAGE = 5, CW = 8;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void l2_prefetch(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

struct Args {
  const unsigned char *buf;
  size_t slice;        // bytes of region B per CTA (multiple of STAGE)
  int mode, idle_ns;
  unsigned long long *t_b;  // per CTA: duration of phase B in ns
  unsigned long long *t_first;  // per CTA: phase-B start -> first stage landed, ns
  float *sink;
};

__global__ void __launch_bounds__((CW + 2) * 32, 1) k_probe(const Args a, const size_t round_off) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)NSTAGE * STAGE), *empty = full + NSTAGE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned char *B = a.buf + round_off + (size_t)blockIdx.x * a.slice;
  const int nst = (int)(a.slice / STAGE);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], CW); }
    mbar_fence_init();
  }
  __syncthreads();
  // ---- phase A: HBM-idle work; helper warp (CW + 1) may prefetch region B into L2 meanwhile
  const unsigned long long t0 = gtime();
  if (warp == CW + 1) {
    if (lane == 0 && a.mode) {
      const uint32_t req = a.mode == 1 ? 32768u : 4096u;
      for (size_t off = 0; off < a.slice; off += req) l2_prefetch(B + off, req);
    }
  }
  while (gtime() - t0 < (unsigned long long)a.idle_ns) {
  }
  __syncthreads();
  // ---- phase B: stream region B through the ring
  const unsigned long long tb = gtime();
  float acc = 0.f;
  if (warp == CW) {
    if (lane == 0) {
      const uint64_t pol = policy_evict_first();
      int s = 0; uint32_t ph = 0;
      for (int i = 0; i < nst; i++) {
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&full[s], STAGE);
        bulk_g2s(smem + (size_t)s * STAGE, B + (size_t)i * STAGE, STAGE, &full[s], pol);
        if (++s == NSTAGE) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp < CW) {
    int s = 0; uint32_t ph = 0;
    for (int i = 0; i < nst; i++) {
      mbar_wait(&full[s], ph);
      if (i == 0 && threadIdx.x == 0) a.t_first[blockIdx.x] = gtime() - tb;
      const uint4 *st = reinterpret_cast<const uint4 *>(smem + (size_t)s * STAGE);
      for (int v = warp * 32 + lane; v < STAGE / 16; v += CW * 32) {
        const uint4 q = st[v];
        acc += __uint_as_float(q.x & 0xffff0000u) + __uint_as_float(q.w << 16);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == NSTAGE) { s = 0; ph ^= 1u; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) a.t_b[blockIdx.x] = gtime() - tb;
  if (acc == 123.456f && a.sink) a.sink[0] = acc;
}

int main(int argc, char **argv) {
  const int region_mb = argc > 1 ? atoi(argv[1]) : 32;
  const int idle_us = argc > 2 ? atoi(argv[2]) : 8;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int grid = p.multiProcessorCount, rounds = 24;
  const size_t slice = ((size_t)region_mb * 1024 * 1024 / grid) / STAGE * STAGE;
  const size_t region = slice * grid, total = region * rounds;
  unsigned char *buf;
  CK(cudaMalloc(&buf, total));
  CK(cudaMemset(buf, 1, total));
  unsigned long long *t_b, *t_first;
  CK(cudaMalloc(&t_b, grid * 8));
  CK(cudaMalloc(&t_first, grid * 8));
  const size_t smem = (size_t)NSTAGE * STAGE + 256;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  printf("%s: region B = %.1f MB (%zu KB per CTA), idle %d us, %d rounds over a %.1f GB buffer\n", p.name, region / 1e6,
         slice / 1024, idle_us, rounds, total / 1e9);
  std::vector<unsigned long long> h(grid), hf(grid);
  unsigned char *flush;  // > L2: written before every mode so that no slice of `buf` starts out L2-resident
  CK(cudaMalloc(&flush, (size_t)256 << 20));
  for (int mode = 0; mode < 3; mode++) {
    CK(cudaMemset(flush, mode, (size_t)256 << 20));
    double sum_max = 0, sum_avg = 0, sum_first = 0;
    for (int r = 0; r < rounds; r++) {
      Args a{buf, slice, mode, idle_us * 1000, t_b, t_first, nullptr};
      k_probe<<<grid, (CW + 2) * 32, smem>>>(a, (size_t)r * region);
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(h.data(), t_b, grid * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hf.data(), t_first, grid * 8, cudaMemcpyDeviceToHost));
      unsigned long long mx = 0, sm = 0, sf = 0;
      for (auto v : h) { mx = v > mx ? v : mx; sm += v; }
      for (auto v : hf) sf += v;
      if (r >= 4) { sum_max += (double)mx; sum_avg += (double)sm / grid; sum_first += (double)sf / grid; }
    }
    const double n = rounds - 4;
    printf("mode %d (%s): phase B  first stage after %.2f us  max-CTA %.2f us  avg-CTA %.2f us  -> %.0f GB/s on the slowest CTA's clock\n", mode,
           mode == 0 ? "no prefetch" : mode == 1 ? "L2 prefetch, 32 KB requests" : "L2 prefetch, 4 KB requests",
           sum_first / n / 1e3, sum_max / n / 1e3, sum_avg / n / 1e3, region / (sum_max / n));
  }
  return 0;
}
