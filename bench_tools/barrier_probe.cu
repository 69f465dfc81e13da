// barrier_probe.cu — experiment harness (not product code): what does one phase boundary of the decode megakernel
// cost, and which grid-barrier formulation is cheapest on 148 SMs?  Answers the "short phases pay ~3-4 us" question
// of DESIGN.md §4.8 in isolation.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cake_b200/csrc -o bench_tools/barrier_probe bench_tools/barrier_probe.cu
//   ./bench_tools/barrier_probe [iters=2000]
// Variants (cooperative launch, one CTA of 512 threads per SM, back-to-back boundaries):
//   0  red.release.gpu + ld.acquire.gpu poll by thread 0 (the product's mk_grid_sync)
//   1  atom.add.acq_rel ticket; the last arriver publishes a generation word, everyone polls that word
//   2  as 0, but every lane of warp 0 polls (does a wider poll see the update sooner?)
//   3  as 0 + the consumer side of a phase boundary: each CTA writes its 1/148 slice of an 8 KB vector before the
//      barrier and stages the whole vector (ld.global.cg, f32 sum of squares, block reduce) after it
//   4  as 3 with a 28 KB vector (the `down` phase: 14336 bf16)
#include <cooperative_groups.h>
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

constexpr int NT = 512;

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void sync_red_poll(unsigned long long *ctr, unsigned long long target, bool wide) {
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ULL) : "memory");
  if (wide ? threadIdx.x < 32 : threadIdx.x == 0) {
    while (ld_acq(ctr) < target) {
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void sync_ticket_gen(unsigned long long *ctr, unsigned long long *gen, unsigned long long round) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("atom.acq_rel.gpu.global.add.u64 %0, [%1], %2;" : "=l"(t) : "l"(ctr), "l"(1ULL) : "memory");
    if (t + 1 == round * gridDim.x) {
      asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(gen), "l"(round) : "memory");
    } else {
      while (ld_acq(gen) < round) {
      }
    }
  }
  __syncthreads();
}

struct Args {
  unsigned long long *ctr, *gen;
  __nv_bfloat16 *vec[2];  // ping-pong activation vectors
  float *sink;
  int variant, iters, K;
};

__global__ void __launch_bounds__(NT, 1) k_probe(const Args a) {
  __shared__ float red[NT / 32];
  extern __shared__ __align__(16) unsigned char smem[];
  float acc = 0.f;
  const int K = a.K;
  for (int it = 1; it <= a.iters; it++) {
    if (a.variant >= 3) {  // produce this CTA's slice of the vector (what a GEMV epilogue does)
      __nv_bfloat16 *dst = a.vec[it & 1];
      const int r0 = (int)((long)K * blockIdx.x / gridDim.x), r1 = (int)((long)K * (blockIdx.x + 1) / gridDim.x);
      for (int i = r0 + threadIdx.x; i < r1; i += NT) dst[i] = __float2bfloat16((float)((i + it) & 255) * 0.01f);
    }
    if (a.variant == 1) sync_ticket_gen(a.ctr, a.gen, (unsigned long long)it);
    else sync_red_poll(a.ctr, (unsigned long long)it * gridDim.x, a.variant == 2);
    if (a.variant >= 3) {  // stage the whole vector: ld.global.cg -> shared, f32 sum of squares, block reduce
      const uint4 *src = reinterpret_cast<const uint4 *>(a.vec[it & 1]);
      uint4 *xs = reinterpret_cast<uint4 *>(smem);
      float ss = 0.f;
      for (int v = threadIdx.x; v < K / 8; v += NT) {
        uint4 q = __ldcg(src + v);
        xs[v] = q;
        float f[8];
        unpack8<__nv_bfloat16>(q, f);
#pragma unroll
        for (int j = 0; j < 8; j++) ss += f[j] * f[j];
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
      __syncthreads();
      float tot = 0.f;
      for (int w = 0; w < NT / 32; w++) tot += red[w];
      acc += tot;
      __syncthreads();
    }
  }
  if (threadIdx.x == 0 && a.sink) a.sink[blockIdx.x] = acc;
}

int main(int argc, char **argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 2000;
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  const int grid = p.multiProcessorCount;
  unsigned long long *ctr;
  CK(cudaMalloc(&ctr, 256));
  __nv_bfloat16 *v0, *v1;
  CK(cudaMalloc(&v0, 65536));
  CK(cudaMalloc(&v1, 65536));
  float *sink;
  CK(cudaMalloc(&sink, grid * 4));
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  printf("%s, %d SMs, %d boundaries per launch\n", p.name, grid, iters);
  const char *names[] = {"red.release + 1-thread acquire poll (product)", "ticket + generation word", "red.release + 32-lane poll",
                         "product barrier + write/stage 8 KB vector", "product barrier + write/stage 28 KB vector"};
  for (int variant = 0; variant < 5; variant++) {
    Args a{ctr, ctr + 16, {v0, v1}, sink, variant, iters, variant == 4 ? 14336 : 4096};
    void *params[] = {(void *)&a};
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
      CK(cudaMemset(ctr, 0, 256));
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      CK(cudaLaunchCooperativeKernel((void *)k_probe, dim3(grid), dim3(NT), params, 65536, 0));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep && ms < best) best = ms;
      CK(cudaEventDestroy(e0));
      CK(cudaEventDestroy(e1));
    }
    printf("variant %d  %-48s %8.3f us per boundary\n", variant, names[variant], best * 1e3f / iters);
  }
  return 0;
}
