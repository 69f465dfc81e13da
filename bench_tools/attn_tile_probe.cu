// attn_tile_probe.cu — experiment harness (not product code): the per-tile arithmetic of the megakernel's attention
// phase (one (kv head, split) item: 128 keys x 128 dims, G = 4 query heads), in two formulations, on tiles that are
// already in shared memory.  DESIGN.md §4.8 candidate (3): is a warp-private online softmax (no CTA barrier inside the
// tile, every warp active in the softmax) faster than the current three-barrier scores -> softmax -> PV pipeline?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I cake_b200/csrc -o bench_tools/attn_tile_probe bench_tools/attn_tile_probe.cu
//   ./bench_tools/attn_tile_probe [reps=200]
// Variant A mirrors mk_consume_attn (decode_mega.cuh): scores to shared memory, barrier, one warp per head does the
//   softmax, barrier, PV, barrier, 16-warp reduction through shared memory.
// Variant B: each half-warp owns rows r = 2*warp + grp + 32*i; it keeps its own running max / sum / accumulator in
//   registers over its 4 rows, and the 32 half-warp states are merged once at the end (one log-sum-exp merge).
// Both write (m, l, acc[G][HD]) partials like the product does; the host checks o = acc / l of A against B and against
// a double-precision reference, then prints the time per tile.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"
using namespace cake;
typedef __nv_bfloat16 T;

#define CK(x)                                                                                                      \
  do {                                                                                                             \
    cudaError_t e = (x);                                                                                           \
    if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1); } \
  } while (0)

constexpr int HD = 128, G = 4, TILE = 128, NW = 16, NT = NW * 32;
constexpr int LPR = HD / 8, RPWI = 32 / LPR;  // 16 lanes per row, 2 rows per warp iteration

// halving butterfly over the 16 lanes of a row group (as decode_mega.cuh reduce16_heads, G = 4)
__device__ __forceinline__ float reduce16_heads4(float (&s)[G], int lane, int &g_out) {
  int g = 0, bit = 8;
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

struct Args {
  const T *K, *V;   // [TILE][HD]
  const float *q;   // [G][HD]
  float *acc, *ml;  // [G][HD], [G][2]
  unsigned long long *ns;
  int reps, variant;
  float scale;
};

__global__ void __launch_bounds__(NT, 1) k_probe(const Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  T *Ks = reinterpret_cast<T *>(smem), *Vs = Ks + TILE * HD;
  float *sc = reinterpret_cast<float *>(Vs + TILE * HD);  // [TILE][G]
  float *red = sc + TILE * G;                              // [NW][G][HD] floats = 32 KB
  __shared__ float m_run[G], l_run[G], fac[G], m_all[32][G], l_all[32][G];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, grp = lane / LPR, gl = lane % LPR;
  for (int i = threadIdx.x; i < TILE * HD / 8; i += NT) {
    reinterpret_cast<uint4 *>(Ks)[i] = reinterpret_cast<const uint4 *>(a.K)[i];
    reinterpret_cast<uint4 *>(Vs)[i] = reinterpret_cast<const uint4 *>(a.V)[i];
  }
  float qreg[G][8];
#pragma unroll
  for (int g = 0; g < G; g++)
#pragma unroll
    for (int i = 0; i < 8; i++) qreg[g][i] = a.q[g * HD + gl * 8 + i];
  __syncthreads();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  float out_acc[G][8];
  float out_m[G], out_l[G];
  for (int rep = 0; rep < a.reps; rep++) {
    float acc[G][8];
#pragma unroll
    for (int g = 0; g < G; g++)
#pragma unroll
      for (int i = 0; i < 8; i++) acc[g][i] = 0.f;
    if (a.variant == 0) {
      // ---------------------------------------------------------------- A: the product's formulation
      if (threadIdx.x < G) { m_run[threadIdx.x] = -INFINITY; l_run[threadIdx.x] = 0.f; }
      __syncthreads();
      for (int pb = warp * RPWI; pb < TILE; pb += NW * RPWI) {
        const int p = pb + grp;
        float kf[8];
        unpack8<T>(*reinterpret_cast<const uint4 *>(Ks + (size_t)p * HD + gl * 8), kf);
        float sg[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; i++) s = fmaf(qreg[g][i], kf[i], s);
          sg[g] = s;
        }
        int gh;
        const float v = reduce16_heads4(sg, lane, gh);
        if ((gl & 3) == 0) sc[p * G + gh] = v * a.scale;
      }
      __syncthreads();
      for (int g = warp; g < G; g += NW) {
        float mx = -INFINITY;
        for (int p = lane; p < TILE; p += 32) mx = fmaxf(mx, sc[p * G + g]);
        mx = warp_max(mx);
        const float m_new = fmaxf(m_run[g], mx);
        float sum = 0.f;
        for (int p = lane; p < TILE; p += 32) {
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
      __syncthreads();
      for (int p = warp * RPWI + grp; p < TILE; p += NW * RPWI) {
        float vf[8];
        unpack8<T>(*reinterpret_cast<const uint4 *>(Vs + (size_t)p * HD + gl * 8), vf);
#pragma unroll
        for (int g = 0; g < G; g++) {
          const float ev = sc[p * G + g];
#pragma unroll
          for (int i = 0; i < 8; i++) acc[g][i] = fmaf(ev, vf[i], acc[g][i]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int g = 0; g < G; g++) { out_m[g] = m_run[g]; out_l[g] = l_run[g]; }
    } else {
      // ---------------------------------------------------------------- B: half-warp-private online softmax
      float sr[TILE / (NW * RPWI)][G];  // this half-warp's 4 rows x G scores, replicated on its 16 lanes
#pragma unroll
      for (int it = 0; it < TILE / (NW * RPWI); it++) {
        const int p = (it * NW + warp) * RPWI + grp;
        float kf[8];
        unpack8<T>(*reinterpret_cast<const uint4 *>(Ks + (size_t)p * HD + gl * 8), kf);
        float sg[G];
#pragma unroll
        for (int g = 0; g < G; g++) {
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; i++) s = fmaf(qreg[g][i], kf[i], s);
          sg[g] = s;
        }
        int gh;
        const float v = reduce16_heads4(sg, lane, gh) * a.scale;  // lanes 4*gh .. 4*gh+3 of the group hold head gh
#pragma unroll
        for (int g = 0; g < G; g++) sr[it][g] = __shfl_sync(0xffffffffu, v, grp * LPR + g * 4);
      }
      float m[G], l[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        float mx = sr[0][g];
#pragma unroll
        for (int it = 1; it < TILE / (NW * RPWI); it++) mx = fmaxf(mx, sr[it][g]);
        m[g] = mx;
        l[g] = 0.f;
      }
#pragma unroll
      for (int it = 0; it < TILE / (NW * RPWI); it++) {
        const int p = (it * NW + warp) * RPWI + grp;
        float vf[8];
        unpack8<T>(*reinterpret_cast<const uint4 *>(Vs + (size_t)p * HD + gl * 8), vf);
#pragma unroll
        for (int g = 0; g < G; g++) {
          const float ev = expf(sr[it][g] - m[g]);
          l[g] += ev;
#pragma unroll
          for (int i = 0; i < 8; i++) acc[g][i] = fmaf(ev, vf[i], acc[g][i]);
        }
      }
      // merge the 32 half-warp states: global max per head, rescale, sum
      if (gl == 0) {
#pragma unroll
        for (int g = 0; g < G; g++) { m_all[warp * 2 + grp][g] = m[g]; l_all[warp * 2 + grp][g] = l[g]; }
      }
      __syncthreads();
      float f[G];
#pragma unroll
      for (int g = 0; g < G; g++) {
        float M = m_all[0][g];
#pragma unroll 8
        for (int s2 = 1; s2 < 32; s2++) M = fmaxf(M, m_all[s2][g]);
        f[g] = expf(m[g] - M);
        out_m[g] = M;
      }
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] *= f[g];
      if (threadIdx.x < G) {  // l of the tile = sum of the rescaled half-warp sums
        float L = 0.f, M = out_m[threadIdx.x];
        for (int s2 = 0; s2 < 32; s2++) L += l_all[s2][threadIdx.x] * expf(m_all[s2][threadIdx.x] - M);
        l_run[threadIdx.x] = L;
      }
    }
    // ---- both: combine the row groups (lanes, then the 16 warps through shared memory)
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int i = 0; i < 8; i++) acc[g][i] += __shfl_xor_sync(0xffffffffu, acc[g][i], o);
    if (lane < LPR) {
      float *dst = red + (size_t)warp * G * HD;
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int i = 0; i < 8; i++) dst[g * HD + lane * 8 + i] = acc[g][i];
    }
    __syncthreads();
    if (threadIdx.x < G * HD) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < NW; w++) sum += red[(size_t)w * G * HD + threadIdx.x];
      if (rep == a.reps - 1) a.acc[threadIdx.x] = sum;
      else if (sum == 123.456f) a.acc[threadIdx.x] = sum;  // keep the work alive
    }
    __syncthreads();
    if (a.variant == 1) {
#pragma unroll
      for (int g = 0; g < G; g++) out_l[g] = l_run[g];
    }
#pragma unroll
    for (int g = 0; g < G; g++)
#pragma unroll
      for (int i = 0; i < 8; i++) out_acc[g][i] = acc[g][i];
  }
  if (threadIdx.x == 0) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    a.ns[blockIdx.x] = t1 - t0;
    for (int g = 0; g < G; g++) { a.ml[g * 2] = out_m[g]; a.ml[g * 2 + 1] = out_l[g]; }
  }
  if (out_acc[0][0] == 123.456f) a.acc[0] = out_acc[0][0];
}

int main(int argc, char **argv) {
  const int reps = argc > 1 ? atoi(argv[1]) : 200;
  std::vector<T> hK(TILE * HD), hV(TILE * HD);
  std::vector<float> hq(G * HD), fK(TILE * HD), fV(TILE * HD);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
  for (int i = 0; i < TILE * HD; i++) { hK[i] = __float2bfloat16(rnd() * 2.f); fK[i] = __bfloat162float(hK[i]); }
  for (int i = 0; i < TILE * HD; i++) { hV[i] = __float2bfloat16(rnd() * 2.f); fV[i] = __bfloat162float(hV[i]); }
  for (int i = 0; i < G * HD; i++) hq[i] = __bfloat162float(__float2bfloat16(rnd() * 2.f));
  const float scale = 1.0f / sqrtf((float)HD);
  // double-precision reference of o = softmax(q K^T * scale) V for the tile
  std::vector<double> ref(G * HD, 0.0);
  for (int g = 0; g < G; g++) {
    std::vector<double> sc(TILE);
    double mx = -1e300;
    for (int p = 0; p < TILE; p++) {
      double d = 0;
      for (int i = 0; i < HD; i++) d += (double)hq[g * HD + i] * fK[p * HD + i];
      sc[p] = d * scale;
      mx = std::max(mx, sc[p]);
    }
    double L = 0;
    for (int p = 0; p < TILE; p++) { sc[p] = exp(sc[p] - mx); L += sc[p]; }
    for (int p = 0; p < TILE; p++)
      for (int i = 0; i < HD; i++) ref[g * HD + i] += sc[p] / L * fV[p * HD + i];
  }
  T *dK, *dV;
  float *dq, *dacc, *dml;
  unsigned long long *dns;
  CK(cudaMalloc(&dK, TILE * HD * 2)); CK(cudaMalloc(&dV, TILE * HD * 2)); CK(cudaMalloc(&dq, G * HD * 4));
  CK(cudaMalloc(&dacc, G * HD * 4)); CK(cudaMalloc(&dml, G * 2 * 4)); CK(cudaMalloc(&dns, 8));
  CK(cudaMemcpy(dK, hK.data(), TILE * HD * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dV, hV.data(), TILE * HD * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dq, hq.data(), G * HD * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)2 * TILE * HD * 2 + TILE * G * 4 + (size_t)NW * G * HD * 4;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int variant = 0; variant < 2; variant++) {
    Args a{dK, dV, dq, dacc, dml, dns, reps, variant, scale};
    double best = 1e30;
    for (int run = 0; run < 5; run++) {
      k_probe<<<1, NT, smem>>>(a);
      CK(cudaDeviceSynchronize());
      unsigned long long ns;
      CK(cudaMemcpy(&ns, dns, 8, cudaMemcpyDeviceToHost));
      if (run) best = std::min(best, (double)ns / reps);
    }
    std::vector<float> acc(G * HD), ml(G * 2);
    CK(cudaMemcpy(acc.data(), dacc, G * HD * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ml.data(), dml, G * 2 * 4, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int g = 0; g < G; g++)
      for (int i = 0; i < HD; i++) err = std::max(err, fabs((double)acc[g * HD + i] / ml[g * 2 + 1] - ref[g * HD + i]));
    printf("variant %c  %-44s %7.1f ns per tile   max |o - ref| = %.2e  (m0 %.4f l0 %.4f)\n", variant ? 'B' : 'A',
           variant ? "half-warp-private softmax, one merge" : "scores -> softmax -> PV with CTA barriers (product)", best, err, ml[0], ml[1]);
  }
  return 0;
}
