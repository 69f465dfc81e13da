// common.cuh — dtype helpers and the sm_100a PTX wrappers (mbarrier, TMA bulk copy, PDL) shared by
// the kernels of libcake_b200.  sm_100a only; no fallbacks.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cake {

// ------------------------------------------------------------------------------------------- dtype
template <typename T> struct DT;
template <> struct DT<__nv_bfloat16> {
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float f) { return __float2bfloat16_rn(f); }
  // two packed elements (little endian: lo = element 0)
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
};
template <> struct DT<__half> {
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float f) { return __float2half_rn(f); }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&u));
  }
};
// round-trip through D: the "->D" rounding points of the reference
template <typename T> __device__ __forceinline__ float rnd(float f) { return DT<T>::to_f(DT<T>::from_f(f)); }

template <typename T> __device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
  float2 a = DT<T>::unpack2(v.x), b = DT<T>::unpack2(v.y), c = DT<T>::unpack2(v.z), d = DT<T>::unpack2(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b) {
  T x = DT<T>::from_f(a), y = DT<T>::from_f(b);
  return (uint32_t)(*reinterpret_cast<uint16_t *>(&x)) | ((uint32_t)(*reinterpret_cast<uint16_t *>(&y)) << 16);
}

// the same two roundings in ONE packed conversion (cvt.rn.{bf16x2|f16x2}.f32 = F2FP on the ALU pipe; the scalar F2F goes
// through the quarter-rate XU pipe, which is what bound the tcgen05 attention's softmax: ncu r02, XU 81 % busy)
template <typename T> __device__ __forceinline__ uint32_t pack2x(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2x<__nv_bfloat16>(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&v);
}
template <> __device__ __forceinline__ uint32_t pack2x<__half>(float a, float b) {
  const __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t *>(&v);
}

// MLP gate activation, mlp.rs:21-31: act 0 = silu(gate) (cpu/mod.rs:87-89), act 1 = gelu_tanh(gate) (use_gelu_mlp: candle's
// `gelu` is the tanh approximation); the activation is rounded to D before the multiplication by `up` (the caller rounds
// the product): two roundings, as the reference's separate ops produce.
template <typename T> __device__ __forceinline__ float gate_act(float g, int act) {
  if (act) return rnd<T>(0.5f * g * (1.0f + tanhf(0.7978845608028654f * g * (1.0f + 0.044715f * g * g))));
  return rnd<T>(g / (1.0f + expf(-g)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// L2 eviction policy for data that is streamed exactly once per token (weights)
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).  bytes % 16 == 0,
// both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
// Programmatic dependent launch: let the next kernel in the stream start its independent prologue
// (weight prefetch) early / wait until everything the previous kernel wrote is visible.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// scalar D load that bypasses L1 (data produced by another SM during the same kernel)
template <typename T> __device__ __forceinline__ T ldcg_T(const T *p) {
  const unsigned short u = __ldcg(reinterpret_cast<const unsigned short *>(p));
  T r;
  *reinterpret_cast<unsigned short *>(&r) = u;
  return r;
}

__device__ __forceinline__ uint4 ld_nc_v4(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace cake
