// Shared device helpers for libdynscaler_sm100 (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "dynscaler.h"

namespace ds {

constexpr int kBlock = 256;          // threads per CTA for every kernel of the path
constexpr int kSMs = 148;            // B200: grids for grid-stride kernels are multiples of this

// ---- element types ---------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<float> {
  using vec4 = float4;                                    // 16-byte vector of 4 elements
  static __device__ __forceinline__ float to_f(float v) { return v; }
  static __device__ __forceinline__ float from_f(float v) { return v; }
  static __device__ __forceinline__ float one() { return 1.0f; }
};
template <> struct Elem<__nv_bfloat16> {
  using vec4 = uint2;                                     // 8-byte vector of 4 elements
  static __device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  static __device__ __forceinline__ __nv_bfloat16 one() { return __float2bfloat16_rn(1.0f); }
};

template <> struct Elem<__half> {                          // the reference's use_fp16 mode (gen_pano_360.py:191-213)
  using vec4 = uint2;
  static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  static __device__ __forceinline__ __half one() { return __float2half_rn(1.0f); }
};

// read-only (LDG.NC) scalar / vector loads
template <typename T> __device__ __forceinline__ T ldg(const T* p) { return __ldg(p); }
template <> __device__ __forceinline__ __nv_bfloat16 ldg(const __nv_bfloat16* p) {
  return __ushort_as_bfloat16(__ldg(reinterpret_cast<const unsigned short*>(p)));
}

template <> __device__ __forceinline__ __half ldg(const __half* p) {
  return __ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(p)));
}

__device__ __forceinline__ int pos_mod(int a, int m) {   // a in [0, 2m)
  return a >= m ? a - m : a;
}

// mix_latents_with_mask (utils/tensor_utils.py:19-39) on one element, every op rounded on its own and -- like
// the reference's elementwise tensor ops -- in the tensor dtype T:
//   mixed = l1*(1-r) + add*r;  out = l1*(1-m) + mixed*m.   (csrc/pm.cu pm_mix1 is the same chain.)
template <typename T>
__device__ __forceinline__ float mix_point(float l1, float add, float m, float r, float omr) {
  auto rt = [](float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); };
  const float a = rt(__fmul_rn(l1, omr));
  const float b = rt(__fmul_rn(add, r));
  const float mixed = rt(__fadd_rn(a, b));
  const float keep = rt(__fmul_rn(l1, rt(__fsub_rn(1.0f, m))));
  return rt(__fadd_rn(keep, rt(__fmul_rn(mixed, m))));
}

// ---- denoise epilogue of the fused reference-order merge (sphere.cu) ---------------------
// (csrc/pm.cu pm_ddim is the same chain over the 4 planes of a pixel-major atom.)
// float32 scalars of one ddim_step (pipeline/scheduler.py:78-86), formed on the host with the
// reference's torch ops: sqrt(a_t), sqrt(1 - a_t), sqrt(a_prev), sqrt(1 - a_prev - sigma^2), sigma,
// and the classifier-free guidance scale.
struct DdimCoef {
  float sa, s1m, sp, sd, sig, g;
};

// One element of  e = eu + g*(ec - eu)  (t2v_sphere_panorama_pipeline.py:258, t2v_normal_pipeline.py:603),
// pred_x0 = (x - s1m*e)/sa,  x_prev = sp*pred_x0 + sd*e [+ sig*noise]  (scheduler.py:84-94); every
// product / sum / quotient rounded on its own like the reference's elementwise chain.
__device__ __forceinline__ void ddim_point(const DdimCoef& c, float x, float ec, float eu, float nz,
                                           bool has_eu, bool has_nz, float& x_prev, float& pred_x0) {
  float e = ec;
  if (has_eu) e = __fadd_rn(eu, __fmul_rn(c.g, __fsub_rn(ec, eu)));
  const float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.s1m, e)), c.sa);
  float xp = __fadd_rn(__fmul_rn(c.sp, x0), __fmul_rn(c.sd, e));
  if (has_nz) xp = __fadd_rn(xp, __fmul_rn(c.sig, nz));
  pred_x0 = x0;
  x_prev = xp;
}

// Host-side launch helpers ----------------------------------------------------------
inline int check_launch() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? DS_OK : static_cast<int>(e);
}

inline unsigned div_up(long long a, long long b) { return static_cast<unsigned>((a + b - 1) / b); }

}  // namespace ds

#define DS_REQUIRE(cond, code) \
  do {                         \
    if (!(cond)) return (code); \
  } while (0)

#define DS_DISPATCH_DTYPE(dtype, ...)                   \
  do {                                                  \
    if ((dtype) == DS_F32) {                            \
      using ET = float;                                  \
      __VA_ARGS__                                       \
    } else if ((dtype) == DS_BF16) {                    \
      using ET = __nv_bfloat16;                          \
      __VA_ARGS__                                       \
    } else if ((dtype) == DS_F16) {                     \
      using ET = __half;                                 \
      __VA_ARGS__                                       \
    } else {                                            \
      return DS_ERR_DTYPE;                              \
    }                                                   \
  } while (0)
