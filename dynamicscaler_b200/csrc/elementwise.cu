// Elementwise steps either side of the tiling path: mix_latents_with_mask
// (reference utils/tensor_utils.py:19-39) and the final select of the sharded merge
// (SURVEY §8e).  Grid-stride, 16-byte vectors when everything is aligned.
#include <type_traits>

#include "ds_common.cuh"

namespace ds {

// out = l1*(1-m) + (l1*(1-r) + add*r)*m with every product / sum rounded on its own,
// exactly like the reference's chain of torch elementwise ops.
__device__ __forceinline__ float mix_one(float l1, float add, float m, float r, float omr) {
  const float mixed = __fadd_rn(__fmul_rn(l1, omr), __fmul_rn(add, r));
  return __fadd_rn(__fmul_rn(l1, __fsub_rn(1.0f, m)), __fmul_rn(mixed, m));
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
mix_mask_kernel(const T* __restrict__ l1, const T* __restrict__ add, const T* __restrict__ mask,
                T* __restrict__ out, float r, float omr, long long n, int mask_hw) {
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n; i += stride) {
    const long long mi = mask_hw > 0 ? i % mask_hw : i;
    out[i] = Elem<T>::from_f(mix_one(Elem<T>::to_f(ldg(l1 + i)), Elem<T>::to_f(ldg(add + i)),
                                     Elem<T>::to_f(ldg(mask + mi)), r, omr));
  }
}

// fp32, n % 4 == 0, mask_hw % 4 == 0, all pointers 16-byte aligned
__global__ void __launch_bounds__(kBlock)
mix_mask_f32x4_kernel(const float4* __restrict__ l1, const float4* __restrict__ add,
                      const float4* __restrict__ mask, float4* __restrict__ out, float r, float omr,
                      long long n4, int mask_hw4) {
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(l1 + i), b = __ldg(add + i);
    const float4 m = __ldg(mask + (mask_hw4 > 0 ? i % mask_hw4 : i));
    float4 o;
    o.x = mix_one(a.x, b.x, m.x, r, omr);
    o.y = mix_one(a.y, b.y, m.y, r, omr);
    o.z = mix_one(a.z, b.z, m.z, r, omr);
    o.w = mix_one(a.w, b.w, m.w, r, omr);
    out[i] = o;
  }
}

// The read side of the planar I2V loop after the two window gathers (i2v_sphere_panorama_pipeline.py:875-884):
//   noised = c*x + s*eps              re_noise, pipeline/scheduler.py:98-110, epsilon injected
//   out    = mix_latents_with_mask(x, noised, m, r)
// in ONE pass over the window tensors.  c, s are 0-dim float32 tensors in the reference and are cast to the
// latent dtype by the multiply; every op is rounded on its own in T (csrc/pm.cu pm_renoise_pass is the same chain).
template <typename T, int VEC>
__global__ void __launch_bounds__(kBlock)
renoise_mix_kernel(const T* __restrict__ x, const T* __restrict__ eps, const T* __restrict__ mask,
                   T* __restrict__ out, float c, float s, float r, float omr, long long n_vec,
                   long long win_vec, int hw_vec) {
  using V = typename std::conditional<VEC == 4, typename Elem<T>::vec4, T>::type;
  auto rt = [](float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); };
  const float ct = rt(c), st = rt(s);
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n_vec; i += stride) {
    const V vx = __ldg(reinterpret_cast<const V*>(x) + i);
    const V ve = __ldg(reinterpret_cast<const V*>(eps) + i);
    // hw_vec > 0: every window is mixed with plane (batch 0, channel 0, frame 0) of its own mask window
    const long long mi = hw_vec > 0 ? (i / win_vec) * win_vec + i % hw_vec : i;
    const V vm = __ldg(reinterpret_cast<const V*>(mask) + mi);
    V vo;
    const T* ex = reinterpret_cast<const T*>(&vx);
    const T* ee = reinterpret_cast<const T*>(&ve);
    const T* em = reinterpret_cast<const T*>(&vm);
    T* eo = reinterpret_cast<T*>(&vo);
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const float xv = Elem<T>::to_f(ex[e]);
      const float noised = rt(__fadd_rn(rt(__fmul_rn(ct, xv)), rt(__fmul_rn(st, Elem<T>::to_f(ee[e])))));
      eo[e] = Elem<T>::from_f(mix_point<T>(xv, noised, Elem<T>::to_f(em[e]), r, omr));
    }
    reinterpret_cast<V*>(out)[i] = vo;
  }
}

template <typename T>
__global__ void __launch_bounds__(kBlock)
select_merged_kernel(T* __restrict__ pano, const T* __restrict__ sum, const T* __restrict__ weight,
                     long long n, long long period, long long group) {
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n; i += stride) {
    const float w = Elem<T>::to_f(ldg(weight + (i / group) % period));
    if (w > 0.0f) pano[i] = ldg(sum + i);
  }
}

// fp32, n % 4 == 0, 16-byte aligned, group == 1 and period % 4 == 0: whole float4 groups
__global__ void __launch_bounds__(kBlock)
select_merged_f32x4_kernel(float4* __restrict__ pano, const float4* __restrict__ sum,
                           const float4* __restrict__ weight, long long n4, long long period4) {
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n4; i += stride) {
    const float4 w = __ldg(weight + i % period4);
    if (w.x > 0.0f || w.y > 0.0f || w.z > 0.0f || w.w > 0.0f) {
      const float4 s = __ldg(sum + i);
      float4 p = pano[i];
      p.x = w.x > 0.0f ? s.x : p.x;
      p.y = w.y > 0.0f ? s.y : p.y;
      p.z = w.z > 0.0f ? s.z : p.z;
      p.w = w.w > 0.0f ? s.w : p.w;
      pano[i] = p;
    }
  }
}

// fp32, n % 4 == 0, 16-byte aligned, group % 4 == 0: one weight per float4 (pixel-major rows)
__global__ void __launch_bounds__(kBlock)
select_merged_f32x4_grouped_kernel(float4* __restrict__ pano, const float4* __restrict__ sum,
                                   const float* __restrict__ weight, long long n4, long long period,
                                   long long group4) {
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n4; i += stride) {
    if (__ldg(weight + (i / group4) % period) > 0.0f) pano[i] = __ldg(sum + i);
  }
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static unsigned stride_grid(long long work_items) {
  // enough CTAs for 8 resident per SM, in multiples of the SM count
  const long long want = (work_items + kBlock - 1) / kBlock;
  const long long cap = static_cast<long long>(kSMs) * 8;
  if (want >= cap) return static_cast<unsigned>(cap);
  return static_cast<unsigned>(want > 0 ? want : 1);
}

}  // namespace ds

extern "C" int ds_mix_mask(const void* l1, const void* add, const void* mask, void* out,
                           float ratio, float one_minus_ratio, long long n, int mask_hw, int dtype,
                           void* stream) {
  using namespace ds;
  DS_REQUIRE(l1 != nullptr && add != nullptr && mask != nullptr && out != nullptr, DS_ERR_NULL);
  DS_REQUIRE(n > 0 && mask_hw >= 0, DS_ERR_DIM);
  DS_REQUIRE(mask_hw == 0 || n % mask_hw == 0, DS_ERR_DIM);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == DS_F32 && n % 4 == 0 && mask_hw % 4 == 0 && aligned16(l1) && aligned16(add) &&
      aligned16(mask) && aligned16(out)) {
    mix_mask_f32x4_kernel<<<stride_grid(n / 4), kBlock, 0, st>>>(
        static_cast<const float4*>(l1), static_cast<const float4*>(add), static_cast<const float4*>(mask),
        static_cast<float4*>(out), ratio, one_minus_ratio, n / 4, mask_hw / 4);
    return check_launch();
  }
  DS_DISPATCH_DTYPE(dtype, {
    mix_mask_kernel<ET><<<stride_grid(n), kBlock, 0, st>>>(static_cast<const ET*>(l1), static_cast<const ET*>(add),
                                                            static_cast<const ET*>(mask), static_cast<ET*>(out),
                                                            ratio, one_minus_ratio, n, mask_hw);
  });
  return check_launch();
}

extern "C" int ds_renoise_mix(const void* x, const void* eps, const void* mask, void* out, float c, float s,
                              float ratio, float one_minus_ratio, long long n, long long window_elems,
                              int mask_plane_hw, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(x != nullptr && eps != nullptr && mask != nullptr && out != nullptr, DS_ERR_NULL);
  DS_REQUIRE(n > 0 && mask_plane_hw >= 0, DS_ERR_DIM);
  DS_REQUIRE(mask_plane_hw == 0 || (window_elems > 0 && n % window_elems == 0 && window_elems % mask_plane_hw == 0),
             DS_ERR_DIM);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DS_DISPATCH_DTYPE(dtype, {
    const uintptr_t al = sizeof(typename Elem<ET>::vec4) - 1;
    const bool vec = n % 4 == 0 && mask_plane_hw % 4 == 0 &&
                     ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(eps) |
                       reinterpret_cast<uintptr_t>(mask) | reinterpret_cast<uintptr_t>(out)) & al) == 0;
    if (vec)
      renoise_mix_kernel<ET, 4><<<stride_grid(n / 4), kBlock, 0, st>>>(
          static_cast<const ET*>(x), static_cast<const ET*>(eps), static_cast<const ET*>(mask), static_cast<ET*>(out),
          c, s, ratio, one_minus_ratio, n / 4, mask_plane_hw > 0 ? window_elems / 4 : 1, mask_plane_hw / 4);
    else
      renoise_mix_kernel<ET, 1><<<stride_grid(n), kBlock, 0, st>>>(
          static_cast<const ET*>(x), static_cast<const ET*>(eps), static_cast<const ET*>(mask), static_cast<ET*>(out),
          c, s, ratio, one_minus_ratio, n, mask_plane_hw > 0 ? window_elems : 1, mask_plane_hw);
  });
  return check_launch();
}

extern "C" int ds_select_merged(void* pano, const void* sum, const void* weight, long long n,
                                long long weight_period, long long weight_group, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(pano != nullptr && sum != nullptr && weight != nullptr, DS_ERR_NULL);
  DS_REQUIRE(n > 0 && weight_period > 0 && weight_group > 0, DS_ERR_DIM);
  DS_REQUIRE(n % (weight_period * weight_group) == 0, DS_ERR_DIM);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec_ok = dtype == DS_F32 && n % 4 == 0 && aligned16(pano) && aligned16(sum);
  if (vec_ok && weight_group == 1 && weight_period % 4 == 0 && aligned16(weight)) {
    select_merged_f32x4_kernel<<<stride_grid(n / 4), kBlock, 0, st>>>(
        static_cast<float4*>(pano), static_cast<const float4*>(sum), static_cast<const float4*>(weight),
        n / 4, weight_period / 4);
    return check_launch();
  }
  if (vec_ok && weight_group % 4 == 0) {
    select_merged_f32x4_grouped_kernel<<<stride_grid(n / 4), kBlock, 0, st>>>(
        static_cast<float4*>(pano), static_cast<const float4*>(sum), static_cast<const float*>(weight),
        n / 4, weight_period, weight_group / 4);
    return check_launch();
  }
  DS_DISPATCH_DTYPE(dtype, {
    select_merged_kernel<ET><<<stride_grid(n), kBlock, 0, st>>>(static_cast<ET*>(pano), static_cast<const ET*>(sum),
                                                                 static_cast<const ET*>(weight), n, weight_period,
                                                                 weight_group);
  });
  return check_launch();
}

extern "C" int ds_abi_version(void) { return DS_ABI_VERSION; }

extern "C" const char* ds_error_string(int code) {
  switch (code) {
    case DS_OK: return "ok";
    case DS_ERR_NULL: return "dynscaler: required pointer is NULL";
    case DS_ERR_DIM: return "dynscaler: bad or inconsistent dimension";
    case DS_ERR_DTYPE: return "dynscaler: unknown dtype";
    case DS_ERR_WINDOW: return "dynscaler: illegal ring window";
    case DS_ERR_COUNT: return "dynscaler: bad payload count";
    case DS_ERR_ALIGN: return "dynscaler: vectorised variant on unaligned data";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "dynscaler: unknown error";
  }
}
