// Per-frame spatial resize of a video latent: the CUDA side of resize_video_latent
// (reference utils/diffusion_utils.py:21-33 = F.interpolate on [B*F, C, h, w]; modes 'nearest',
// 'bilinear', 'bicubic').  Used at the stage transitions (gen_pano_360.py:287-289 nearest
// down, :345-347 bicubic x2 up) and in the loop when view_get/set_scale_factor != 1
// (t2v_sphere_panorama_pipeline.py:200-203,268-274).
//
// Every (b, c, f) plane is resized independently, so the tensor is addressed as [planes][h][w]
// whatever order the reference's permutes describe.  Index and weight arithmetic follows ATen's
// float32 formulas (UpSample.h: nearest_neighbor_compute_source_index, area_pixel_compute_
// source_index, get_cubic_upsample_coefficients with A = -0.75) so nearest is bit-exact and the
// interpolating modes agree to ~1e-6.  HBM-bound; a thread produces 4 consecutive output pixels
// of one plane row (16-byte store).
#include "ds_common.cuh"

namespace ds {

struct ResizeDims {
  int planes, h, w, H, W;
  float scale_h, scale_w;
  int align;
};

__device__ __forceinline__ float src_index(float scale, int dst, bool align, bool cubic) {
  if (align) return scale * dst;
  const float s = scale * (dst + 0.5f) - 0.5f;
  return (!cubic && s < 0.f) ? 0.f : s;
}

__device__ __forceinline__ void cubic_coeffs(float t, float (&c)[4]) {
  const float A = -0.75f;
  const float x0 = t + 1.f, x1 = t, x2 = 1.f - t, x3 = 2.f - t;
  c[0] = ((A * x0 - 5.f * A) * x0 + 8.f * A) * x0 - 4.f * A;
  c[1] = ((A + 2.f) * x1 - (A + 3.f)) * x1 * x1 + 1.f;
  c[2] = ((A + 2.f) * x2 - (A + 3.f)) * x2 * x2 + 1.f;
  c[3] = ((A * x3 - 5.f * A) * x3 + 8.f * A) * x3 - 4.f * A;
}

// MODE: 0 nearest, 1 bilinear, 2 bicubic
template <typename T, int MODE>
__global__ void __launch_bounds__(kBlock)
resize_kernel(const T* __restrict__ in, T* __restrict__ out, ResizeDims d) {
  const int xg = (d.W + 3) / 4;
  const long long total = static_cast<long long>(d.planes) * d.H * xg;
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long g = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; g < total; g += stride) {
    const int x0 = static_cast<int>(g % xg) * 4;
    const long long rest = g / xg;
    const int y = static_cast<int>(rest % d.H);
    const long long pl = rest / d.H;
    const T* __restrict__ src = in + pl * d.h * d.w;
    float r[4];
    if (MODE == 0) {
      const int sy = min(static_cast<int>(floorf(y * d.scale_h)), d.h - 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int x = min(x0 + e, d.W - 1);
        const int sx = min(static_cast<int>(floorf(x * d.scale_w)), d.w - 1);
        r[e] = Elem<T>::to_f(ldg(src + sy * d.w + sx));
      }
    } else if (MODE == 1) {
      const float fy = src_index(d.scale_h, y, d.align, false);
      const int y0 = static_cast<int>(fy), y1 = y0 + (y0 < d.h - 1 ? 1 : 0);
      const float ly1 = fy - y0, ly0 = 1.f - ly1;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int x = min(x0 + e, d.W - 1);
        const float fx = src_index(d.scale_w, x, d.align, false);
        const int xa = static_cast<int>(fx), xb = xa + (xa < d.w - 1 ? 1 : 0);
        const float lx1 = fx - xa, lx0 = 1.f - lx1;
        const float v00 = Elem<T>::to_f(ldg(src + y0 * d.w + xa)), v01 = Elem<T>::to_f(ldg(src + y0 * d.w + xb));
        const float v10 = Elem<T>::to_f(ldg(src + y1 * d.w + xa)), v11 = Elem<T>::to_f(ldg(src + y1 * d.w + xb));
        r[e] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
      }
    } else {
      const float fy = src_index(d.scale_h, y, d.align, true);
      const float flo_y = floorf(fy);
      const int iy = static_cast<int>(flo_y);
      float cy[4];
      cubic_coeffs(fy - flo_y, cy);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int x = min(x0 + e, d.W - 1);
        const float fx = src_index(d.scale_w, x, d.align, true);
        const float flo_x = floorf(fx);
        const int ix = static_cast<int>(flo_x);
        float cx[4];
        cubic_coeffs(fx - flo_x, cx);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int yy = min(max(iy - 1 + i, 0), d.h - 1);
          float row = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int xx = min(max(ix - 1 + j, 0), d.w - 1);
            row += Elem<T>::to_f(ldg(src + yy * d.w + xx)) * cx[j];
          }
          acc += row * cy[i];
        }
        r[e] = acc;
      }
    }
    T* __restrict__ o = out + (pl * d.H + y) * d.W + x0;
    if (x0 + 3 < d.W && (reinterpret_cast<uintptr_t>(o) & (4 * sizeof(T) - 1)) == 0) {
      T v[4] = {Elem<T>::from_f(r[0]), Elem<T>::from_f(r[1]), Elem<T>::from_f(r[2]), Elem<T>::from_f(r[3])};
      *reinterpret_cast<typename Elem<T>::vec4*>(o) = *reinterpret_cast<const typename Elem<T>::vec4*>(v);
    } else {
      for (int e = 0; e < 4 && x0 + e < d.W; ++e) o[e] = Elem<T>::from_f(r[e]);
    }
  }
}

}  // namespace ds

extern "C" int ds_resize_planes(const void* in, void* out, long long planes, int h, int w, int H, int W,
                                int mode, int align_corners, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(in != nullptr && out != nullptr, DS_ERR_NULL);
  DS_REQUIRE(planes > 0 && planes < (1ll << 31) && h > 0 && w > 0 && H > 0 && W > 0, DS_ERR_DIM);
  DS_REQUIRE(mode == DS_RESIZE_NEAREST || mode == DS_RESIZE_BILINEAR || mode == DS_RESIZE_BICUBIC, DS_ERR_DIM);
  DS_REQUIRE(static_cast<long long>(h) * w < (1ll << 31), DS_ERR_DIM);
  ResizeDims d;
  d.planes = static_cast<int>(planes); d.h = h; d.w = w; d.H = H; d.W = W;
  d.align = (mode != DS_RESIZE_NEAREST && align_corners) ? 1 : 0;
  // ATen area_pixel_compute_scale / compute_scales_value (no scale_factor given): float division
  if (d.align) {
    d.scale_h = H > 1 ? static_cast<float>(h - 1) / (H - 1) : 0.f;
    d.scale_w = W > 1 ? static_cast<float>(w - 1) / (W - 1) : 0.f;
  } else {
    d.scale_h = static_cast<float>(h) / H;
    d.scale_w = static_cast<float>(w) / W;
  }
  const long long groups = planes * H * ((W + 3) / 4);
  const long long want = (groups + kBlock - 1) / kBlock;
  const unsigned grid = static_cast<unsigned>(want < kSMs * 16 ? (want > 0 ? want : 1) : kSMs * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DS_DISPATCH_DTYPE(dtype, {
    const ET* ip = static_cast<const ET*>(in);
    ET* op = static_cast<ET*>(out);
    if (mode == DS_RESIZE_NEAREST) resize_kernel<ET, 0><<<grid, kBlock, 0, st>>>(ip, op, d);
    else if (mode == DS_RESIZE_BILINEAR) resize_kernel<ET, 1><<<grid, kBlock, 0, st>>>(ip, op, d);
    else resize_kernel<ET, 2><<<grid, kBlock, 0, st>>>(ip, op, d);
  });
  return check_launch();
}
