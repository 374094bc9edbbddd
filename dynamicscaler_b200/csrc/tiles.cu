// Pixel-side neighbours of the tiling path: the data movement around the VAE calls
// of the I2V sphere pipeline.  The VAE itself stays the reference's torch module.
//
//   tile_accumulate_kernel / tile_normalise_kernel
//       tiled_vae_encode_tensor_simple (pipeline/i2v_sphere_panorama_pipeline.py:505-562): every
//       encoded tile is cropped to its core and ADDED into a float32 accumulator, a count plane is
//       incremented, and the result is acc / clamp(count, 1) -- the only literal sum / count
//       normalisation in the reference.
//   repeat_renoise_kernel
//       paste_on_static (i2v_...:245-254): the encoded frame-0 latent repeated over total_f frames
//       (torch.cat([frame_0_latent] * total_f, dim=2)) and re-noised (pipeline/scheduler.py:98-110,
//       epsilon injected) in one pass, instead of a cat + 3 elementwise passes.
// The wrap-padded chunked decode (i2v_...:974-991) needs no kernel of its own: padding and cropping
// are ring windows (ds_ring_gather).
// All HBM-bound elementwise / copy work: grid-stride, coalesced, 16-byte vectors where aligned.
#include "ds_common.cuh"

namespace ds {

// acc[pl][dst_top + y][dst_left + x] += tile[pl][top_cut + y][left_cut + x]   (y < ch, x < cw)
// count[dst_top + y][dst_left + x] += 1                                        (once, by plane 0)
template <typename T>
__global__ void __launch_bounds__(kBlock)
tile_accumulate_kernel(float* __restrict__ acc, float* __restrict__ count, const T* __restrict__ tile,
                       long long planes, int H, int W, int th, int tw, int top_cut, int left_cut, int ch, int cw,
                       int dst_top, int dst_left, long long count_stride, long long planes_per_count) {
  const long long per_plane = static_cast<long long>(ch) * cw;
  const long long n = planes * per_plane;
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n; i += stride) {
    const long long pl = i / per_plane;
    const int r = static_cast<int>(i - pl * per_plane);
    const int y = r / cw, x = r - y * cw;
    const float v = Elem<T>::to_f(ldg(tile + (pl * th + top_cut + y) * tw + left_cut + x));
    float* a = acc + (pl * H + dst_top + y) * W + dst_left + x;
    *a = __fadd_rn(*a, v);
    if (pl % planes_per_count == 0) {
      float* c = count + (pl / planes_per_count) * count_stride + static_cast<long long>(dst_top + y) * W + dst_left + x;
      *c = __fadd_rn(*c, 1.0f);
    }
  }
}

// out[pl][p] = acc[pl][p] / max(count[b(pl)][p], 1)
__global__ void __launch_bounds__(kBlock)
tile_normalise_kernel(const float* __restrict__ acc, const float* __restrict__ count, float* __restrict__ out,
                      long long planes, long long HW, long long planes_per_count) {
  const long long n = planes * HW;
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n; i += stride) {
    const long long pl = i / HW, p = i - pl * HW;
    const float c = fmaxf(__ldg(count + (pl / planes_per_count) * HW + p), 1.0f);
    out[i] = __fdiv_rn(__ldg(acc + i), c);
  }
}

// out[pl][f][p] = c * x0[pl][p] + s * eps[pl][f][p]      (every op rounded on its own, in T)
template <typename T>
__global__ void __launch_bounds__(kBlock)
repeat_renoise_kernel(const T* __restrict__ x0, const T* __restrict__ eps, T* __restrict__ out, float c, float s,
                      long long planes, int F, long long HW) {
  const long long n = planes * F * HW;
  const long long stride = static_cast<long long>(gridDim.x) * kBlock;
  auto rt = [](float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); };
  // the reference's c and s are 0-dim float32 tensors (scheduler.py:104-105): in `c * x_a` torch casts a 0-dim
  // operand to the dimensioned operand's dtype first, so for half latents the coefficients are rounded to T
  c = rt(c);
  s = rt(s);
  for (long long i = static_cast<long long>(blockIdx.x) * kBlock + threadIdx.x; i < n; i += stride) {
    const long long pl = i / (F * HW), p = i % HW;
    const float a = rt(__fmul_rn(c, Elem<T>::to_f(ldg(x0 + pl * HW + p))));
    const float b = eps != nullptr ? rt(__fmul_rn(s, Elem<T>::to_f(ldg(eps + i)))) : 0.f;
    out[i] = Elem<T>::from_f(eps != nullptr ? __fadd_rn(a, b) : Elem<T>::to_f(ldg(x0 + pl * HW + p)));
  }
}

static unsigned tiles_grid(long long work_items) {
  const long long want = (work_items + kBlock - 1) / kBlock;
  const long long cap = static_cast<long long>(kSMs) * 8;
  return static_cast<unsigned>(want >= cap ? cap : (want > 0 ? want : 1));
}

}  // namespace ds

extern "C" int ds_tile_accumulate(float* acc, float* count, const void* tile, long long planes, long long planes_per_count,
                                  int H, int W, int th, int tw, int top_cut, int left_cut, int ch, int cw,
                                  int dst_top, int dst_left, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(acc != nullptr && count != nullptr && tile != nullptr, DS_ERR_NULL);
  DS_REQUIRE(planes > 0 && planes_per_count > 0 && planes % planes_per_count == 0, DS_ERR_DIM);
  DS_REQUIRE(H > 0 && W > 0 && th > 0 && tw > 0 && ch > 0 && cw > 0, DS_ERR_DIM);
  DS_REQUIRE(top_cut >= 0 && left_cut >= 0 && top_cut + ch <= th && left_cut + cw <= tw, DS_ERR_WINDOW);
  DS_REQUIRE(dst_top >= 0 && dst_left >= 0 && dst_top + ch <= H && dst_left + cw <= W, DS_ERR_WINDOW);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = planes * ch * cw;
  DS_DISPATCH_DTYPE(dtype, {
    tile_accumulate_kernel<ET><<<tiles_grid(n), kBlock, 0, st>>>(acc, count, static_cast<const ET*>(tile), planes, H, W,
                                                                  th, tw, top_cut, left_cut, ch, cw, dst_top, dst_left,
                                                                  static_cast<long long>(H) * W, planes_per_count);
  });
  return check_launch();
}

extern "C" int ds_tile_normalise(const float* acc, const float* count, float* out, long long planes,
                                 long long planes_per_count, long long HW, void* stream) {
  using namespace ds;
  DS_REQUIRE(acc != nullptr && count != nullptr && out != nullptr, DS_ERR_NULL);
  DS_REQUIRE(planes > 0 && planes_per_count > 0 && planes % planes_per_count == 0 && HW > 0, DS_ERR_DIM);
  tile_normalise_kernel<<<tiles_grid(planes * HW), kBlock, 0, static_cast<cudaStream_t>(stream)>>>(
      acc, count, out, planes, HW, planes_per_count);
  return check_launch();
}

extern "C" int ds_repeat_renoise(const void* x0, const void* eps, void* out, float c, float s, long long planes, int F,
                                 long long HW, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(x0 != nullptr && out != nullptr, DS_ERR_NULL);
  DS_REQUIRE(planes > 0 && F > 0 && HW > 0, DS_ERR_DIM);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DS_DISPATCH_DTYPE(dtype, {
    repeat_renoise_kernel<ET><<<tiles_grid(planes * F * HW), kBlock, 0, st>>>(
        static_cast<const ET*>(x0), static_cast<const ET*>(eps), static_cast<ET*>(out), c, s, planes, F, HW);
  });
  return check_launch();
}
