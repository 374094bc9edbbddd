// Planar wrap-around ("ring") windows of the panorama latent: the CUDA side of
// RingLatent.get_window_latent / set_window_latent and RingImageTensor.get_window_tensor
// (reference utils/shift_window_utils.py:48-114, 116-206, 232-263).
//
// Pure HBM-bound copies.  One CTA moves one (window, plane, frame) tile of h x w
// elements; with W, w and the window's left edge on multiples of 4 every access is a
// 16-byte (fp32) vector, otherwise the scalar path applies the modulo per element.
#include "ds_common.cuh"

namespace ds {

struct RingDims {
  int P, F, H, W, f_len, h, w;
};

// Generic tile copy: any window size / alignment, modulo per element.  Kept out of line so
// its registers do not burden the fast path.
template <typename T, int VEC, bool GATHER>
__device__ __forceinline__ void ring_tile_generic(T* __restrict__ src_plane, T* __restrict__ win_tile,
                                               int left, int top, const RingDims& d) {
  using V4 = typename Elem<T>::vec4;
  const int wg = d.w / VEC;
  const int groups = d.h * wg;
  constexpr int U = 4;                            // independent accesses in flight per thread
  for (int base = threadIdx.x; base < groups; base += kBlock * U) {
    int y[U], x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = base + u * kBlock;
      y[u] = g / wg;
      x[u] = (g - y[u] * wg) * VEC;
    }
    if (VEC == 4 && (left & 3) == 0) {
      V4 v[U];
      if (GATHER) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (base + u * kBlock < groups) {
            const int r = (top + y[u]) % d.H, c = (left + x[u]) % d.W;
            v[u] = __ldg(reinterpret_cast<const V4*>(src_plane + static_cast<size_t>(r) * d.W + c));
          }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (base + u * kBlock < groups)
            *reinterpret_cast<V4*>(win_tile + static_cast<size_t>(y[u]) * d.w + x[u]) = v[u];
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (base + u * kBlock < groups)
            v[u] = __ldg(reinterpret_cast<const V4*>(win_tile + static_cast<size_t>(y[u]) * d.w + x[u]));
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (base + u * kBlock < groups) {
            const int r = (top + y[u]) % d.H, c = (left + x[u]) % d.W;
            *reinterpret_cast<V4*>(src_plane + static_cast<size_t>(r) * d.W + c) = v[u];
          }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (base + u * kBlock >= groups) continue;
        const int r = (top + y[u]) % d.H;
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const int c = (left + x[u] + e) % d.W;
          T* pp = src_plane + static_cast<size_t>(r) * d.W + c;
          T* wp = win_tile + static_cast<size_t>(y[u]) * d.w + x[u] + e;
          if (GATHER) *wp = ldg(pp); else *pp = ldg(wp);
        }
      }
    }
  }
}

// VEC = 4: w % 4 == 0 and W % 4 == 0 (host-checked); a window whose left edge is not a
// multiple of 4 falls back to scalar accesses (warp-uniform branch).
// Persistent CTAs loop over the (window, plane, frame) tiles.  A thread owns a fixed column
// group and walks down the rows, so the only per-element index work is one compare-subtract
// wrap per axis (no division): the kernel is otherwise a plain 16-byte copy.
// WG > 0: w / VEC known at compile time (power of two dividing the CTA), else runtime.
template <typename T, int VEC, bool GATHER, int WG>
__global__ void __launch_bounds__(kBlock)
ring_copy_kernel(T* __restrict__ pano, T* __restrict__ win, const int32_t* __restrict__ d_wins,
                 int h_left, int h_top, int h_f0, RingDims d, int n_tiles) {
  using V4 = typename Elem<T>::vec4;
  constexpr int WGS = WG > 0 ? WG : 1;            // column groups per window row (fast path only)
  constexpr int rows_per_pass = kBlock / WGS;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    // tile = (window j, plane pl, window frame fo)
    const int fo = tile % d.f_len;
    const int pl = (tile / d.f_len) % d.P;
    const int j = tile / (d.f_len * d.P);
    int left, top, f0;
    if (d_wins != nullptr) {
      left = __ldg(d_wins + 3 * j);
      top = __ldg(d_wins + 3 * j + 1);
      f0 = __ldg(d_wins + 3 * j + 2);
    } else {
      left = h_left; top = h_top; f0 = h_f0;
    }
    left %= d.W; top %= d.H;                       // begin may lie in [size, 2*size)
    const int f = (f0 + fo) % d.F;
    T* __restrict__ src_plane = pano + (static_cast<size_t>(pl) * d.F + f) * d.H * d.W;
    T* __restrict__ win_tile = win + static_cast<size_t>(tile) * d.h * d.w;

    if (WG > 0 && VEC == 4 && (left & 3) == 0 && d.w <= d.W && d.h <= d.H) {
      // fast path: fixed column group per thread, rows strided by rows_per_pass
      const int xg = threadIdx.x % WGS, y0 = threadIdx.x / WGS;
      int c = left + xg * VEC;
      c = c >= d.W ? c - d.W : c;
      constexpr int U = 4;
      for (int yb = y0; yb < d.h; yb += rows_per_pass * U) {
        V4 v[U];
        int r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int y = yb + u * rows_per_pass;
          r[u] = top + y;
          r[u] = r[u] >= d.H ? r[u] - d.H : r[u];
          if (y < d.h) {
            if (GATHER) v[u] = __ldg(reinterpret_cast<const V4*>(src_plane + static_cast<size_t>(r[u]) * d.W + c));
            else v[u] = __ldg(reinterpret_cast<const V4*>(win_tile + y * d.w + xg * VEC));
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int y = yb + u * rows_per_pass;
          if (y < d.h) {
            if (GATHER) *reinterpret_cast<V4*>(win_tile + y * d.w + xg * VEC) = v[u];
            else *reinterpret_cast<V4*>(src_plane + static_cast<size_t>(r[u]) * d.W + c) = v[u];
          }
        }
      }
      continue;
    }

    if (WG > 0) {
      // rare for this instantiation (left edge off the 4-grid, or window larger than the panorama):
      // compact scalar loop, modulo per element
      for (int g = threadIdx.x; g < d.h * d.w; g += kBlock) {
        const int y = g / d.w, x = g - y * d.w;
        T* pp = src_plane + static_cast<size_t>((top + y) % d.H) * d.W + (left + x) % d.W;
        if (GATHER) win_tile[g] = ldg(pp); else *pp = ldg(win_tile + g);
      }
    } else {
      ring_tile_generic<T, VEC, GATHER>(src_plane, win_tile, left, top, d);
    }
  }
}

// plan_src[p] = j*h*w + y*w + x of the last window (highest j) covering p, else -1.
__global__ void __launch_bounds__(kBlock)
ring_merge_plan_kernel(int32_t* __restrict__ plan_src, const int32_t* __restrict__ d_wins, int n,
                       int H, int W, int h, int w) {
  extern __shared__ int32_t s_wins[];             // [n][2] reduced modulo (H, W)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_wins[2 * i] = __ldg(d_wins + 2 * i) % W;
    s_wins[2 * i + 1] = __ldg(d_wins + 2 * i + 1) % H;
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= H * W) return;
  const int r = p / W, c = p - r * W;
  int src = -1;
  for (int j = n - 1; j >= 0; --j) {
    int dx = c - s_wins[2 * j];
    dx += dx < 0 ? W : 0;
    if (dx >= w) continue;
    int dy = r - s_wins[2 * j + 1];
    dy += dy < 0 ? H : 0;
    if (dy >= h) continue;
    src = (j * h + dy) * w + dx;
    break;
  }
  plan_src[p] = src;
}

template <bool GATHER>
static int ring_copy(void* pano, void* win, const int32_t* d_wins, int n, int left, int top, int f0,
                     int P, int F, int H, int W, int f_len, int h, int w, int dtype, void* stream) {
  DS_REQUIRE(pano != nullptr && win != nullptr, DS_ERR_NULL);
  DS_REQUIRE(n > 0 && P > 0 && F > 0 && H > 0 && W > 0 && f_len > 0 && h > 0 && w > 0, DS_ERR_DIM);
  DS_REQUIRE(d_wins != nullptr || n == 1, DS_ERR_DIM);
  if (d_wins == nullptr)
    DS_REQUIRE(left >= 0 && left < 2 * W && top >= 0 && top < 2 * H && f0 >= 0 && f0 < 2 * F &&
               left + w <= 2 * W && top + h <= 2 * H && f0 + f_len <= 2 * F, DS_ERR_WINDOW);
  if (!GATHER) DS_REQUIRE(f_len <= F && h <= H && w <= W, DS_ERR_WINDOW);
  const long long tiles = static_cast<long long>(n) * P * f_len;
  DS_REQUIRE(tiles < (1ll << 31), DS_ERR_DIM);
  RingDims d{P, F, H, W, f_len, h, w};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (w % 4 == 0) && (W % 4 == 0);
  // persistent: at most 8 CTAs per SM (the kernels use <= 32 registers), each looping over tiles
  const unsigned grid = static_cast<unsigned>(tiles < static_cast<long long>(kSMs) * 8 ? tiles : kSMs * 8);
  const int nt = static_cast<int>(tiles);
  DS_DISPATCH_DTYPE(dtype, {
    ET* pp = static_cast<ET*>(pano);
    ET* wp = static_cast<ET*>(win);
    if (vec && w == 64)
      ring_copy_kernel<ET, 4, GATHER, 16><<<grid, kBlock, 0, st>>>(pp, wp, d_wins, left, top, f0, d, nt);
    else if (vec)
      ring_copy_kernel<ET, 4, GATHER, 0><<<grid, kBlock, 0, st>>>(pp, wp, d_wins, left, top, f0, d, nt);
    else
      ring_copy_kernel<ET, 1, GATHER, 0><<<grid, kBlock, 0, st>>>(pp, wp, d_wins, left, top, f0, d, nt);
  });
  return check_launch();
}

}  // namespace ds

extern "C" int ds_ring_gather(const void* pano, void* out, const int32_t* d_wins, int n,
                              int h_left, int h_top, int h_f0, int P, int F, int H, int W,
                              int f_len, int h, int w, int dtype, void* stream) {
  return ds::ring_copy<true>(const_cast<void*>(pano), out, d_wins, n, h_left, h_top, h_f0,
                             P, F, H, W, f_len, h, w, dtype, stream);
}

extern "C" int ds_ring_scatter(void* pano, const void* in, int left, int top, int f0,
                               int P, int F, int H, int W, int f_len, int h, int w, int dtype,
                               void* stream) {
  return ds::ring_copy<false>(pano, const_cast<void*>(in), nullptr, 1, left, top, f0,
                              P, F, H, W, f_len, h, w, dtype, stream);
}

extern "C" int ds_ring_merge_plan(int32_t* plan_src, const int32_t* d_wins, int n,
                                  int H, int W, int h, int w, void* stream) {
  DS_REQUIRE(plan_src != nullptr && d_wins != nullptr, DS_ERR_NULL);
  DS_REQUIRE(n > 0 && H > 0 && W > 0 && h > 0 && w > 0, DS_ERR_DIM);
  DS_REQUIRE(h <= H && w <= W, DS_ERR_WINDOW);
  DS_REQUIRE(static_cast<long long>(n) * h * w < (1ll << 31), DS_ERR_DIM);
  const size_t smem = static_cast<size_t>(n) * 2 * sizeof(int32_t);
  DS_REQUIRE(smem <= 48 * 1024, DS_ERR_DIM);
  ds::ring_merge_plan_kernel<<<ds::div_up(static_cast<long long>(H) * W, ds::kBlock), ds::kBlock, smem,
                               static_cast<cudaStream_t>(stream)>>>(plan_src, d_wins, n, H, W, h, w);
  return ds::check_launch();
}
