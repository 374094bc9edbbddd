// Pixel-major resident panorama path ("channel/frame-contiguous latent").
//
// The batched step driver keeps the panorama of a denoising stage resident in HBM as
//     pm[pixel][frame][plane]      (H*W pixels, F frames, P = 4 planes; one "atom" = the 4
//                                   planes of one (pixel, frame) = 16 bytes in fp32)
// instead of the reference's [B, C, F, H, W] order (utils/shift_window_utils.py:42).  The
// index maps of the equirect<->perspective path scatter over PIXELS only, so in this order
// every tap of a view pixel is one contiguous F*P*4-byte row (256 B at F = 16): a warp-wide
// 16-byte load touches 4 cache lines instead of the 17-27 sectors per request of the
// plane-major kernels in sphere.cu (profiles/r01a_ncu_full_summary.md).
//
// Window tensors (gather output, UNet payloads) stay in the reference's [job][plane][frame]
// [h*w] order -- the UNet consumes / produces them -- so every kernel here is a gather or
// scatter fused with a [pixel][frame*plane] <-> [frame*plane][pixel] transpose through
// shared memory:
//   pm_gather_kernel         get_view_tensor_no_interpolate / _interpolate, all views of a step
//   pm_merge_nearest_kernel  set_view_tensor_no_interpolation, all views, last writer wins
//   pm_merge_bilinear_kernel set_view_tensor_bilinear, all views (weighted splat + normalise)
// and, with identity maps, the pack / unpack between the two layouts.
//
// Merges are in SCATTER form over a host-built, per-view compacted list of the view pixels
// (nearest) / panorama rows (bilinear) that survive the ordered sequence of writes, so no
// two threads ever write the same panorama atom: race-free without atomics, bit-identical
// to the reference's sequential order, and per-rank outputs have disjoint support when the
// job list is sharded.
#include "ds_common.cuh"

namespace ds {

constexpr int kPmPlanes = 4;     // planes per atom (B*C of the latent); other counts use sphere.cu
constexpr int kPmKT = 64;        // view pixels per gather CTA
constexpr int kPmET = 32;        // compacted entries (panorama pixels) per merge CTA
constexpr int kWarps = kBlock / 32;

template <typename T> struct Atom;
template <> struct Atom<float> {
  using type = float4;
  static __device__ __forceinline__ void unpack(const float4& a, float (&v)[4]) { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
  static __device__ __forceinline__ float4 pack(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Atom<__nv_bfloat16> {
  using type = uint2;
  static __device__ __forceinline__ void unpack(const uint2& a, float (&v)[4]) {
    v[0] = __uint_as_float(a.x << 16); v[1] = __uint_as_float(a.x & 0xffff0000u);
    v[2] = __uint_as_float(a.y << 16); v[3] = __uint_as_float(a.y & 0xffff0000u);
  }
  static __device__ __forceinline__ uint2 pack(const float (&v)[4]) {
    const unsigned a = __bfloat16_as_ushort(__float2bfloat16_rn(v[0])), b = __bfloat16_as_ushort(__float2bfloat16_rn(v[1]));
    const unsigned c = __bfloat16_as_ushort(__float2bfloat16_rn(v[2])), d = __bfloat16_as_ushort(__float2bfloat16_rn(v[3]));
    return make_uint2(a | (b << 16), c | (d << 16));
  }
};

// raw element <-> atom component without a float round trip (pure copies)
template <typename T> __device__ __forceinline__ void atom_split(const typename Atom<T>::type& a, T (&v)[4]);
template <> __device__ __forceinline__ void atom_split<float>(const float4& a, float (&v)[4]) { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
template <> __device__ __forceinline__ void atom_split<__nv_bfloat16>(const uint2& a, __nv_bfloat16 (&v)[4]) {
  v[0] = __ushort_as_bfloat16(static_cast<unsigned short>(a.x & 0xffffu)); v[1] = __ushort_as_bfloat16(static_cast<unsigned short>(a.x >> 16));
  v[2] = __ushort_as_bfloat16(static_cast<unsigned short>(a.y & 0xffffu)); v[3] = __ushort_as_bfloat16(static_cast<unsigned short>(a.y >> 16));
}
template <typename T> __device__ __forceinline__ typename Atom<T>::type atom_join(const T (&v)[4]);
template <> __device__ __forceinline__ float4 atom_join<float>(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ uint2 atom_join<__nv_bfloat16>(const __nv_bfloat16 (&v)[4]) {
  return make_uint2(__bfloat16_as_ushort(v[0]) | (static_cast<unsigned>(__bfloat16_as_ushort(v[1])) << 16),
                    __bfloat16_as_ushort(v[2]) | (static_cast<unsigned>(__bfloat16_as_ushort(v[3])) << 16));
}
template <typename T> __device__ __forceinline__ typename Atom<T>::type atom_ones();
template <> __device__ __forceinline__ float4 atom_ones<float>() { return make_float4(1.f, 1.f, 1.f, 1.f); }
template <> __device__ __forceinline__ uint2 atom_ones<__nv_bfloat16>() { return make_uint2(0x3f803f80u, 0x3f803f80u); }

template <typename A> __device__ __forceinline__ A ldg_atom(const A* p) { return __ldg(p); }

// ----------------------------------------------------------------------- gather ----
struct PmGatherArgs {
  int F, HW, f_len, hw, n_fblk;     // n_fblk = ceil(f_len / 8)
};

// CTA = (tile of kPmKT view pixels, job).  Phase 1: warp lanes = 4 pixels x 8 frames, each
// lane one 16-byte atom per tap (4 rows x 128 B per request).  Phase 2: the transposed
// tile [plane*f_len + fo][k] leaves shared memory as 16-byte vectors along k.
template <typename T, int TAPS>
__global__ void __launch_bounds__(kBlock)
pm_gather_kernel(const typename Atom<T>::type* __restrict__ pm, T* __restrict__ out,
                 const int32_t* __restrict__ d_idx, const float* __restrict__ d_wt,
                 const int32_t* __restrict__ d_jobs, int h_f0, PmGatherArgs a) {
  using A = typename Atom<T>::type;
  constexpr int VE = 16 / sizeof(T);            // elements per 16-byte vector
  constexpr int S = kPmKT + VE;                 // padded row stride (elements): 4 banks mod 32
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = kPmPlanes * a.f_len;
  T* tile = reinterpret_cast<T*>(smem_raw);                                   // [R][S]
  int32_t* s_idx = reinterpret_cast<int32_t*>(tile + static_cast<size_t>(R) * S);   // [TAPS][KT]
  float* s_wt = reinterpret_cast<float*>(s_idx + TAPS * kPmKT);               // [TAPS][KT] (bilinear)

  const int job = blockIdx.y;
  const int k0 = blockIdx.x * kPmKT;
  int map = job, f0 = h_f0;
  if (d_jobs != nullptr) {
    map = __ldg(d_jobs + 2 * job);
    f0 = __ldg(d_jobs + 2 * job + 1);
  }
  for (int t = threadIdx.x; t < TAPS * kPmKT; t += kBlock) {
    const int tap = t / kPmKT, kl = t - tap * kPmKT;
    const int k = k0 + kl;
    int idx = 0;
    float w = 0.f;
    if (k < a.hw) {
      if (d_idx != nullptr) idx = __ldg(d_idx + (static_cast<size_t>(map) * TAPS + tap) * a.hw + k);
      else idx = k;                                                  // identity map: unpack
      if (TAPS > 1) w = __ldg(d_wt + (static_cast<size_t>(map) * TAPS + tap) * a.hw + k);
    }
    s_idx[t] = idx;
    if (TAPS > 1) s_wt[t] = w;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kq = lane >> 3, fl = lane & 7;
  const int G = (kPmKT / 4) * a.n_fblk;
  constexpr int U = TAPS > 1 ? 2 : 4;           // groups in flight per warp
  for (int g0 = warp * U; g0 < G; g0 += kWarps * U) {
    A v[U][TAPS];
    int kl[U], fo[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int g = g0 + u;
      const int k4 = g / a.n_fblk, fb = g - k4 * a.n_fblk;
      kl[u] = k4 * 4 + kq;
      fo[u] = fb * 8 + fl;
      ok[u] = g < G && fo[u] < a.f_len && k0 + kl[u] < a.hw;
      if (ok[u]) {
        int f = f0 + fo[u];
        f = f % a.F;
#pragma unroll
        for (int t = 0; t < TAPS; ++t)
          v[u][t] = ldg_atom(pm + static_cast<size_t>(s_idx[t * kPmKT + kl[u]]) * a.F + f);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
      T* dstp = tile + static_cast<size_t>(fo[u]) * S + kl[u];
      if (TAPS == 1) {
        T e[4];
        atom_split<T>(v[u][0], e);
#pragma unroll
        for (int p = 0; p < 4; ++p) dstp[static_cast<size_t>(p) * a.f_len * S] = e[p];
      } else {
        // ATen grid_sample (CPU) order: nw*w + ne*w + sw*w + se*w contracted to an FMA chain
        float r[4], x[4];
        Atom<T>::unpack(v[u][0], x);
        const float w0 = s_wt[kl[u]];
#pragma unroll
        for (int p = 0; p < 4; ++p) r[p] = __fmul_rn(x[p], w0);
#pragma unroll
        for (int t = 1; t < TAPS; ++t) {
          Atom<T>::unpack(v[u][t], x);
          const float wt = s_wt[t * kPmKT + kl[u]];
#pragma unroll
          for (int p = 0; p < 4; ++p) r[p] = __fmaf_rn(x[p], wt, r[p]);
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) dstp[static_cast<size_t>(p) * a.f_len * S] = Elem<T>::from_f(r[p]);
      }
    }
  }
  __syncthreads();

  constexpr int VPR = kPmKT / VE;                // 16-byte vectors per tile row
  T* __restrict__ o = out + static_cast<size_t>(job) * R * a.hw + k0;
  for (int u = threadIdx.x; u < R * VPR; u += kBlock) {
    const int row = u / VPR, c = u - row * VPR;
    if (k0 + c * VE < a.hw)
      *reinterpret_cast<uint4*>(o + static_cast<size_t>(row) * a.hw + c * VE) =
          *reinterpret_cast<const uint4*>(tile + static_cast<size_t>(row) * S + c * VE);
  }
}

// ------------------------------------------------------------------------ merges ----
struct PmMergeArgs {
  void* pm[DS_MAX_PAYLOADS];
  const void* pay[DS_MAX_PAYLOADS];
  void* mask;
  int mask_mode;
  int Q, n_tiles, n_ent;
  int vpf, job_begin, job_count;
  int F, HW, f_len, hw;
};

// dynamic indexing of a kernel-parameter array would force a local-memory copy of the struct
template <typename PT> __device__ __forceinline__ PT pm_sel(PT const (&arr)[DS_MAX_PAYLOADS], int q) {
  PT r = arr[0];
#pragma unroll
  for (int i = 1; i < DS_MAX_PAYLOADS; ++i) r = q == i ? arr[i] : r;
  return r;
}

struct PmTileCtx {
  int v, begin, count, s, job, fb;
};

// Resolve this CTA's (tile, frame window) to its payload job; false => nothing to do.
__device__ __forceinline__ bool pm_tile_ctx(const PmMergeArgs& a, const int32_t* __restrict__ d_tiles,
                                            const int32_t* __restrict__ d_fbegin, PmTileCtx& c) {
  const int t = blockIdx.x;
  if (d_tiles != nullptr) {
    c.v = __ldg(d_tiles + 3 * t);
    c.begin = __ldg(d_tiles + 3 * t + 1);
    c.count = __ldg(d_tiles + 3 * t + 2);
  } else {                                        // implicit identity tiling (pack)
    c.v = 0;
    c.begin = t * kPmET;
    c.count = min(kPmET, a.n_ent - c.begin);
  }
  c.s = blockIdx.y;
  c.job = c.s * a.vpf + c.v - a.job_begin;
  c.fb = d_fbegin != nullptr ? __ldg(d_fbegin + c.s) : 0;
  return c.count > 0 && c.job >= 0 && c.job < a.job_count;
}

// Window frame fo of frame window s lands on panorama frame f; it is written only if s is the
// LAST frame window (in job order) containing f (d_ftab, maps.frame_table).
__device__ __forceinline__ bool pm_frame_live(const PmMergeArgs& a, const int32_t* __restrict__ d_ftab,
                                              const PmTileCtx& c, int fo, int& f) {
  f = (c.fb + fo) % a.F;
  return d_ftab == nullptr || __ldg(d_ftab + f) == c.s * a.f_len + fo;
}

// CTA = (tile of <= 32 surviving view pixels of one view, frame window).
// Load: warp <-> (payload q, window frame fo), lane <-> entry: 4 plane loads -> one atom ->
// shared tile[q][entry][fo].  Store: 16 lanes write one full panorama row (f_len atoms).
template <typename T>
__global__ void __launch_bounds__(kBlock)
pm_merge_nearest_kernel(PmMergeArgs a, const int32_t* __restrict__ d_tiles,
                        const int32_t* __restrict__ d_ent_k, const int32_t* __restrict__ d_ent_dst,
                        const int32_t* __restrict__ d_ftab, const int32_t* __restrict__ d_fbegin) {
  using A = typename Atom<T>::type;
  PmTileCtx c;
  if (!pm_tile_ctx(a, d_tiles, d_fbegin, c)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int SA = a.f_len | 1;                                        // odd row stride in atoms
  A* tile = reinterpret_cast<A*>(smem_raw);                          // [Q][ET][SA]
  int32_t* s_dst = reinterpret_cast<int32_t*>(tile + static_cast<size_t>(a.Q) * kPmET * SA);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool have = lane < c.count;
  int k = 0;
  if (have) k = d_ent_k != nullptr ? __ldg(d_ent_k + c.begin + lane) : c.begin + lane;
  if (warp == 0 && have) {
    const int dst = d_ent_dst != nullptr ? __ldg(d_ent_dst + c.begin + lane) : c.begin + lane;
    s_dst[lane] = dst;
    if (a.mask_mode == DS_MASK_PLANE) static_cast<T*>(a.mask)[dst] = Elem<T>::one();
  }
  const size_t plane_stride = static_cast<size_t>(a.f_len) * a.hw;
  const int n_pairs = a.Q * a.f_len;
  constexpr int U = 4;
  for (int p0 = warp * U; p0 < n_pairs; p0 += kWarps * U) {
    T e[U][4];
    bool live[U];
    int q[U], fo[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pr = p0 + u;
      q[u] = pr / a.f_len;
      fo[u] = pr - q[u] * a.f_len;
      int f;
      live[u] = pr < n_pairs && have && pm_frame_live(a, d_ftab, c, fo[u], f);
      if (live[u]) {
        const T* __restrict__ src = static_cast<const T*>(pm_sel(a.pay, q[u])) +
                                    (static_cast<size_t>(c.job) * kPmPlanes * a.f_len + fo[u]) * a.hw + k;
#pragma unroll
        for (int p = 0; p < 4; ++p) e[u][p] = ldg(src + p * plane_stride);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (live[u]) tile[(static_cast<size_t>(q[u]) * kPmET + lane) * SA + fo[u]] = atom_join<T>(e[u]);
  }
  __syncthreads();

  const int per_q = kPmET * a.f_len;
  for (int u = threadIdx.x; u < a.Q * per_q; u += kBlock) {
    const int q = u / per_q, r = u - q * per_q;
    const int i = r / a.f_len, fo = r - i * a.f_len;
    int f;
    if (i < c.count && pm_frame_live(a, d_ftab, c, fo, f)) {
      const size_t off = static_cast<size_t>(s_dst[i]) * a.F + f;
      static_cast<A*>(pm_sel(a.pm, q))[off] = tile[(static_cast<size_t>(q) * kPmET + i) * SA + fo];
      if (q == 0 && a.mask_mode == DS_MASK_FULL) static_cast<A*>(a.mask)[off] = atom_ones<T>();
    }
  }
}

// CTA = (tile of <= 32 panorama rows owned by one view, frame window).  Each lane owns a row
// and walks its CSR entries (tap-major, k ascending -- index_add_'s order) accumulating
//   acc = acc + (payload * w);  ws = ws + w;   row = acc / ws     (each op rounded on its own)
template <typename T>
__global__ void __launch_bounds__(kBlock)
pm_merge_bilinear_kernel(PmMergeArgs a, const int32_t* __restrict__ d_tiles,
                         const int32_t* __restrict__ d_row_dst, const int32_t* __restrict__ d_row_ptr,
                         const int32_t* __restrict__ d_ent_k, const float* __restrict__ d_ent_w,
                         const int32_t* __restrict__ d_ftab, const int32_t* __restrict__ d_fbegin) {
  using A = typename Atom<T>::type;
  PmTileCtx c;
  if (!pm_tile_ctx(a, d_tiles, d_fbegin, c)) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int SA = a.f_len | 1;
  A* tile = reinterpret_cast<A*>(smem_raw);                          // [Q][ET][SA]
  int32_t* s_dst = reinterpret_cast<int32_t*>(tile + static_cast<size_t>(a.Q) * kPmET * SA);   // -1: row not written

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool have = lane < c.count;
  int e0 = 0, e1 = 0;
  if (have) {
    e0 = __ldg(d_row_ptr + c.begin + lane);
    e1 = __ldg(d_row_ptr + c.begin + lane + 1);
  }
  const size_t plane_stride = static_cast<size_t>(a.f_len) * a.hw;
  const int n_pairs = a.Q * a.f_len;
  constexpr int U = 2;
  float ws = 0.f;
  for (int p0 = warp * U; p0 < n_pairs; p0 += kWarps * U) {
    float acc[U][4];
    const T* src[U];
    bool live[U];
    int q[U], fo[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int pr = p0 + u;
      q[u] = pr / a.f_len;
      fo[u] = pr - q[u] * a.f_len;
      int f;
      live[u] = pr < n_pairs && have && pm_frame_live(a, d_ftab, c, fo[u], f);
      src[u] = live[u] ? static_cast<const T*>(pm_sel(a.pay, q[u])) +
                             (static_cast<size_t>(c.job) * kPmPlanes * a.f_len + fo[u]) * a.hw
                       : nullptr;
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[u][p] = 0.f;
    }
    ws = 0.f;
    for (int e = e0; e < e1; ++e) {
      const int k = __ldg(d_ent_k + e);
      const float w = __ldg(d_ent_w + e);
      ws = __fadd_rn(ws, w);
      float x[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (live[u]) {
#pragma unroll
          for (int p = 0; p < 4; ++p) x[u][p] = Elem<T>::to_f(ldg(src[u] + p * plane_stride + k));
        }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (live[u]) {
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[u][p] = __fadd_rn(acc[u][p], __fmul_rn(x[u][p], w));
        }
    }
    if (ws > 0.f) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (live[u]) {
          float r[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) r[p] = __fdiv_rn(acc[u][p], ws);
          tile[(static_cast<size_t>(q[u]) * kPmET + lane) * SA + fo[u]] = Atom<T>::pack(r);
        }
    }
  }
  if (warp == 0) {
    // every warp walks the same entries, so warp 0's ws is the row's weight sum
    if (n_pairs <= 0) {                         // mask-only call: weight sum still needed
      for (int e = e0; e < e1; ++e) ws = __fadd_rn(ws, __ldg(d_ent_w + e));
    }
    const bool wr = have && ws > 0.f;
    const int dst = wr ? __ldg(d_row_dst + c.begin + lane) : -1;
    s_dst[lane] = dst;
    if (wr && a.mask_mode == DS_MASK_PLANE) static_cast<T*>(a.mask)[dst] = Elem<T>::one();
  }
  __syncthreads();

  const int per_q = kPmET * a.f_len;
  const int n_units = (a.Q > 0 ? a.Q : 1) * per_q;
  for (int u = threadIdx.x; u < n_units; u += kBlock) {
    const int q = u / per_q, r = u - q * per_q;
    const int i = r / a.f_len, fo = r - i * a.f_len;
    int f;
    if (i < c.count && s_dst[i] >= 0 && pm_frame_live(a, d_ftab, c, fo, f)) {
      const size_t off = static_cast<size_t>(s_dst[i]) * a.F + f;
      if (a.Q > 0) static_cast<A*>(pm_sel(a.pm, q))[off] = tile[(static_cast<size_t>(q) * kPmET + i) * SA + fo];
      if (q == 0 && a.mask_mode == DS_MASK_FULL) static_cast<A*>(a.mask)[off] = atom_ones<T>();
    }
  }
}

static bool pm_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int pm_fill_merge_args(PmMergeArgs& a, void* const* h_pms, const void* const* h_payloads, int Q,
                              void* mask, int mask_mode, int n_tiles, int n_ent, const int32_t* d_ftab,
                              const int32_t* d_fbegin, int n_fwin, int vpf, int job_begin, int job_count,
                              int P, int F, int HW, int f_len, int hw) {
  DS_REQUIRE(Q >= 0 && Q <= DS_MAX_PAYLOADS, DS_ERR_COUNT);
  DS_REQUIRE(Q == 0 || (h_pms != nullptr && h_payloads != nullptr), DS_ERR_NULL);
  DS_REQUIRE(P == kPmPlanes, DS_ERR_DIM);
  DS_REQUIRE(n_tiles >= 0 && n_ent >= 0 && n_fwin > 0 && vpf > 0 && F > 0 && HW > 0 && f_len > 0 && hw > 0, DS_ERR_DIM);
  DS_REQUIRE(f_len <= F && n_fwin <= 65535, DS_ERR_DIM);
  DS_REQUIRE(mask_mode == DS_MASK_NONE || mask_mode == DS_MASK_PLANE || mask_mode == DS_MASK_FULL, DS_ERR_DIM);
  DS_REQUIRE(mask_mode == DS_MASK_NONE || mask != nullptr, DS_ERR_NULL);
  DS_REQUIRE(Q > 0 || mask_mode != DS_MASK_NONE, DS_ERR_COUNT);
  DS_REQUIRE(job_begin >= 0 && job_count > 0, DS_ERR_DIM);
  // frame ring: both tables or neither; without them the single frame window is the identity
  DS_REQUIRE((d_ftab == nullptr) == (d_fbegin == nullptr), DS_ERR_NULL);
  DS_REQUIRE(d_ftab != nullptr || (f_len == F && n_fwin == 1), DS_ERR_DIM);
  DS_REQUIRE(mask_mode != DS_MASK_PLANE || d_ftab == nullptr, DS_ERR_DIM);
  for (int q = 0; q < DS_MAX_PAYLOADS; ++q) {
    a.pm[q] = q < Q ? h_pms[q] : nullptr;
    a.pay[q] = q < Q ? h_payloads[q] : nullptr;
    if (q < Q) DS_REQUIRE(a.pm[q] != nullptr && a.pay[q] != nullptr, DS_ERR_NULL);
    if (q < Q) DS_REQUIRE(pm_aligned16(a.pm[q]), DS_ERR_ALIGN);
  }
  if (mask_mode == DS_MASK_FULL) DS_REQUIRE(pm_aligned16(mask), DS_ERR_ALIGN);
  a.mask = mask; a.mask_mode = mask_mode;
  a.Q = Q; a.n_tiles = n_tiles; a.n_ent = n_ent;
  a.vpf = vpf; a.job_begin = job_begin; a.job_count = job_count;
  a.F = F; a.HW = HW; a.f_len = f_len; a.hw = hw;
  return DS_OK;
}

template <typename T>
static size_t pm_merge_smem(int Q, int f_len) {
  return static_cast<size_t>(Q > 0 ? Q : 1) * kPmET * (f_len | 1) * sizeof(typename Atom<T>::type) + kPmET * sizeof(int32_t);
}

}  // namespace ds

extern "C" int ds_pm_gather(const void* pm, void* out, const int32_t* d_idx, const float* d_wt, int taps,
                            const int32_t* d_jobs, int h_f0, int n_jobs, int P, int F, int HW, int f_len,
                            int hw, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(pm != nullptr && out != nullptr, DS_ERR_NULL);
  DS_REQUIRE(taps == 1 || taps == 4, DS_ERR_DIM);
  DS_REQUIRE(taps == 1 || (d_idx != nullptr && d_wt != nullptr), DS_ERR_NULL);
  DS_REQUIRE(P == kPmPlanes, DS_ERR_DIM);
  DS_REQUIRE(n_jobs > 0 && n_jobs <= 65535 && F > 0 && HW > 0 && f_len > 0 && hw > 0 && h_f0 >= 0, DS_ERR_DIM);
  DS_REQUIRE(d_idx != nullptr || hw <= HW, DS_ERR_DIM);
  DS_REQUIRE(pm_aligned16(pm) && pm_aligned16(out), DS_ERR_ALIGN);
  PmGatherArgs a{F, HW, f_len, hw, (f_len + 7) / 8};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(div_up(hw, kPmKT), n_jobs);
  DS_DISPATCH_DTYPE(dtype, {
    constexpr int VE = 16 / sizeof(ET);
    DS_REQUIRE(hw % VE == 0, DS_ERR_ALIGN);
    const size_t smem = static_cast<size_t>(kPmPlanes) * f_len * (kPmKT + VE) * sizeof(ET) + static_cast<size_t>(taps) * kPmKT * 8;
    DS_REQUIRE(smem <= 200 * 1024, DS_ERR_DIM);
    using AT = typename Atom<ET>::type;
    if (taps == 1) {
      if (smem > 48 * 1024) cudaFuncSetAttribute(pm_gather_kernel<ET, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      pm_gather_kernel<ET, 1><<<grid, kBlock, smem, st>>>(static_cast<const AT*>(pm), static_cast<ET*>(out), d_idx, d_wt, d_jobs, h_f0, a);
    } else {
      if (smem > 48 * 1024) cudaFuncSetAttribute(pm_gather_kernel<ET, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      pm_gather_kernel<ET, 4><<<grid, kBlock, smem, st>>>(static_cast<const AT*>(pm), static_cast<ET*>(out), d_idx, d_wt, d_jobs, h_f0, a);
    }
  });
  return check_launch();
}

extern "C" int ds_pm_merge_nearest(void* const* h_pms, const void* const* h_payloads, int Q,
                                   void* mask, int mask_mode, const int32_t* d_tiles, int n_tiles,
                                   const int32_t* d_ent_k, const int32_t* d_ent_dst, int n_ent,
                                   const int32_t* d_ftab, const int32_t* d_fbegin, int n_fwin,
                                   int views_per_fwin, int job_begin, int job_count,
                                   int P, int F, int HW, int f_len, int hw, int dtype, void* stream) {
  using namespace ds;
  PmMergeArgs a;
  int rc = pm_fill_merge_args(a, h_pms, h_payloads, Q, mask, mask_mode, n_tiles, n_ent, d_ftab, d_fbegin, n_fwin,
                              views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw);
  if (rc != DS_OK) return rc;
  DS_REQUIRE((d_tiles == nullptr) == (d_ent_k == nullptr) && (d_tiles == nullptr) == (d_ent_dst == nullptr), DS_ERR_NULL);
  if (d_tiles == nullptr) {                      // identity plan: entry i = (k = i, dst = i)
    DS_REQUIRE(n_ent <= HW && n_ent <= hw && views_per_fwin == 1, DS_ERR_DIM);
    n_tiles = static_cast<int>(div_up(n_ent, kPmET));
    a.n_tiles = n_tiles;
  }
  if (n_tiles == 0) return DS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(n_tiles, n_fwin);
  DS_DISPATCH_DTYPE(dtype, {
    const size_t smem = pm_merge_smem<ET>(Q, f_len);
    DS_REQUIRE(smem <= 200 * 1024, DS_ERR_DIM);
    if (smem > 48 * 1024) cudaFuncSetAttribute(pm_merge_nearest_kernel<ET>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    pm_merge_nearest_kernel<ET><<<grid, kBlock, smem, st>>>(a, d_tiles, d_ent_k, d_ent_dst, d_ftab, d_fbegin);
  });
  return check_launch();
}

extern "C" int ds_pm_merge_bilinear(void* const* h_pms, const void* const* h_payloads, int Q,
                                    void* mask, int mask_mode, const int32_t* d_tiles, int n_tiles,
                                    const int32_t* d_row_dst, const int32_t* d_row_ptr, int n_rows,
                                    const int32_t* d_ent_k, const float* d_ent_w,
                                    const int32_t* d_ftab, const int32_t* d_fbegin, int n_fwin,
                                    int views_per_fwin, int job_begin, int job_count,
                                    int P, int F, int HW, int f_len, int hw, int dtype, void* stream) {
  using namespace ds;
  PmMergeArgs a;
  int rc = pm_fill_merge_args(a, h_pms, h_payloads, Q, mask, mask_mode, n_tiles, n_rows, d_ftab, d_fbegin, n_fwin,
                              views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw);
  if (rc != DS_OK) return rc;
  DS_REQUIRE(d_tiles != nullptr && d_row_dst != nullptr && d_row_ptr != nullptr && d_ent_k != nullptr && d_ent_w != nullptr, DS_ERR_NULL);
  if (n_tiles == 0) return DS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(n_tiles, n_fwin);
  DS_DISPATCH_DTYPE(dtype, {
    const size_t smem = pm_merge_smem<ET>(Q, f_len);
    DS_REQUIRE(smem <= 200 * 1024, DS_ERR_DIM);
    if (smem > 48 * 1024) cudaFuncSetAttribute(pm_merge_bilinear_kernel<ET>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    pm_merge_bilinear_kernel<ET><<<grid, kBlock, smem, st>>>(a, d_tiles, d_row_dst, d_row_ptr, d_ent_k, d_ent_w, d_ftab, d_fbegin);
  });
  return check_launch();
}
