// Index-map driven window extraction and merge: the CUDA side of the equirectangular <->
// perspective path (reference utils/panorama_tensor_utils.py, utils/ring_panorama_tensor_utils.py)
// and, with plans from ds_ring_merge_plan, of the batched planar-ring merge.
//
// The transcendental uv map is NOT evaluated here: floor(u), floor(v) flip for thousands
// of pixels between libm implementations, so the host builds the int32 maps
// with the reference's own op sequence and these kernels only consume them.
//
// All four kernels are HBM/L2-bound gathers with one fully coalesced side:
//   gather_*  : scattered (but spatially coherent) panorama reads -> coalesced window writes
//   merge_*   : coherent payload reads -> panorama writes in ascending pixel order.
// The merge is written in GATHER FORM (one thread owns one panorama pixel and pulls the
// winning view pixel(s)), so it needs no atomics, is race-free and reproduces the
// reference's sequential "last writer wins" / index_add_ order bit for bit.
#include <type_traits>

#include "ds_common.cuh"

namespace ds {

struct GatherDims {
  int P, F, HW, f_len, hw;
};

// ---------------------------------------------------------------- nearest gather ----
// thread = VPT consecutive view pixels of one job, CPT (plane, frame) slices.
template <typename T, int VPT, int CPT>
__global__ void __launch_bounds__(kBlock)
gather_nearest_kernel(const T* __restrict__ pano, T* __restrict__ out,
                      const int32_t* __restrict__ d_idx, const int32_t* __restrict__ d_jobs,
                      int h_f0, GatherDims d) {
  const int k0 = (blockIdx.x * kBlock + threadIdx.x) * VPT;
  if (k0 >= d.hw) return;
  const int job = blockIdx.y;
  int map = job, f0 = h_f0;
  if (d_jobs != nullptr) {
    map = __ldg(d_jobs + 2 * job);
    f0 = __ldg(d_jobs + 2 * job + 1);
  }
  int idx[VPT];
  if (VPT == 4) {
    const int4 t = __ldg(reinterpret_cast<const int4*>(d_idx + static_cast<size_t>(map) * d.hw + k0));
    idx[0] = t.x; idx[1 % VPT] = t.y; idx[2 % VPT] = t.z; idx[3 % VPT] = t.w;
  } else {
#pragma unroll
    for (int e = 0; e < VPT; ++e) idx[e] = __ldg(d_idx + static_cast<size_t>(map) * d.hw + k0 + e);
  }
  const int CF = d.P * d.f_len;
  const int cf0 = blockIdx.z * CPT;
  int pl = cf0 / d.f_len, fo = cf0 - pl * d.f_len;
  T v[CPT][VPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (cf0 + c < CF) {
      const int f = (f0 + fo) % d.F;
      const T* __restrict__ plane = pano + (static_cast<size_t>(pl) * d.F + f) * d.HW;
#pragma unroll
      for (int e = 0; e < VPT; ++e) v[c][e] = ldg(plane + idx[e]);
    }
    if (++fo == d.f_len) { fo = 0; ++pl; }
  }
  T* __restrict__ o = out + (static_cast<size_t>(job) * CF + cf0) * d.hw + k0;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (cf0 + c < CF) {
      if (VPT == 4) {
        *reinterpret_cast<typename Elem<T>::vec4*>(o + static_cast<size_t>(c) * d.hw) =
            *reinterpret_cast<const typename Elem<T>::vec4*>(&v[c][0]);
      } else {
#pragma unroll
        for (int e = 0; e < VPT; ++e) o[static_cast<size_t>(c) * d.hw + e] = v[c][e];
      }
    }
  }
}

// --------------------------------------------------------------- bilinear gather ----
// thread = one view pixel of one job, CPT (plane, frame) slices; 4*CPT loads in flight.
template <typename T, int CPT>
__global__ void __launch_bounds__(kBlock)
gather_bilinear_kernel(const T* __restrict__ pano, T* __restrict__ out,
                       const int32_t* __restrict__ d_idx4, const float* __restrict__ d_wt4,
                       const int32_t* __restrict__ d_jobs, int h_f0, GatherDims d) {
  const int k = blockIdx.x * kBlock + threadIdx.x;
  if (k >= d.hw) return;
  const int job = blockIdx.y;
  int map = job, f0 = h_f0;
  if (d_jobs != nullptr) {
    map = __ldg(d_jobs + 2 * job);
    f0 = __ldg(d_jobs + 2 * job + 1);
  }
  int idx[4];
  float wt[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    idx[t] = __ldg(d_idx4 + (static_cast<size_t>(map) * 4 + t) * d.hw + k);
    wt[t] = __ldg(d_wt4 + (static_cast<size_t>(map) * 4 + t) * d.hw + k);
  }
  const int CF = d.P * d.f_len;
  const int cf0 = blockIdx.z * CPT;
  int pl = cf0 / d.f_len, fo = cf0 - pl * d.f_len;
  float v[CPT][4];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (cf0 + c < CF) {
      const int f = (f0 + fo) % d.F;
      const T* __restrict__ plane = pano + (static_cast<size_t>(pl) * d.F + f) * d.HW;
#pragma unroll
      for (int t = 0; t < 4; ++t) v[c][t] = Elem<T>::to_f(ldg(plane + idx[t]));
    }
    if (++fo == d.f_len) { fo = 0; ++pl; }
  }
  T* __restrict__ o = out + (static_cast<size_t>(job) * CF + cf0) * d.hw + k;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (cf0 + c < CF) {
      // ATen grid_sample (CPU) order: nw*w + ne*w + sw*w + se*w contracted to an FMA chain.
      float r = __fmul_rn(v[c][0], wt[0]);
      r = __fmaf_rn(v[c][1], wt[1], r);
      r = __fmaf_rn(v[c][2], wt[2], r);
      r = __fmaf_rn(v[c][3], wt[3], r);
      o[static_cast<size_t>(c) * d.hw] = Elem<T>::from_f(r);
    }
  }
}

// ------------------------------------------------------------------------ merges ----
struct MergeArgs {
  void* pano[DS_MAX_PAYLOADS];
  const void* pay[DS_MAX_PAYLOADS];
  void* mask;
  int mask_mode, mask_planes;
  int Q, T;
  int vpf;                  // views per frame window
  int job_begin, job_count; // sharding: only jobs in [job_begin, job_begin + job_count) are applied
  int P, F, HW, f_len, hw;
  const ds_peer_table_t* peer;   // exchange mode: every write goes to the replicas listed there (pano / mask above unused)
};

// One element (or 16-byte vector) to the local panorama, or to every rank's replica over NVLink.
template <typename V, typename T, bool PEER>
__device__ __forceinline__ void merge_store(const MergeArgs& a, int q, size_t off, const V& v) {
  if (!PEER) {
    *reinterpret_cast<V*>(static_cast<T*>(a.pano[q]) + off) = v;
  } else {
    for (int r = 0; r < a.peer->world; ++r) *reinterpret_cast<V*>(reinterpret_cast<T*>(a.peer->pm[r][q]) + off) = v;
  }
}
template <typename T, bool PEER>
__device__ __forceinline__ void merge_store_mask(const MergeArgs& a, size_t off) {
  if (!PEER) {
    static_cast<T*>(a.mask)[off] = Elem<T>::one();
  } else {
    for (int r = 0; r < a.peer->world; ++r) reinterpret_cast<T*>(a.peer->mask[r])[off] = Elem<T>::one();
  }
}

// Resolve panorama slice (pl, f) + view v to the local payload job and window frame.
// Returns false when the frame is not written this step or the job belongs to another rank.
__device__ __forceinline__ bool merge_source(const MergeArgs& a, const int32_t* __restrict__ d_ftab,
                                             int f, int v, int& job, int& fo) {
  const int sfo = d_ftab != nullptr ? __ldg(d_ftab + f) : f;
  if (sfo < 0) return false;
  const int s = sfo / a.f_len;
  fo = sfo - s * a.f_len;
  job = s * a.vpf + v - a.job_begin;
  return job >= 0 && job < a.job_count;
}

// thread = one plan entry (VEC consecutive panorama pixels), CPT panorama (plane, frame)
// slices, all Q payloads.
template <typename T, int VEC, int CPT, int QMAX, bool PEER = false>
__global__ void __launch_bounds__(kBlock)
merge_nearest_kernel(MergeArgs a, const int32_t* __restrict__ d_dst,
                     const int32_t* __restrict__ d_src, const int32_t* __restrict__ d_ftab) {
  using V = typename std::conditional<VEC == 4, typename Elem<T>::vec4, T>::type;
  const int t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= a.T) return;
  const int dst = d_dst != nullptr ? __ldg(d_dst + t) : t * VEC;
  const int src = __ldg(d_src + (d_dst != nullptr ? t : t * VEC));
  if (src < 0) return;
  const int v = src / a.hw, k = src - v * a.hw;
  const int PF = a.P * a.F;
  const int pf0 = blockIdx.y * CPT;
  int pl = pf0 / a.F, f = pf0 - pl * a.F;
  V val[QMAX][CPT];
  bool live[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    int job, fo;
    live[c] = pf0 + c < PF && merge_source(a, d_ftab, f, v, job, fo);
    if (live[c]) {
      const size_t off = ((static_cast<size_t>(job) * a.P + pl) * a.f_len + fo) * a.hw + k;
#pragma unroll
      for (int q = 0; q < QMAX; ++q)
        if (q < a.Q) val[q][c] = __ldg(reinterpret_cast<const V*>(static_cast<const T*>(a.pay[q]) + off));
    }
    if (++f == a.F) { f = 0; ++pl; }
  }
  pl = pf0 / a.F; f = pf0 - pl * a.F;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (live[c]) {
      const size_t off = (static_cast<size_t>(pl) * a.F + f) * a.HW + dst;
#pragma unroll
      for (int q = 0; q < QMAX; ++q)
        if (q < a.Q) merge_store<V, T, PEER>(a, q, off, val[q][c]);
      if (a.mask_mode == DS_MASK_FULL && pl < a.mask_planes) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) merge_store_mask<T, PEER>(a, off + e);
      }
    }
    if (++f == a.F) { f = 0; ++pl; }
  }
  // DS_MASK_PLANE (host guarantees d_ftab == nullptr): one write per pixel, from slice 0
  if (a.mask_mode == DS_MASK_PLANE && blockIdx.y == 0 && live[0]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) merge_store_mask<T, PEER>(a, dst + e);
  }
}

// The nearest merge fused with the denoise epilogue, reference-order tensors (the planar ring loop,
// t2v_normal_pipeline.py:526-557, and any nearest plan): pay[0] = x (the windows as gathered),
// pay[1] = eps_cond, pay[2] = eps_uncond or NULL, pay[3] = ddim noise or NULL; pano[0] receives
// x_prev and pano[1] pred_x0 of ddim_step, computed for the winning elements only -- elements
// that lose to a later window are never denoised and x_prev / pred_x0 never exist as tensors.
// Same thread mapping as merge_nearest_kernel.
// merge_prev_denoised_ratio mix between ddim_step and the latent set (i2v_sphere_panorama_pipeline.py:938-943):
//   x_prev <- mix_latents_with_mask(x_prev, prev, m, r);  prev = the window as gathered (:866), m = the
// denoised-mask window read at gather time (:868-873); both in window layout like x.  prev NULL: off.
struct PrevMix {
  const void* prev;
  const void* mview;
  float r, omr;
};

template <typename T, int VEC, int CPT, bool PEER, bool PV>
__global__ void __launch_bounds__(kBlock)
merge_nearest_ddim_kernel(MergeArgs a, DdimCoef dc, PrevMix pv, const int32_t* __restrict__ d_dst,
                          const int32_t* __restrict__ d_src, const int32_t* __restrict__ d_ftab) {
  using V = typename std::conditional<VEC == 4, typename Elem<T>::vec4, T>::type;
  const int t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= a.T) return;
  const int dst = d_dst != nullptr ? __ldg(d_dst + t) : t * VEC;
  const int src = __ldg(d_src + (d_dst != nullptr ? t : t * VEC));
  if (src < 0) return;
  const int v = src / a.hw, k = src - v * a.hw;
  const bool has_eu = a.pay[2] != nullptr, has_nz = a.pay[3] != nullptr;
  const int PF = a.P * a.F;
  const int pf0 = blockIdx.y * CPT;
  int pl = pf0 / a.F, f = pf0 - pl * a.F;
  constexpr bool has_pv = PV;
  V val[PV ? 6 : 4][CPT];
  bool live[CPT];
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    int job, fo;
    live[c] = pf0 + c < PF && merge_source(a, d_ftab, f, v, job, fo);
    if (live[c]) {
      const size_t off = ((static_cast<size_t>(job) * a.P + pl) * a.f_len + fo) * a.hw + k;
      val[0][c] = __ldg(reinterpret_cast<const V*>(static_cast<const T*>(a.pay[0]) + off));
      val[1][c] = __ldg(reinterpret_cast<const V*>(static_cast<const T*>(a.pay[1]) + off));
      if (has_eu) val[2][c] = __ldg(reinterpret_cast<const V*>(static_cast<const T*>(a.pay[2]) + off));
      if (has_nz) val[3][c] = __ldg(reinterpret_cast<const V*>(static_cast<const T*>(a.pay[3]) + off));
      if constexpr (PV) {
        val[4][c] = __ldg(reinterpret_cast<const V*>(static_cast<const T*>(pv.prev) + off));
        val[5][c] = __ldg(reinterpret_cast<const V*>(static_cast<const T*>(pv.mview) + off));
      }
    }
    if (++f == a.F) { f = 0; ++pl; }
  }
  pl = pf0 / a.F; f = pf0 - pl * a.F;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (live[c]) {
      V xp, x0;
      const T* ex = reinterpret_cast<const T*>(&val[0][c]);
      const T* ec = reinterpret_cast<const T*>(&val[1][c]);
      const T* eu = reinterpret_cast<const T*>(&val[2][c]);
      const T* nz = reinterpret_cast<const T*>(&val[3][c]);
      const T* pr = reinterpret_cast<const T*>(&val[PV ? 4 : 0][c]);
      const T* mv = reinterpret_cast<const T*>(&val[PV ? 5 : 0][c]);
      T* oxp = reinterpret_cast<T*>(&xp);
      T* ox0 = reinterpret_cast<T*>(&x0);
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        float fxp, fx0;
        ddim_point(dc, Elem<T>::to_f(ex[e]), Elem<T>::to_f(ec[e]), has_eu ? Elem<T>::to_f(eu[e]) : 0.f,
                   has_nz ? Elem<T>::to_f(nz[e]) : 0.f, has_eu, has_nz, fxp, fx0);
        if constexpr (PV)                      // the reference mixes the ROUNDED x_prev tensor
          fxp = mix_point<T>(Elem<T>::to_f(Elem<T>::from_f(fxp)), Elem<T>::to_f(pr[e]), Elem<T>::to_f(mv[e]), pv.r, pv.omr);
        oxp[e] = Elem<T>::from_f(fxp);
        ox0[e] = Elem<T>::from_f(fx0);
      }
      const size_t off = (static_cast<size_t>(pl) * a.F + f) * a.HW + dst;
      merge_store<V, T, PEER>(a, 0, off, xp);
      merge_store<V, T, PEER>(a, 1, off, x0);
      if (a.mask_mode == DS_MASK_FULL && pl < a.mask_planes) {
#pragma unroll
        for (int e = 0; e < VEC; ++e) merge_store_mask<T, PEER>(a, off + e);
      }
    }
    if (++f == a.F) { f = 0; ++pl; }
  }
  if (a.mask_mode == DS_MASK_PLANE && blockIdx.y == 0 && live[0]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) merge_store_mask<T, PEER>(a, dst + e);
  }
}

// thread = one CSR row (one panorama pixel), CPT panorama (plane, frame) slices, Q payloads.
template <typename T, int CPT, int QMAX>
__global__ void __launch_bounds__(kBlock)
merge_bilinear_kernel(MergeArgs a, const int32_t* __restrict__ d_row_dst,
                      const int32_t* __restrict__ d_row_ptr, const int32_t* __restrict__ d_ent_src,
                      const float* __restrict__ d_ent_w, const int32_t* __restrict__ d_ftab) {
  const int t = blockIdx.x * kBlock + threadIdx.x;
  if (t >= a.T) return;
  const int dst = __ldg(d_row_dst + t);
  const int e0 = __ldg(d_row_ptr + t), e1 = __ldg(d_row_ptr + t + 1);
  if (e1 <= e0) return;
  const int v = __ldg(d_ent_src + e0) / a.hw;          // every entry of a row comes from one view
  const int PF = a.P * a.F;
  const int pf0 = blockIdx.y * CPT;
  int pl = pf0 / a.F, f = pf0 - pl * a.F;
  size_t base[CPT];                                    // payload offset of (job, pl, fo), pixel 0
  bool live[CPT];
  bool any = false;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    int job, fo;
    live[c] = pf0 + c < PF && merge_source(a, d_ftab, f, v, job, fo);
    if (live[c]) base[c] = ((static_cast<size_t>(job) * a.P + pl) * a.f_len + fo) * a.hw;
    any |= live[c];
    if (++f == a.F) { f = 0; ++pl; }
  }
  if (!any) return;
  float acc[QMAX][CPT];
#pragma unroll
  for (int q = 0; q < QMAX; ++q)
#pragma unroll
    for (int c = 0; c < CPT; ++c) acc[q][c] = 0.0f;
  float ws = 0.0f;
  for (int e = e0; e < e1; ++e) {
    const int k = __ldg(d_ent_src + e) - v * a.hw;
    const float w = __ldg(d_ent_w + e);
    ws = __fadd_rn(ws, w);
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      if (live[c]) {
#pragma unroll
        for (int q = 0; q < QMAX; ++q)
          if (q < a.Q) {
            const float x = Elem<T>::to_f(ldg(static_cast<const T*>(a.pay[q]) + base[c] + k));
            acc[q][c] = __fadd_rn(acc[q][c], __fmul_rn(x, w));   // index_add_(view * w): two roundings
          }
      }
    }
  }
  if (!(ws > 0.0f)) return;
  pl = pf0 / a.F; f = pf0 - pl * a.F;
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    if (live[c]) {
      const size_t off = (static_cast<size_t>(pl) * a.F + f) * a.HW + dst;
#pragma unroll
      for (int q = 0; q < QMAX; ++q)
        if (q < a.Q) static_cast<T*>(a.pano[q])[off] = Elem<T>::from_f(__fdiv_rn(acc[q][c], ws));
      if (a.mask_mode == DS_MASK_FULL && pl < a.mask_planes) static_cast<T*>(a.mask)[off] = Elem<T>::one();
    }
    if (++f == a.F) { f = 0; ++pl; }
  }
  if (a.mask_mode == DS_MASK_PLANE && blockIdx.y == 0 && live[0]) static_cast<T*>(a.mask)[dst] = Elem<T>::one();
}

static int fill_merge_args(MergeArgs& a, void* const* h_panos, const void* const* h_payloads, int Q,
                           void* mask, int mask_mode, int mask_planes, int T, int vpf,
                           int job_begin, int job_count, const int32_t* d_ftab,
                           int P, int F, int HW, int f_len, int hw) {
  DS_REQUIRE(h_panos != nullptr && h_payloads != nullptr, DS_ERR_NULL);
  DS_REQUIRE(Q >= 0 && Q <= DS_MAX_PAYLOADS, DS_ERR_COUNT);
  DS_REQUIRE(T >= 0 && vpf > 0 && P > 0 && F > 0 && HW > 0 && f_len > 0 && hw > 0, DS_ERR_DIM);
  DS_REQUIRE(mask_mode == DS_MASK_NONE || mask_mode == DS_MASK_PLANE || mask_mode == DS_MASK_FULL, DS_ERR_DIM);
  DS_REQUIRE(mask_mode == DS_MASK_NONE || mask != nullptr, DS_ERR_NULL);
  DS_REQUIRE(mask_mode != DS_MASK_FULL || (mask_planes > 0 && mask_planes <= P), DS_ERR_DIM);
  DS_REQUIRE(Q > 0 || mask_mode != DS_MASK_NONE, DS_ERR_COUNT);
  DS_REQUIRE(job_begin >= 0 && job_count > 0, DS_ERR_DIM);
  DS_REQUIRE(d_ftab != nullptr || f_len == F, DS_ERR_DIM);
  DS_REQUIRE(mask_mode != DS_MASK_PLANE || d_ftab == nullptr, DS_ERR_DIM);
  for (int q = 0; q < DS_MAX_PAYLOADS; ++q) {
    a.pano[q] = q < Q ? h_panos[q] : nullptr;
    a.pay[q] = q < Q ? h_payloads[q] : nullptr;
    if (q < Q) DS_REQUIRE(a.pano[q] != nullptr && a.pay[q] != nullptr, DS_ERR_NULL);
  }
  a.mask = mask; a.mask_mode = mask_mode; a.mask_planes = mask_planes;
  a.Q = Q; a.T = T; a.vpf = vpf; a.job_begin = job_begin; a.job_count = job_count; a.P = P; a.F = F; a.HW = HW; a.f_len = f_len; a.hw = hw;
  return DS_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace ds

extern "C" int ds_gather_nearest(const void* pano, void* out, const int32_t* d_idx,
                                 const int32_t* d_jobs, int h_f0, int n_jobs, int P, int F, int HW,
                                 int f_len, int hw, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(pano != nullptr && out != nullptr && d_idx != nullptr, DS_ERR_NULL);
  DS_REQUIRE(n_jobs > 0 && P > 0 && F > 0 && HW > 0 && f_len > 0 && hw > 0, DS_ERR_DIM);
  DS_REQUIRE(n_jobs <= 65535 && h_f0 >= 0, DS_ERR_DIM);
  GatherDims d{P, F, HW, f_len, hw};
  constexpr int CPT = 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = (hw % 4 == 0) && aligned16(out) && aligned16(d_idx);
  const int CF = P * f_len;
  DS_REQUIRE(div_up(CF, CPT) <= 65535, DS_ERR_DIM);
  DS_DISPATCH_DTYPE(dtype, {
    if (vec) {
      dim3 grid(div_up(hw, kBlock * 4), n_jobs, div_up(CF, CPT));
      gather_nearest_kernel<ET, 4, CPT><<<grid, kBlock, 0, st>>>(static_cast<const ET*>(pano), static_cast<ET*>(out), d_idx, d_jobs, h_f0, d);
    } else {
      dim3 grid(div_up(hw, kBlock), n_jobs, div_up(CF, CPT));
      gather_nearest_kernel<ET, 1, CPT><<<grid, kBlock, 0, st>>>(static_cast<const ET*>(pano), static_cast<ET*>(out), d_idx, d_jobs, h_f0, d);
    }
  });
  return check_launch();
}

extern "C" int ds_gather_bilinear(const void* pano, void* out, const int32_t* d_idx4,
                                  const float* d_wt4, const int32_t* d_jobs, int h_f0, int n_jobs, int P,
                                  int F, int HW, int f_len, int hw, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(pano != nullptr && out != nullptr && d_idx4 != nullptr && d_wt4 != nullptr, DS_ERR_NULL);
  DS_REQUIRE(n_jobs > 0 && P > 0 && F > 0 && HW > 0 && f_len > 0 && hw > 0, DS_ERR_DIM);
  DS_REQUIRE(n_jobs <= 65535 && h_f0 >= 0, DS_ERR_DIM);
  GatherDims d{P, F, HW, f_len, hw};
  constexpr int CPT = 8;
  const int CF = P * f_len;
  DS_REQUIRE(div_up(CF, CPT) <= 65535, DS_ERR_DIM);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DS_DISPATCH_DTYPE(dtype, {
    dim3 grid(div_up(hw, kBlock), n_jobs, div_up(CF, CPT));
    gather_bilinear_kernel<ET, CPT><<<grid, kBlock, 0, st>>>(static_cast<const ET*>(pano), static_cast<ET*>(out), d_idx4, d_wt4, d_jobs, h_f0, d);
  });
  return check_launch();
}

static int merge_nearest_impl(void* const* h_panos, const void* const* h_payloads, int Q,
                              void* mask, int mask_mode, int mask_planes,
                              const int32_t* d_dst, const int32_t* d_src, int T,
                              const int32_t* d_ftab, int views_per_fwin, int job_begin,
                              int job_count, int P, int F, int HW, int f_len, int hw, int vec,
                              void* d_peer, int dtype, void* stream) {
  using namespace ds;
  MergeArgs a;
  int rc = fill_merge_args(a, h_panos, h_payloads, Q, mask, mask_mode, mask_planes, T,
                           views_per_fwin, job_begin, job_count, d_ftab, P, F, HW, f_len, hw);
  if (rc != DS_OK) return rc;
  a.peer = static_cast<const ds_peer_table_t*>(d_peer);
  DS_REQUIRE(d_src != nullptr, DS_ERR_NULL);
  DS_REQUIRE(vec == 1 || vec == 4, DS_ERR_DIM);
  if (T == 0) return DS_OK;
  if (vec == 4) {
    DS_REQUIRE(HW % 4 == 0 && hw % 4 == 0, DS_ERR_ALIGN);
    for (int q = 0; q < Q; ++q) DS_REQUIRE(aligned16(a.pano[q]) && aligned16(a.pay[q]), DS_ERR_ALIGN);
    if (mask_mode != DS_MASK_NONE) DS_REQUIRE(aligned16(mask), DS_ERR_ALIGN);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int PF = P * F;
  DS_DISPATCH_DTYPE(dtype, {
    if (vec == 4) {
      constexpr int CPT = 4;
      DS_REQUIRE(div_up(PF, CPT) <= 65535, DS_ERR_DIM);
      dim3 grid(div_up(T, kBlock), div_up(PF, CPT));
      if (d_peer != nullptr) merge_nearest_kernel<ET, 4, CPT, DS_MAX_PAYLOADS, true><<<grid, kBlock, 0, st>>>(a, d_dst, d_src, d_ftab);
      else merge_nearest_kernel<ET, 4, CPT, DS_MAX_PAYLOADS><<<grid, kBlock, 0, st>>>(a, d_dst, d_src, d_ftab);
    } else {
      constexpr int CPT = 8;
      DS_REQUIRE(div_up(PF, CPT) <= 65535, DS_ERR_DIM);
      dim3 grid(div_up(T, kBlock), div_up(PF, CPT));
      if (d_peer != nullptr) merge_nearest_kernel<ET, 1, CPT, DS_MAX_PAYLOADS, true><<<grid, kBlock, 0, st>>>(a, d_dst, d_src, d_ftab);
      else merge_nearest_kernel<ET, 1, CPT, DS_MAX_PAYLOADS><<<grid, kBlock, 0, st>>>(a, d_dst, d_src, d_ftab);
    }
  });
  return check_launch();
}

extern "C" int ds_merge_nearest(void* const* h_panos, const void* const* h_payloads, int Q,
                                void* mask, int mask_mode, int mask_planes,
                                const int32_t* d_dst, const int32_t* d_src, int T,
                                const int32_t* d_ftab, int views_per_fwin, int job_begin,
                                int job_count, int P, int F, int HW, int f_len, int hw, int vec,
                                int dtype, void* stream) {
  return merge_nearest_impl(h_panos, h_payloads, Q, mask, mask_mode, mask_planes, d_dst, d_src, T, d_ftab,
                            views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw, vec, nullptr, dtype, stream);
}

extern "C" int ds_merge_nearest_peer(void* const* h_panos, const void* const* h_payloads, int Q,
                                     void* mask, int mask_mode, int mask_planes,
                                     const int32_t* d_dst, const int32_t* d_src, int T,
                                     const int32_t* d_ftab, int views_per_fwin, int job_begin,
                                     int job_count, int P, int F, int HW, int f_len, int hw, int vec,
                                     void* d_peer, int dtype, void* stream) {
  DS_REQUIRE(d_peer != nullptr, DS_ERR_NULL);
  return merge_nearest_impl(h_panos, h_payloads, Q, mask, mask_mode, mask_planes, d_dst, d_src, T, d_ftab,
                            views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw, vec, d_peer, dtype, stream);
}

extern "C" int ds_merge_nearest_ddim(void* const* h_panos, const void* x, const void* eps_cond,
                                     const void* eps_uncond, const void* noise, const void* prev,
                                     const void* mask_view, const float* h_coef,
                                     void* mask, int mask_mode, int mask_planes,
                                     const int32_t* d_dst, const int32_t* d_src, int T,
                                     const int32_t* d_ftab, int views_per_fwin, int job_begin,
                                     int job_count, int P, int F, int HW, int f_len, int hw, int vec,
                                     void* d_peer, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(x != nullptr && eps_cond != nullptr && h_coef != nullptr, DS_ERR_NULL);
  const void* pays[2] = {x, eps_cond};
  MergeArgs a;
  int rc = fill_merge_args(a, h_panos, pays, 2, mask, mask_mode, mask_planes, T, views_per_fwin,
                           job_begin, job_count, d_ftab, P, F, HW, f_len, hw);
  if (rc != DS_OK) return rc;
  a.pay[2] = eps_uncond;
  a.pay[3] = noise;
  a.peer = static_cast<const ds_peer_table_t*>(d_peer);
  const DdimCoef dc{h_coef[0], h_coef[1], h_coef[2], h_coef[3], h_coef[4], h_coef[5]};
  DS_REQUIRE(prev == nullptr || mask_view != nullptr, DS_ERR_NULL);
  const PrevMix pv{prev, mask_view, h_coef[6], h_coef[7]};
  DS_REQUIRE(d_src != nullptr, DS_ERR_NULL);
  DS_REQUIRE(vec == 1 || vec == 4, DS_ERR_DIM);
  if (T == 0) return DS_OK;
  if (vec == 4) {
    DS_REQUIRE(HW % 4 == 0 && hw % 4 == 0, DS_ERR_ALIGN);
    for (int q = 0; q < 2; ++q) DS_REQUIRE(aligned16(a.pano[q]), DS_ERR_ALIGN);
    for (int q = 0; q < 4; ++q) DS_REQUIRE(aligned16(a.pay[q]), DS_ERR_ALIGN);
    DS_REQUIRE(aligned16(prev) && aligned16(mask_view), DS_ERR_ALIGN);
    if (mask_mode != DS_MASK_NONE) DS_REQUIRE(aligned16(mask), DS_ERR_ALIGN);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int PF = P * F;
  const bool peer = d_peer != nullptr, mix = prev != nullptr;
#define DS_LAUNCH_DDIM(VEC_, CPT_)                                                                          \
  do {                                                                                                      \
    DS_REQUIRE(div_up(PF, CPT_) <= 65535, DS_ERR_DIM);                                                      \
    dim3 grid(div_up(T, kBlock), div_up(PF, CPT_));                                                         \
    if (peer && mix) merge_nearest_ddim_kernel<ET, VEC_, CPT_, true, true><<<grid, kBlock, 0, st>>>(a, dc, pv, d_dst, d_src, d_ftab);   \
    else if (peer) merge_nearest_ddim_kernel<ET, VEC_, CPT_, true, false><<<grid, kBlock, 0, st>>>(a, dc, pv, d_dst, d_src, d_ftab);    \
    else if (mix) merge_nearest_ddim_kernel<ET, VEC_, CPT_, false, true><<<grid, kBlock, 0, st>>>(a, dc, pv, d_dst, d_src, d_ftab);     \
    else merge_nearest_ddim_kernel<ET, VEC_, CPT_, false, false><<<grid, kBlock, 0, st>>>(a, dc, pv, d_dst, d_src, d_ftab);             \
  } while (0)
  DS_DISPATCH_DTYPE(dtype, {
    if (vec == 4) DS_LAUNCH_DDIM(4, 2);
    else DS_LAUNCH_DDIM(1, 8);
  });
#undef DS_LAUNCH_DDIM
  return check_launch();
}

extern "C" int ds_merge_bilinear(void* const* h_panos, const void* const* h_payloads, int Q,
                                 void* mask, int mask_mode, int mask_planes,
                                 const int32_t* d_row_dst, const int32_t* d_row_ptr, int T,
                                 const int32_t* d_ent_src, const float* d_ent_w,
                                 const int32_t* d_ftab, int views_per_fwin, int job_begin,
                                 int job_count, int P, int F, int HW, int f_len, int hw, int dtype,
                                 void* stream) {
  using namespace ds;
  MergeArgs a;
  int rc = fill_merge_args(a, h_panos, h_payloads, Q, mask, mask_mode, mask_planes, T,
                           views_per_fwin, job_begin, job_count, d_ftab, P, F, HW, f_len, hw);
  if (rc != DS_OK) return rc;
  DS_REQUIRE(d_row_dst != nullptr && d_row_ptr != nullptr && d_ent_src != nullptr && d_ent_w != nullptr, DS_ERR_NULL);
  if (T == 0) return DS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  constexpr int CPT = 4;
  const int PF = P * F;
  DS_REQUIRE(div_up(PF, CPT) <= 65535, DS_ERR_DIM);
  DS_DISPATCH_DTYPE(dtype, {
    dim3 grid(div_up(T, kBlock), div_up(PF, CPT));
    merge_bilinear_kernel<ET, CPT, DS_MAX_PAYLOADS><<<grid, kBlock, 0, st>>>(a, d_row_dst, d_row_ptr, d_ent_src, d_ent_w, d_ftab);
  });
  return check_launch();
}
