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
//   pm_gather_kernel                get_view_tensor_no_interpolate / _interpolate, all views of a step
//   pm_merge_nearest_kernel         set_view_tensor_no_interpolation, all views, last writer wins
//   pm_merge_weighted_tiled_kernel  set_view_tensor_bilinear, all views (weighted splat + normalise)
//   pm_merge_bilinear_kernel        the same for any window frame count / row length
// and, with identity maps, the pack / unpack between the two layouts.
//
// Merges are in SCATTER form over a host-built list, sorted by view, of the view pixels
// (nearest) / panorama rows (bilinear) that survive the ordered sequence of writes, so no
// two threads ever write the same panorama atom: race-free without atomics, bit-identical
// to the reference's sequential order, and per-rank outputs have disjoint support when the
// job list is sharded.
//
// Multi-GPU (windows sharded over ranks): with a ds_peer_table_t every rank holds a
// replica of the panorama in peer-mapped (symmetric) memory and the merge kernels are the
// exchange step themselves: each surviving row is stored straight into EVERY replica over
// NVLink -- one multimem.st to the NVLS multicast alias, or one TMA bulk copy / peer st.global per
// replica -- so there is no zeroed sum buffer, no all-reduce and no select pass.
// Cross-GPU ordering is two monotonic counters per peer: a 32-thread kernel after every gather /
// merge publishes this rank's count into every peer's table (st.release.sys), and the data
// kernels wait on them (ld.acquire.sys, prefetched so the round trip overlaps their loads):
//   gather  reads the panorama only when every peer has pushed the previous step's rows
//   merge   stores rows only when every peer has finished reading its replica
//
// Scheduling: these kernels run for 10-20 us, so wave quantisation and tails matter as much
// as bandwidth: the gather runs persistent CTAs with an even static split, the merges one
// small tile per CTA under the hardware CTA scheduler (measured faster than a static split).
#include "ds_common.cuh"

namespace ds {

constexpr int kPmPlanes = 4;     // planes per atom (B*C of the latent); other counts use sphere.cu
constexpr int kPmKT = 64;        // view pixels per gather work item
constexpr int kPmET = 32;        // entries (nearest) / panorama rows (weighted) per merge CTA
constexpr int kPmEC = 64;        // padded CSR entries per tile of the tiled weighted merge
constexpr int kPmTR = 32;        // panorama rows per tile of the tiled weighted merge (64 measured slower on cfg2: 15.3 -> 16.2 us)
constexpr int kWarps = kBlock / 32;

// Launch-shape knobs (ds_tune); defaults are the measured best on B200 for the bench workload.
static int g_gather_ctas_per_sm = 0;            // 0: occupancy limit
static int g_merge_publish = 0;                 // exchange mode: 0 = the host enqueues ds_pm_peer_sync after every gather / merge; 1 = deferred:
                                                // a data kernel counts itself as issued and the next kernel publishes it in its prologue
static int g_hw_const = 1;                      // 40 x 64 windows: kernels specialised on the window size (immediate load offsets)
static int g_gather_pipe = 2;                   // bit 1: copy-only gather, bit 2: 4-tap gather use the pipelined tile loop
static long long g_peer_timeout_ns = 60000000000ll;  // a cross-GPU wait longer than this (60 s: first-step autotune,
                                                       // lazy module loads, stragglers are all far shorter) aborts the kernel
static int g_pdl = 11;                          // programmatic dependent launch, bit mask: 1 copy-only gather, 2 4-tap gather,
                                                // 4 nearest merge, 8 weighted merges.  Measured (tools/kbench.py, cfg2): the
                                                // 4-tap gather / weighted merge chain gains 7 % (31.7 -> 29.6 us); with BOTH
                                                // nearest kernels on PDL their chain loses 8 % (25.5 -> 27.7 us), so the
                                                // nearest merge keeps plain stream order

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

template <> struct Atom<__half> {
  using type = uint2;
  static __device__ __forceinline__ void unpack(const uint2& a, float (&v)[4]) {
    const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
    const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
  }
  static __device__ __forceinline__ uint2 pack(const float (&v)[4]) {
    const __half2 lo = __floats2half2_rn(v[0], v[1]), hi = __floats2half2_rn(v[2], v[3]);
    return make_uint2(*reinterpret_cast<const unsigned*>(&lo), *reinterpret_cast<const unsigned*>(&hi));
  }
};

// raw element <-> atom component without a float round trip (pure copies)
template <typename T> __device__ __forceinline__ void atom_split(const typename Atom<T>::type& a, T (&v)[4]);
template <> __device__ __forceinline__ void atom_split<float>(const float4& a, float (&v)[4]) { v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
template <> __device__ __forceinline__ void atom_split<__nv_bfloat16>(const uint2& a, __nv_bfloat16 (&v)[4]) {
  v[0] = __ushort_as_bfloat16(static_cast<unsigned short>(a.x & 0xffffu)); v[1] = __ushort_as_bfloat16(static_cast<unsigned short>(a.x >> 16));
  v[2] = __ushort_as_bfloat16(static_cast<unsigned short>(a.y & 0xffffu)); v[3] = __ushort_as_bfloat16(static_cast<unsigned short>(a.y >> 16));
}
template <> __device__ __forceinline__ void atom_split<__half>(const uint2& a, __half (&v)[4]) {
  v[0] = __ushort_as_half(static_cast<unsigned short>(a.x & 0xffffu)); v[1] = __ushort_as_half(static_cast<unsigned short>(a.x >> 16));
  v[2] = __ushort_as_half(static_cast<unsigned short>(a.y & 0xffffu)); v[3] = __ushort_as_half(static_cast<unsigned short>(a.y >> 16));
}
template <typename T> __device__ __forceinline__ typename Atom<T>::type atom_join(const T (&v)[4]);
template <> __device__ __forceinline__ float4 atom_join<float>(const float (&v)[4]) { return make_float4(v[0], v[1], v[2], v[3]); }
template <> __device__ __forceinline__ uint2 atom_join<__nv_bfloat16>(const __nv_bfloat16 (&v)[4]) {
  return make_uint2(__bfloat16_as_ushort(v[0]) | (static_cast<unsigned>(__bfloat16_as_ushort(v[1])) << 16),
                    __bfloat16_as_ushort(v[2]) | (static_cast<unsigned>(__bfloat16_as_ushort(v[3])) << 16));
}
template <> __device__ __forceinline__ uint2 atom_join<__half>(const __half (&v)[4]) {
  return make_uint2(__half_as_ushort(v[0]) | (static_cast<unsigned>(__half_as_ushort(v[1])) << 16),
                    __half_as_ushort(v[2]) | (static_cast<unsigned>(__half_as_ushort(v[3])) << 16));
}
template <typename T> __device__ __forceinline__ typename Atom<T>::type atom_ones();
template <> __device__ __forceinline__ float4 atom_ones<float>() { return make_float4(1.f, 1.f, 1.f, 1.f); }
template <> __device__ __forceinline__ uint2 atom_ones<__nv_bfloat16>() { return make_uint2(0x3f803f80u, 0x3f803f80u); }
template <> __device__ __forceinline__ uint2 atom_ones<__half>() { return make_uint2(0x3c003c00u, 0x3c003c00u); }

template <typename A> __device__ __forceinline__ A ldg_atom(const A* p) { return __ldg(p); }

// mix_latents_with_mask on one element, every op rounded on its own and -- like the reference's
// elementwise tensor ops -- in the tensor dtype T.
template <typename T>
__device__ __forceinline__ float pm_mix1(float l1, float add, float m, float r, float omr) {
  auto rt = [](float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); };
  const float a = rt(__fmul_rn(l1, omr));
  const float b = rt(__fmul_rn(add, r));
  const float mixed = rt(__fadd_rn(a, b));
  const float keep = rt(__fmul_rn(l1, rt(__fsub_rn(1.0f, m))));
  return rt(__fadd_rn(keep, rt(__fmul_rn(mixed, m))));
}

// ------------------------------------------- programmatic dependent launch (PDL) ----
// The gather / merge kernels of consecutive steps form a chain of 10-20 us kernels; with PDL the
// next kernel's CTAs are scheduled while the current one drains, run their prologue (plan / map
// loads, which no earlier kernel writes) and only then wait for the predecessor's memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait_predecessors() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------- async-proxy (TMA bulk) helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// Make this thread's generic-proxy shared-memory writes visible to the async proxy (TMA).
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 1-D bulk copy shared -> global (SASS UBLKCP): bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_commit_and_wait_read() {
  bulk_commit();
  bulk_wait_read();
}
// 4-byte asynchronous copy global -> shared (SASS LDGSTS): no register holds the value, so a
// prefetch costs nothing while it is in flight.  bytes == 0 writes a zero instead.
__device__ __forceinline__ void cp_async4(void* sdst, const void* gsrc, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" :: "r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ------------------------------------------------------- cross-GPU ordering (peers) ----
__device__ __forceinline__ int ld_acquire_sys(const int32_t* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int32_t* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
// NVLS multicast stores: one store to a multicast address lands in EVERY GPU's replica (the
// NVSwitch replicates it), so the exchange costs one store per row instead of one per peer.
__device__ __forceinline__ void multimem_st(void* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void multimem_st(void* p, const uint2& v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};"
               :: "l"(p), "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)) : "memory");
}
__device__ __forceinline__ void multimem_st(void* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// A cross-GPU wait that outlives the timeout must never let the data phase run on a replica that
// is still being read / not yet complete: record the error word (host-visible through
// PanoramaState.peer_error) and abort the kernel -- the context is then in a sticky launch-failure
// state, so the next call through the C ABI fails loudly instead of computing on corrupted rows.
__device__ __forceinline__ void pm_peer_timeout(ds_peer_table_t* tab) {
  tab->error = 1;
  __threadfence_system();
  asm volatile("trap;");
}

// Counters of the exchange (ds_peer_table_t): issued_gathers / issued_merges count the gather / merge
// launches of this rank that have COMPLETED (bumped by the launch itself or by the sync kernel below);
// my_gathers / my_merges are how many of them have been PUBLISHED into every peer's table
// (st.release.sys into gathers_done[rank] / merges_done[rank] there).  Two protocols:
//   host-side (default)  the host enqueues ds_pm_peer_sync after every data kernel: it bumps the
//                        issued count and publishes it at once;
//   deferred (DS_TUNE_PEER_MERGE_PUBLISH = 1)  a data kernel bumps the issued count as its last act and
//                        the NEXT kernel on the state publishes it in its prologue (pm_publish_pending, CTA 0,
//                        after griddepcontrol.wait: the predecessor has retired, so its reads are done
//                        and its peer stores performed) -- no extra launches in the step.
// Either way a data kernel waits on `issued`: a gather reads the replica only when every peer has
// published as many merges as this rank has issued, a merge stores rows only when every peer has
// published as many gathers.

// Warp 0 of one CTA: publish whatever is issued but not yet published.  lane p talks to peer p.
__device__ __forceinline__ void pm_publish_pending(ds_peer_table_t* tab, bool gathers) {
  if (threadIdx.x >= 32) return;
  int32_t* pub = gathers ? &tab->my_gathers : &tab->my_merges;
  const int issued = gathers ? tab->issued_gathers : tab->issued_merges;
  const int have = *pub;
  __syncwarp();
  if (have >= issued) return;
  const size_t off = gathers ? offsetof(ds_peer_table_t, gathers_done) : offsetof(ds_peer_table_t, merges_done);
  __threadfence_system();
  const int p = threadIdx.x;
  if (p < tab->world && p != tab->rank)
    st_release_sys(reinterpret_cast<int32_t*>(reinterpret_cast<char*>(tab->ctl[p]) + off) + tab->rank, issued);
  if (p == 0) *pub = issued;
}

// Host-enqueued ordering kernel (32 threads; stream order does the rest):
//   DS_PEER_AFTER_GATHER / DS_PEER_AFTER_MERGE: one more gather / merge of this rank has completed ->
//     bump the issued count and publish it;
//   DS_PEER_WAIT_MERGES: bump nothing, publish what is pending and wait for the peers' merges.
// wait != 0 additionally blocks until every peer has published as much (a full barrier); the data
// kernels normally do that part themselves (pm_peer_spin).  A wait longer than timeout_ns sets
// tab->error and traps (pm_peer_timeout) instead of hanging.
__global__ void pm_peer_sync_kernel(ds_peer_table_t* tab, int phase, int wait, long long timeout_ns) {
  const int p = threadIdx.x;
  const bool gathers = phase == DS_PEER_AFTER_GATHER;
  int32_t* issued = gathers ? &tab->issued_gathers : &tab->issued_merges;
  const int c = *issued + (phase != DS_PEER_WAIT_MERGES ? 1 : 0);
  __syncwarp();
  if (p == 0) *issued = c;
  __syncwarp();
  if (phase == DS_PEER_WAIT_MERGES) pm_publish_pending(tab, true);
  pm_publish_pending(tab, gathers);
  if (wait && p < tab->world && p != tab->rank) {
    const int32_t* done = (gathers ? tab->gathers_done : tab->merges_done) + p;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(done) < c) {
      if (global_ns() - t0 > static_cast<unsigned long long>(timeout_ns)) pm_peer_timeout(tab);
    }
  }
}

// Deferred protocol: last act of a data kernel (one thread of the grid): this launch is done as far as
// the NEXT kernel on the stream is concerned, which is the one that publishes it.
__device__ __forceinline__ void pm_count_issued(ds_peer_table_t* tab, bool gathers) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    int32_t* issued = gathers ? &tab->issued_gathers : &tab->issued_merges;
    *issued = *issued + 1;
  }
}

// In-kernel half of the ordering: lane p of every CTA waits until peer p's counter `done[p]` has
// reached `want` (normally true already, or a few hundred ns away); `v` is a value of the flag
// prefetched earlier so that the first round trip overlaps other work.
__device__ __forceinline__ void pm_peer_spin(ds_peer_table_t* tab, const int32_t* done, int v, int want,
                                             long long timeout_ns) {
  if (static_cast<int>(threadIdx.x) < tab->world && static_cast<int>(threadIdx.x) != tab->rank) {
    const unsigned long long t0 = global_ns();
    while (v < want) {
      v = ld_acquire_sys(done + threadIdx.x);
      if (global_ns() - t0 > static_cast<unsigned long long>(timeout_ns)) pm_peer_timeout(tab);
    }
  }
}

// ----------------------------------------------------------------------- gather ----
struct PmGatherArgs {
  int F, HW, f_len, hw;
  int n_tiles, tiles_per_job;
  int publish;                   // exchange mode, deferred protocol: count this gather as issued at the end (DS_TUNE_PEER_MERGE_PUBLISH)
};

// Persistent CTAs; work item = (tile of kPmKT view pixels, job).
// Load: warp lanes = 4 pixel groups x 8 frames; a lane loads the atoms of 4 CONSECUTIVE view
// pixels of one window frame (each request: 4 panorama rows x 128 B), transposes them in
// registers and stores, per plane, the 4 pixels as one 16-byte vector into the shared tile
// [plane*f_len + fo][k] (row stride kPmKT + 16 B: conflict-free).
// Store: every tile row is 256 contiguous bytes of the window tensor and leaves through one
// TMA bulk copy (cp.async.bulk shared -> global), issued by one thread per row.
// FL > 0: f_len known at compile time; FL == 0: runtime f_len.  MINB: min resident CTAs per SM
// (register budget).
// PIPE: software pipeline over the CTA's tiles.  Per tile the plain loop is a serial chain -- job
// descriptor -> map entries -> barrier -> tap loads -> transpose -> barrier -> bulk store -> wait
// for the copy to have read the tile -- and a CTA walks it 2-3 times, so the kernel is bound by
// that latency, not by bandwidth (3 instead of 4 CTAs per SM costs 25 %).  The pipelined loop
// takes the first three links and the last one out of the chain: the NEXT tile's map entries
// arrive through cp.async (LDGSTS, no registers) into the other half of a double-buffered
// s_idx / s_wt while this tile's taps are in flight, the job descriptor is fetched two tiles
// ahead, and the wait for the bulk copy moves to the top of the next tile.
// HWC: view pixels per window known at compile time (2560 = 40 x 64, see the merges), else 0: the tile -> job
// divisions and every window-row stride become constants.
template <typename T, int TAPS, int FL, int MINB, bool PIPE, int HWC = 0>
__global__ void __launch_bounds__(kBlock, MINB)
pm_gather_kernel(const typename Atom<T>::type* __restrict__ pm, T* __restrict__ out,
                 const int32_t* __restrict__ d_idx, const float* __restrict__ d_wt,
                 const int32_t* __restrict__ d_jobs, int h_f0, PmGatherArgs a, ds_peer_table_t* __restrict__ peer,
                 long long peer_timeout_ns) {
  using A = typename Atom<T>::type;
  static_assert(TAPS * kPmKT <= kBlock, "one map entry per thread");
  pdl_launch_dependents();
  bool waited = false;                           // predecessors (the previous merge) not yet awaited
  constexpr int VE = 16 / sizeof(T);            // elements per 16-byte vector
  constexpr int S = kPmKT + VE;                 // padded row stride (elements): 4 banks mod 32
  constexpr int NB = PIPE ? 2 : 1;              // map-entry buffers
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int f_len = FL > 0 ? FL : a.f_len;
  const int hw = HWC > 0 ? HWC : a.hw;
  const int tpj = HWC > 0 ? (HWC + kPmKT - 1) / kPmKT : a.tiles_per_job;
  const int R = kPmPlanes * f_len;
  T* tile = reinterpret_cast<T*>(smem_raw);                                   // [R][S]
  int32_t* s_idx0 = reinterpret_cast<int32_t*>(tile + R * S);                 // [NB][TAPS][KT]
  float* s_wt0 = reinterpret_cast<float*>(s_idx0 + NB * TAPS * kPmKT);        // [NB][TAPS][KT] (bilinear)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kgl = lane >> 3, fl = lane & 7;
  const int n_fblk = (f_len + 7) >> 3;
  const int n_it = (kPmKT / 16) * n_fblk;       // warp iterations: 4 pixel groups x 8 frames each

  // (map row, first frame) of the job of tile tt
  auto job_of = [&](int tt, int& map, int& f0) {
    const int job = tt / tpj;
    map = job;
    f0 = h_f0;
    if (d_jobs != nullptr) {
      map = __ldg(d_jobs + 2 * job);
      f0 = __ldg(d_jobs + 2 * job + 1);
    }
  };
  // PIPE: start the asynchronous copy of tile tt's map entries into buffer b
  auto stage_async = [&](int tt, int map, int b) {
    if (threadIdx.x < TAPS * kPmKT) {
      const int tap = threadIdx.x / kPmKT, kl = threadIdx.x % kPmKT;
      const int k = (tt % tpj) * kPmKT + kl;
      const int ok = k < hw ? 4 : 0;                                   // ragged last tile: zero-fill
      const size_t e = static_cast<size_t>(map * TAPS + tap) * hw + min(k, hw - 1);
      int32_t* di = s_idx0 + b * TAPS * kPmKT + threadIdx.x;
      if (d_idx != nullptr) cp_async4(di, d_idx + e, ok);
      else *di = ok ? k : 0;                                             // identity map: unpack
      if (TAPS > 1) cp_async4(s_wt0 + b * TAPS * kPmKT + threadIdx.x, d_wt + e, ok);
    }
  };

  if (peer != nullptr && a.n_tiles == 0) {       // an idle rank (no window of its own): still publish its last merge
    pdl_wait_predecessors();
    pm_publish_pending(peer, false);
  }
  int cur_f0 = 0, nxt_map = 0, nxt_f0 = 0, buf = 0;
  bool in_flight = false;                        // PIPE: this thread's bulk copies of the previous tile
  if (PIPE && static_cast<int>(blockIdx.x) < a.n_tiles) {
    int map;
    job_of(blockIdx.x, map, cur_f0);
    stage_async(blockIdx.x, map, 0);
    if (blockIdx.x + gridDim.x < a.n_tiles) job_of(blockIdx.x + gridDim.x, nxt_map, nxt_f0);
  }

  for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int job = t / tpj;
    const int k0 = (t - job * tpj) * kPmKT;
    const int32_t* s_idx = s_idx0 + buf * TAPS * kPmKT;
    const float* s_wt = s_wt0 + buf * TAPS * kPmKT;
    int f0;
    if (PIPE) {
      f0 = cur_f0 % a.F;
      cp_async_wait_all();                       // this tile's map entries (copied during the previous tile)
      if (in_flight) bulk_wait_read();           // the previous tile's rows have left the shared tile
    } else {
      int map;
      job_of(t, map, f0);
      f0 %= a.F;
      for (int i = threadIdx.x; i < TAPS * kPmKT; i += kBlock) {
        const int tap = i / kPmKT, kl = i % kPmKT;
        const int k = k0 + kl;
        int idx = 0;
        float w = 0.f;
        if (k < hw) {
          if (d_idx != nullptr) idx = __ldg(d_idx + (map * TAPS + tap) * hw + k);
          else idx = k;                                                  // identity map: unpack
          if (TAPS > 1) w = __ldg(d_wt + (map * TAPS + tap) * hw + k);
        }
        s_idx0[i] = idx;
        if (TAPS > 1) s_wt0[i] = w;
      }
    }
    if (!waited) {
      // the map entries above are static; everything below reads the panorama
      pdl_wait_predecessors();
      if (peer != nullptr) {  // exchange mode: every peer must have pushed the previous step's rows too
        if (blockIdx.x == 0) pm_publish_pending(peer, false);      // deferred protocol: my own last merge first
        pm_peer_spin(peer, peer->merges_done, -1, peer->issued_merges, peer_timeout_ns);
      }
      waited = true;
    }
    __syncthreads();        // also orders the previous tile's bulk-store reads before the STS below
    if (PIPE) {
      const int tn = t + gridDim.x;
      cur_f0 = nxt_f0;
      if (tn < a.n_tiles) {
        stage_async(tn, nxt_map, buf ^ 1);       // lands while this tile's taps are in flight
        if (tn + gridDim.x < a.n_tiles) job_of(tn + gridDim.x, nxt_map, nxt_f0);
      }
      buf ^= 1;
    }

    for (int it = warp; it < n_it; it += kWarps) {
      const int kgb = it / n_fblk, fb = it - kgb * n_fblk;
      const int kl = (kgb * 4 + kgl) * 4;         // first of this lane's 4 consecutive view pixels
      const int fo = fb * 8 + fl;
      if (fo >= f_len || k0 + kl >= hw) continue;
      int f = f0 + fo;
      f = f >= a.F ? f - a.F : f;
      A v[TAPS][4];
#pragma unroll
      for (int tp = 0; tp < TAPS; ++tp) {
        const int4 id = *reinterpret_cast<const int4*>(s_idx + tp * kPmKT + kl);
        v[tp][0] = ldg_atom(pm + static_cast<size_t>(id.x) * a.F + f);
        v[tp][1] = ldg_atom(pm + static_cast<size_t>(id.y) * a.F + f);
        v[tp][2] = ldg_atom(pm + static_cast<size_t>(id.z) * a.F + f);
        v[tp][3] = ldg_atom(pm + static_cast<size_t>(id.w) * a.F + f);
      }
      // keep every tap load in flight before the first use (ptxas otherwise interleaves
      // loads and FMAs to save registers, serialising the L2 round trips)
      asm volatile("" ::: "memory");
      T e[4][4];                                  // [pixel j][plane p]
      if (TAPS == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) atom_split<T>(v[0][j], e[j]);
      } else {
        // ATen grid_sample (CPU) order: nw*w + ne*w + sw*w + se*w contracted to an FMA chain
        float r[4][4], x[4];
        float w[4];
        *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(s_wt + kl);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          Atom<T>::unpack(v[0][j], x);
#pragma unroll
          for (int p = 0; p < 4; ++p) r[j][p] = __fmul_rn(x[p], w[j]);
        }
#pragma unroll
        for (int tp = 1; tp < TAPS; ++tp) {
          *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(s_wt + tp * kPmKT + kl);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            Atom<T>::unpack(v[tp][j], x);
#pragma unroll
            for (int p = 0; p < 4; ++p) r[j][p] = __fmaf_rn(x[p], w[j], r[j][p]);
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < 4; ++p) e[j][p] = Elem<T>::from_f(r[j][p]);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const T col[4] = {e[0][p], e[1][p], e[2][p], e[3][p]};
        *reinterpret_cast<A*>(tile + (p * f_len + fo) * S + kl) = atom_join<T>(col);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

    T* __restrict__ o = out + static_cast<size_t>(job) * R * hw + k0;
    if (hw % VE != 0) {
      // window rows are not 16-byte multiples (odd view sizes): element-wise copy of the tile
      const int kt = min(kPmKT, hw - k0);
      for (int u = threadIdx.x; u < R * kPmKT; u += kBlock) {
        const int row = u / kPmKT, c = u - row * kPmKT;
        if (c < kt) o[static_cast<size_t>(row) * hw + c] = tile[row * S + c];
      }
      __syncthreads();                            // the tile is rewritten by the next work item
      continue;
    }
    const uint32_t row_bytes = static_cast<uint32_t>(min(kPmKT, hw - k0) * sizeof(T));
    bool issued = false;
    for (int row = threadIdx.x; row < R; row += kBlock) {
      bulk_store(o + static_cast<size_t>(row) * hw, tile + row * S, row_bytes);
      issued = true;
    }
    if (PIPE) {
      if (issued) bulk_commit();
      in_flight = issued;
    } else if (issued) {
      bulk_commit_and_wait_read();              // the tile is reused / freed only after the copy read it
    }
  }
  if (PIPE && in_flight) bulk_wait_read();      // shared memory must outlive the last copy
  if (peer != nullptr && a.publish) pm_count_issued(peer, true);   // deferred protocol: the next kernel publishes it
}

// -------------------------------------------------- gather + re-noise + mask mix ----
// The sequential / grouped regime of the reference loop reads, per view, the latent AND the
// denoised-mask panorama and blends a re-noised copy of the window into the already-denoised
// part (t2v_sphere_panorama_pipeline.py:194-221, i2v_...:330-365):
//   view   = get_view_tensor_no_interpolate(latent)                 panorama_tensor_utils.py:53-70
//   m      = get_view_tensor_no_interpolate(mask panorama)          (nearest read, same map family)
//   noised = c * view + s * eps                                      pipeline/scheduler.py:98-110
//   out    = mix_latents_with_mask(view, noised, m, ratio)          utils/tensor_utils.py:19-39
// One launch does all four for every job of the group.  eps is INJECTED (window layout; the
// reference draws randn_like inside re_noise -- SURVEY Q6).  Optional extra outputs, both needed by
// the merge prologue (PmDdim::prev / mview): out_raw = the unmixed view (view_latent_prev_denoise),
// out_mask = m ([job][hw] for the [1,H,W] T2V mask plane, window layout for the I2V mask).
// Every elementwise op is rounded on its own in T, like the reference's tensor ops.
struct PmRenoiseArgs {
  const void* eps;
  const void* mask;              // MM == DS_MASK_PLANE: [H*W] elements; DS_MASK_FULL: pixel-major panorama
  void* out_raw;                 // or NULL
  void* out_mask;                // or NULL
  float c, s, r, omr;
};

// One pass = one quantity through the [plane*f_len + fo][k] shared tile and out to `dst`.
// WHAT: 0 mixed window, 1 raw window, 2 mask view (DS_MASK_FULL only).
template <typename T, int MM, int WHAT>
__device__ __forceinline__ void pm_renoise_pass(const typename Atom<T>::type* __restrict__ pm, T* __restrict__ dst,
                                                const int32_t* s_idx, const int32_t* s_midx, T* tile,
                                                const PmGatherArgs& a, const PmRenoiseArgs& e, int job, int k0, int f0) {
  using A = typename Atom<T>::type;
  constexpr int VE = 16 / sizeof(T);
  constexpr int S = kPmKT + VE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kgl = lane >> 3, fl = lane & 7;
  const int f_len = a.f_len, R = kPmPlanes * f_len;
  const int n_fblk = (f_len + 7) >> 3;
  const int n_it = (kPmKT / 16) * n_fblk;
  auto rt = [](float v) { return Elem<T>::to_f(Elem<T>::from_f(v)); };
  for (int it = warp; it < n_it; it += kWarps) {
    const int kgb = it / n_fblk, fb = it - kgb * n_fblk;
    const int kl = (kgb * 4 + kgl) * 4;
    const int fo = fb * 8 + fl;
    if (fo >= f_len || k0 + kl >= a.hw) continue;
    int f = f0 + fo;
    f = f >= a.F ? f - a.F : f;
    float val[4][4];                                            // [pixel j][plane p]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + kl + j;
      if (k >= a.hw) {
#pragma unroll
        for (int p = 0; p < 4; ++p) val[j][p] = 0.f;
        continue;
      }
      float x[4], m[4];
      if (WHAT != 2) Atom<T>::unpack(ldg_atom(pm + static_cast<size_t>(s_idx[kl + j]) * a.F + f), x);
      if (WHAT != 1) {
        if (MM == DS_MASK_PLANE) {
          const float mp = Elem<T>::to_f(ldg(static_cast<const T*>(e.mask) + s_midx[kl + j]));
          m[0] = m[1] = m[2] = m[3] = mp;
        } else {
          Atom<T>::unpack(ldg_atom(static_cast<const A*>(e.mask) + static_cast<size_t>(s_midx[kl + j]) * a.F + f), m);
        }
      }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (WHAT == 1) {
          val[j][p] = x[p];
        } else if (WHAT == 2) {
          val[j][p] = m[p];
        } else {
          const float ep = Elem<T>::to_f(ldg(static_cast<const T*>(e.eps) +
                                             (static_cast<size_t>(job) * R + p * f_len + fo) * a.hw + k));
          // c, s: 0-dim float32 tensors in the reference, cast to the latent dtype by the multiply (python-float
          // ratios of the mix are NOT: torch keeps wrapped scalars in the op's float math type)
          const float noised = rt(__fadd_rn(rt(__fmul_rn(rt(e.c), x[p])), rt(__fmul_rn(rt(e.s), ep))));
          val[j][p] = pm_mix1<T>(x[p], noised, m[p], e.r, e.omr);
        }
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const T col[4] = {Elem<T>::from_f(val[0][p]), Elem<T>::from_f(val[1][p]), Elem<T>::from_f(val[2][p]),
                        Elem<T>::from_f(val[3][p])};
      *reinterpret_cast<A*>(tile + (p * f_len + fo) * S + kl) = atom_join<T>(col);
    }
  }
  __syncthreads();
  // tile rows -> window tensor (plain coalesced stores: this is the small-group path)
  T* __restrict__ o = dst + static_cast<size_t>(job) * R * a.hw + k0;
  const int kt = min(kPmKT, a.hw - k0);
  for (int u = threadIdx.x; u < R * kPmKT; u += kBlock) {
    const int row = u / kPmKT, c = u - row * kPmKT;
    if (c < kt) o[static_cast<size_t>(row) * a.hw + c] = tile[row * S + c];
  }
  __syncthreads();
}

template <typename T, int MM>
__global__ void __launch_bounds__(kBlock)
pm_gather_renoise_kernel(const typename Atom<T>::type* __restrict__ pm, T* __restrict__ out,
                         const int32_t* __restrict__ d_idx, const int32_t* __restrict__ d_midx,
                         const int32_t* __restrict__ d_jobs, int h_f0, PmGatherArgs a, PmRenoiseArgs e,
                         ds_peer_table_t* __restrict__ peer, long long peer_timeout_ns) {
  constexpr int VE = 16 / sizeof(T);
  constexpr int S = kPmKT + VE;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int R = kPmPlanes * a.f_len;
  T* tile = reinterpret_cast<T*>(smem_raw);                                   // [R][S]
  int32_t* s_idx = reinterpret_cast<int32_t*>(tile + R * S);                  // [KT] latent map
  int32_t* s_midx = s_idx + kPmKT;                                            // [KT] mask map
  if (peer != nullptr) {
    if (blockIdx.x == 0) pm_publish_pending(peer, false);
    pm_peer_spin(peer, peer->merges_done, -1, peer->issued_merges, peer_timeout_ns);
  }
  for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
    const int job = t / a.tiles_per_job;
    const int k0 = (t - job * a.tiles_per_job) * kPmKT;
    int map = job, f0 = h_f0;
    if (d_jobs != nullptr) {
      map = __ldg(d_jobs + 2 * job);
      f0 = __ldg(d_jobs + 2 * job + 1);
    }
    f0 %= a.F;
    if (threadIdx.x < kPmKT) {
      const int k = min(k0 + static_cast<int>(threadIdx.x), a.hw - 1);
      s_idx[threadIdx.x] = __ldg(d_idx + static_cast<size_t>(map) * a.hw + k);
      s_midx[threadIdx.x] = __ldg(d_midx + static_cast<size_t>(map) * a.hw + k);
    }
    __syncthreads();
    if (MM == DS_MASK_PLANE && e.out_mask != nullptr && threadIdx.x < kPmKT && k0 + threadIdx.x < a.hw)
      static_cast<T*>(e.out_mask)[static_cast<size_t>(job) * a.hw + k0 + threadIdx.x] =
          ldg(static_cast<const T*>(e.mask) + s_midx[threadIdx.x]);
    pm_renoise_pass<T, MM, 0>(pm, out, s_idx, s_midx, tile, a, e, job, k0, f0);
    if (e.out_raw != nullptr) pm_renoise_pass<T, MM, 1>(pm, static_cast<T*>(e.out_raw), s_idx, s_midx, tile, a, e, job, k0, f0);
    if (MM == DS_MASK_FULL && e.out_mask != nullptr)
      pm_renoise_pass<T, MM, 2>(pm, static_cast<T*>(e.out_mask), s_idx, s_midx, tile, a, e, job, k0, f0);
  }
  if (peer != nullptr && a.publish) pm_count_issued(peer, true);
}

// ------------------------------------------------------------------------ merges ----
// Fused denoise epilogue (DD kernels): the two payloads are not read but computed, per surviving
// element only, from the window latent x and the UNet output(s), exactly like the reference's
// elementwise torch chain (every op rounded on its own, no FMA):
//   e       = eu + g * (ec - eu)                              classifier-free guidance
//                                                              (t2v_sphere_panorama_pipeline.py:258)
//   pred_x0 = (x - s1m * e) / sa                              pipeline/scheduler.py:84
//   x_prev  = (sp * pred_x0 + sd * e) + sig * noise           :86-89  (sig * noise only if given)
// pay[0] = x, pay[1] = ec, pay[2] = eu or NULL (then e = ec), pay[3] = noise or NULL;
// pm[0] receives x_prev (latent panorama), pm[1] pred_x0 (denoised panorama).
struct PmDdim {
  float sa, s1m, sp, sd, sig, g;
  // merge_prev_denoised_ratio mix that precedes the latent set in the I2V loop
  // (i2v_sphere_panorama_pipeline.py:430-435, :938-943):
  //   x_prev <- mix_latents_with_mask(x_prev, prev, m, r)      utils/tensor_utils.py:19-39
  // prev  = the window as gathered (view_latent_prev_denoise, :345), window layout like x;
  // mview = the denoised-mask view read at gather time (:347-355): [job][hw] (mview_full == 0,
  //         T2V plane mask) or window layout (mview_full == 1, I2V latent-shaped mask).
  const void* prev;
  const void* mview;
  int mview_full;
  float mr, momr;                // mix ratio r and (1 - r), formed in double on the host like the reference
};

template <typename T>
__device__ __forceinline__ void pm_ddim(const PmDdim& c, const T (&x)[4], const T (&ec)[4], const T (&eu)[4],
                                        const T (&nz)[4], bool has_eu, bool has_nz, float (&x_prev)[4],
                                        float (&pred_x0)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float e = Elem<T>::to_f(ec[p]);
    if (has_eu) {
      const float u = Elem<T>::to_f(eu[p]);
      e = __fadd_rn(u, __fmul_rn(c.g, __fsub_rn(e, u)));
    }
    const float x0 = __fdiv_rn(__fsub_rn(Elem<T>::to_f(x[p]), __fmul_rn(c.s1m, e)), c.sa);
    float xp = __fadd_rn(__fmul_rn(c.sp, x0), __fmul_rn(c.sd, e));
    if (has_nz) xp = __fadd_rn(xp, __fmul_rn(c.sig, Elem<T>::to_f(nz[p])));
    pred_x0[p] = x0;
    x_prev[p] = xp;
  }
}

// x_prev (4 planes of one window element) <- mix(x_prev, prev, mask view, r).  o = element offset of
// (job, plane 0, frame fo, k) in window layout, o0 = the same for frame 0 (job*4*plane_stride + k).
template <typename T>
__device__ __forceinline__ void pm_prev_mix(const PmDdim& c, int o, int o0, int plane_stride, int hw, float (&xp)[4]) {
  const T* __restrict__ prev = static_cast<const T*>(c.prev);
  const T* __restrict__ mv = static_cast<const T*>(c.mview);
  float m_plane = 0.f;
  if (!c.mview_full) {                           // [job][hw]: job = o0 / (4*plane_stride), k = o0 mod plane_stride mod hw
    const int job = o0 / (kPmPlanes * plane_stride), k = o0 - job * kPmPlanes * plane_stride;
    m_plane = Elem<T>::to_f(ldg(mv + job * hw + k));
  }
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const float m = c.mview_full ? Elem<T>::to_f(ldg(mv + o + p * plane_stride)) : m_plane;
    const float l1 = Elem<T>::to_f(Elem<T>::from_f(xp[p]));     // the reference mixes the rounded x_prev tensor
    xp[p] = pm_mix1<T>(l1, Elem<T>::to_f(ldg(prev + o + p * plane_stride)), m, c.mr, c.momr);
  }
}

struct PmMergeArgs {
  void* pm[DS_MAX_PAYLOADS];
  const void* pay[DS_MAX_PAYLOADS];
  void* mask;
  int mask_mode;
  int Q, n_ent;
  int vpf, job_begin, job_count;
  int F, HW, f_len, hw;
  ds_peer_table_t* peer;         // NULL: single replica (pm / mask above)
  int publish;                   // exchange mode, deferred protocol: count this merge as issued at the end
  long long peer_timeout_ns;
  PmDdim dd;                     // DD kernels only
};

// dynamic indexing of a kernel-parameter array would force a local-memory copy of the struct
template <typename PT> __device__ __forceinline__ PT pm_sel(PT const (&arr)[DS_MAX_PAYLOADS], int q) {
  PT r = arr[0];
#pragma unroll
  for (int i = 1; i < DS_MAX_PAYLOADS; ++i) r = q == i ? arr[i] : r;
  return r;
}

// Replica addresses of this launch, staged once per CTA: [r][q] panoramas, [r][DS_MAX_PAYLOADS] mask.
struct PmReplicas {
  unsigned long long ptr[DS_MAX_PEERS][DS_MAX_PAYLOADS + 1];
  unsigned long long mc[DS_MAX_PAYLOADS + 1];     // multicast alias of [q] / mask, 0 = none
  int world;
};

// Every merge CTA first (peer mode): stage the replica table in shared memory and prefetch this
// lane's peer flag.  No barrier here: the table is read, and pm_merge_wait called, only right
// before the __syncthreads that precedes the store phase.
struct PmPeerFlag {
  int v, want;
};
template <bool PEER>
__device__ __forceinline__ PmPeerFlag pm_merge_begin(const PmMergeArgs& a, PmReplicas& rep) {
  PmPeerFlag f{0, 0};
  if (!PEER) return f;
  if (threadIdx.x == 0) rep.world = a.peer->world;
  if (threadIdx.x >= 64 && threadIdx.x < 64 + DS_MAX_PAYLOADS + 1) {
    const int q = threadIdx.x - 64;
    rep.mc[q] = q < DS_MAX_PAYLOADS ? a.peer->mc_pm[q] : a.peer->mc_mask;
  }
  if (threadIdx.x < DS_MAX_PEERS * (DS_MAX_PAYLOADS + 1)) {
    const int r = threadIdx.x / (DS_MAX_PAYLOADS + 1), q = threadIdx.x % (DS_MAX_PAYLOADS + 1);
    rep.ptr[r][q] = q < DS_MAX_PAYLOADS ? a.peer->pm[r][q] : a.peer->mask[r];
  }
  if (blockIdx.x == 0 && blockIdx.y == 0) pm_publish_pending(a.peer, true);   // deferred protocol: my own last gather first
  if (static_cast<int>(threadIdx.x) < a.peer->world && static_cast<int>(threadIdx.x) != a.peer->rank) {
    f.want = a.peer->issued_gathers;
    f.v = ld_acquire_sys(a.peer->gathers_done + threadIdx.x);
  }
  return f;
}
// ... and right before that barrier: every peer must have finished reading its replica (this
// step's gather) before rows land in it.
template <bool PEER>
__device__ __forceinline__ void pm_merge_wait(const PmMergeArgs& a, const PmPeerFlag& f) {
  if (PEER) pm_peer_spin(a.peer, a.peer->gathers_done, f.v, f.want, a.peer_timeout_ns);
}

// Store one atom of payload q at element offset `o` of the panorama -- into the single local
// replica, or into every rank's replica (peer stores over NVLink).
template <typename A, bool PEER>
__device__ __forceinline__ void pm_store(const PmMergeArgs& a, const PmReplicas& rep, int q, size_t o, const A& v) {
  if (!PEER) {
    static_cast<A*>(pm_sel(a.pm, q))[o] = v;
  } else if (rep.mc[q] != 0) {
    multimem_st(reinterpret_cast<A*>(rep.mc[q]) + o, v);
  } else {
    for (int r = 0; r < rep.world; ++r) reinterpret_cast<A*>(rep.ptr[r][q])[o] = v;
  }
}
template <typename A, bool PEER>
__device__ __forceinline__ void pm_store_mask_atom(const PmMergeArgs& a, const PmReplicas& rep, size_t o, const A& v) {
  if (!PEER) {
    static_cast<A*>(a.mask)[o] = v;
  } else if (rep.mc[DS_MAX_PAYLOADS] != 0) {
    multimem_st(reinterpret_cast<A*>(rep.mc[DS_MAX_PAYLOADS]) + o, v);
  } else {
    for (int r = 0; r < rep.world; ++r) reinterpret_cast<A*>(rep.ptr[r][DS_MAX_PAYLOADS])[o] = v;
  }
}
template <typename T, bool PEER>
__device__ __forceinline__ void pm_store_mask_plane(const PmMergeArgs& a, const PmReplicas& rep, int dst) {
  if (!PEER) {
    static_cast<T*>(a.mask)[dst] = Elem<T>::one();
  } else if (sizeof(T) == 4 && rep.mc[DS_MAX_PAYLOADS] != 0) {
    multimem_st(reinterpret_cast<float*>(rep.mc[DS_MAX_PAYLOADS]) + dst, 1.0f);
  } else {
    for (int r = 0; r < rep.world; ++r) reinterpret_cast<T*>(rep.ptr[r][DS_MAX_PAYLOADS])[dst] = Elem<T>::one();
  }
}

// Last statement of every exchange-mode merge CTA (also of CTAs that found no work).  Deferred protocol:
// count this merge as issued; the next kernel on the state publishes it once this one has retired.
template <bool PEER>
__device__ __forceinline__ void pm_merge_end(const PmMergeArgs& a) {
  if (PEER && a.publish) pm_count_issued(a.peer, false);
}

struct PmRange {
  int lo, hi;      // this CTA's slice of the entry (nearest) / row (weighted) list
  int s, fb;       // frame window and its first panorama frame
};

// The list is sorted by view and d_view_ptr[v] is the first entry of view v, so the jobs of
// this rank in frame window s = blockIdx.y are one contiguous slice; CTA b owns entries
// [b*32, b*32 + 32) of it.  false => nothing to do.
__device__ __forceinline__ bool pm_range(const PmMergeArgs& a, const int32_t* __restrict__ d_view_ptr,
                                         const int32_t* __restrict__ d_fbegin, PmRange& r) {
  r.s = blockIdx.y;
  r.fb = d_fbegin != nullptr ? __ldg(d_fbegin + r.s) % a.F : 0;
  int e_lo = 0, e_hi = a.n_ent;
  if (d_view_ptr != nullptr) {
    const int v_lo = min(max(a.job_begin - r.s * a.vpf, 0), a.vpf);
    const int v_hi = min(max(a.job_begin + a.job_count - r.s * a.vpf, 0), a.vpf);
    e_lo = __ldg(d_view_ptr + v_lo);
    e_hi = __ldg(d_view_ptr + v_hi);
  }
  r.lo = e_lo + blockIdx.x * kPmET;
  r.hi = min(r.lo + kPmET, e_hi);
  return r.hi > r.lo;
}

// Window frame fo of frame window s lands on panorama frame f; it is written only if s is the
// LAST frame window (in job order) containing f (d_ftab, maps.frame_table).
__device__ __forceinline__ bool pm_frame_live(const PmMergeArgs& a, const int32_t* __restrict__ d_ftab,
                                              const PmRange& r, int fo, int f_len, int& f) {
  f = r.fb + fo;                                 // fb already reduced modulo F, f_len <= F
  f = f >= a.F ? f - a.F : f;
  return d_ftab == nullptr || __ldg(d_ftab + f) == r.s * f_len + fo;
}

// entry -> payload element offset of (job, plane 0, frame 0, k)
__device__ __forceinline__ int pm_job_offset(const PmMergeArgs& a, int s, int src, int plane_stride) {
  const int v = src / a.hw, k = src - v * a.hw;
  return (s * a.vpf + v - a.job_begin) * kPmPlanes * plane_stride + k;
}

// CTA = 32 surviving view pixels.
// Load: warp <-> (payload q, window frame fo), lane <-> entry: 4 plane loads -> one atom ->
// shared tile[q][entry][fo].  Store: f_len consecutive lanes write one full panorama row.
// HWC: as in the weighted merge below -- 40 x 64 windows known at compile time, payload offsets become immediates.
template <typename T, int FL, bool DD, bool PEER, int HWC = 0>
__global__ void __launch_bounds__(kBlock, 5)
pm_merge_nearest_kernel(const PmMergeArgs a, const int32_t* __restrict__ d_view_ptr,
                        const int32_t* __restrict__ d_ent_src, const int32_t* __restrict__ d_ent_dst,
                        const int32_t* __restrict__ d_ftab, const int32_t* __restrict__ d_fbegin) {
  using A = typename Atom<T>::type;
  pdl_launch_dependents();
  __shared__ PmReplicas rep;
  PmRange rg;
  if (pm_range(a, d_view_ptr, d_fbegin, rg)) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int f_len = FL > 0 ? FL : a.f_len;
    const int SA = f_len | 1;                                          // odd row stride in atoms
    A* tile = reinterpret_cast<A*>(smem_raw);                          // [Q][32][SA]
    int32_t* s_dst = reinterpret_cast<int32_t*>(tile + (a.Q > 0 ? a.Q : 1) * kPmET * SA);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int plane_stride = f_len * (HWC > 0 ? HWC : a.hw);
    const int n_pairs = a.Q * f_len;
    const int cnt = rg.hi - rg.lo;
    const bool have = lane < cnt;
    int off = 0;
    if (have) {
      const int src = d_ent_src != nullptr ? __ldg(d_ent_src + rg.lo + lane) : rg.lo + lane;
      if (HWC > 0) {
        const int v = src / HWC, k = src - v * HWC;
        off = (rg.s * a.vpf + v - a.job_begin) * kPmPlanes * plane_stride + k;
      } else {
        off = pm_job_offset(a, rg.s, src, plane_stride);
      }
      if (warp == 0) {
        const int dst = d_ent_dst != nullptr ? __ldg(d_ent_dst + rg.lo + lane) : rg.lo + lane;
        s_dst[lane] = dst;
      }
    }
    // the plan entries above are static; payloads (and, across GPUs, the counters) come from earlier kernels
    pdl_wait_predecessors();
    const PmPeerFlag pflag = pm_merge_begin<PEER>(a, rep);
    if (DD) {
      // fused denoise epilogue: warp <-> window frame, both outputs from (x, ec, eu, noise)
      const bool has_eu = a.pay[2] != nullptr, has_nz = a.pay[3] != nullptr;
      constexpr int U = 2;
      for (int f0 = warp * U; f0 < f_len; f0 += kWarps * U) {
        T x[U][4], ec[U][4], eu[U][4], nz[U][4];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int fo = f0 + u;
          int f;
          live[u] = fo < f_len && have && pm_frame_live(a, d_ftab, rg, fo, f_len, f);
          if (live[u]) {
            const int o = fo * (HWC > 0 ? HWC : a.hw) + off;          // HWC: (frame, plane) offsets fold into immediates
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              x[u][p] = ldg(static_cast<const T*>(a.pay[0]) + o + p * plane_stride);
              ec[u][p] = ldg(static_cast<const T*>(a.pay[1]) + o + p * plane_stride);
              if (has_eu) eu[u][p] = ldg(static_cast<const T*>(a.pay[2]) + o + p * plane_stride);
              if (has_nz) nz[u][p] = ldg(static_cast<const T*>(a.pay[3]) + o + p * plane_stride);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (live[u]) {
            float xp[4], x0[4];
            pm_ddim<T>(a.dd, x[u], ec[u], eu[u], nz[u], has_eu, has_nz, xp, x0);
            if (a.dd.prev != nullptr)
              pm_prev_mix<T>(a.dd, (f0 + u) * (HWC > 0 ? HWC : a.hw) + off, off, plane_stride, HWC > 0 ? HWC : a.hw, xp);
            tile[(0 * kPmET + lane) * SA + f0 + u] = Atom<T>::pack(xp);
            tile[(1 * kPmET + lane) * SA + f0 + u] = Atom<T>::pack(x0);
          }
      }
    } else if (HWC > 0 && FL == 16 && a.Q == 2) {
      // two payloads x 16 frames = 32 (payload, frame) pairs = 4 per warp: payload q = warp / 4, frames
      // (warp % 4) * 4 + u.  One base pointer per thread; (frame, plane) offsets are immediates.
      constexpr int U = 4;
      const int q = warp >> 2, fo0 = (warp & 3) * U;
      const T* __restrict__ b = static_cast<const T*>(q == 0 ? a.pay[0] : a.pay[1]) + off + fo0 * HWC;
      T e[U][4];
      bool live[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        int f;
        live[u] = have && pm_frame_live(a, d_ftab, rg, fo0 + u, FL, f);
        if (live[u]) {
#pragma unroll
          for (int p = 0; p < 4; ++p) e[u][p] = ldg(b + (u + p * FL) * HWC);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (live[u]) tile[(q * kPmET + lane) * SA + fo0 + u] = atom_join<T>(e[u]);
    } else {
    constexpr int U = 4;
    for (int p0 = warp * U; p0 < n_pairs; p0 += kWarps * U) {
      T e[U][4];
      bool live[U];
      int q[U], fo[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pr = p0 + u;
        q[u] = pr / f_len;
        fo[u] = pr - q[u] * f_len;
        int f;
        live[u] = pr < n_pairs && have && pm_frame_live(a, d_ftab, rg, fo[u], f_len, f);
        if (live[u]) {
          const T* __restrict__ src = static_cast<const T*>(pm_sel(a.pay, q[u])) + fo[u] * a.hw + off;
#pragma unroll
          for (int p = 0; p < 4; ++p) e[u][p] = ldg(src + p * plane_stride);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (live[u]) tile[(q[u] * kPmET + lane) * SA + fo[u]] = atom_join<T>(e[u]);
    }
    }
    // Exchange mode with whole rows (no frame ring): every staged 256-B row leaves through one TMA
    // bulk copy per replica, so NVLink back-pressure never stalls the SM's load/store pipeline.
    const bool row_bulk = PEER && a.peer->mc_pm[0] == 0 && sizeof(A) == 16 && d_ftab == nullptr && a.Q > 0 && a.mask_mode != DS_MASK_FULL;   // fp32: 16-byte aligned rows
    if (row_bulk) fence_proxy_async_smem();
    pm_merge_wait<PEER>(a, pflag);
    __syncthreads();

    if (a.mask_mode == DS_MASK_PLANE && static_cast<int>(threadIdx.x) < cnt) pm_store_mask_plane<T, PEER>(a, rep, s_dst[threadIdx.x]);
    if (row_bulk) {
      const uint32_t row_bytes = static_cast<uint32_t>(f_len * sizeof(A));
      const int n_rows = a.Q * cnt;
      bool issued = false;
      for (int u = threadIdx.x; u < n_rows * rep.world; u += kBlock) {
        const int r = u / n_rows, qi = u - r * n_rows;
        const int q = qi / cnt, i = qi - q * cnt;
        bulk_store(reinterpret_cast<A*>(rep.ptr[r][q]) + static_cast<size_t>(s_dst[i]) * a.F,
                   tile + (q * kPmET + i) * SA, row_bytes);
        issued = true;
      }
      if (issued) bulk_commit_and_wait_read();
    } else {
      const int per_q = cnt * f_len;
      const int n_units = (a.Q > 0 ? a.Q : 1) * per_q;
      for (int u = threadIdx.x; u < n_units; u += kBlock) {
        const int q = u / per_q, r = u - q * per_q;
        const int i = r / f_len, fo = r - i * f_len;
        int f;
        if (pm_frame_live(a, d_ftab, rg, fo, f_len, f)) {
          const size_t o = static_cast<size_t>(s_dst[i]) * a.F + f;
          if (a.Q > 0) pm_store<A, PEER>(a, rep, q, o, tile[(q * kPmET + i) * SA + fo]);
          if (q == 0 && a.mask_mode == DS_MASK_FULL) pm_store_mask_atom<A, PEER>(a, rep, o, atom_ones<T>());
        }
      }
    }
  }
  pm_merge_end<PEER>(a);
}

// Tiled weighted merge (fast path): the host pads the CSR entry list so that tile t owns
// exactly entries [t*kPmEC, (t+1)*kPmEC) (d_ent_k < 0 = padding), all of one view, and <= 32
// whole rows: d_tile[t] = (first row, view, rows, entries).  The entry loads therefore depend
// on nothing but blockIdx: global dependency depth 2 (entries -> payload) instead of 3, and
// every loop bound of phase A is a compile-time constant.  d_view_ptr[v] = first tile of view
// v; d_row_start[r] = first (padded) entry of row r, rows of a tile being contiguous.
// HWC: view pixels per window known at compile time (40 x 64 = 2560, the one window size of the reference's
// pipelines: 320 x 512 pixels / 8), else 0 = runtime.  With HWC every (frame, plane) offset of a payload element is
// an immediate of the load instruction, so the 32 loads of a thread need 4 address computations instead of 32.
template <typename T, int FL, int QN, bool DD, bool PEER, int HWC = 0>
__global__ void __launch_bounds__(kBlock)
pm_merge_weighted_tiled_kernel(const PmMergeArgs a, const int32_t* __restrict__ d_view_ptr,
                               const int4* __restrict__ d_tile, const int32_t* __restrict__ d_row_dst,
                               const int32_t* __restrict__ d_row_start, const int32_t* __restrict__ d_ent_k,
                               const float* __restrict__ d_ent_w, const int32_t* __restrict__ d_ftab,
                               const int32_t* __restrict__ d_fbegin) {
  using A = typename Atom<T>::type;
  static_assert(FL == 16 && kBlock == 256, "lane mapping: 8 warps x 2 frames in phase A, 16 rows x 16 frames in phase B");
  constexpr int SA = FL + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* prod = reinterpret_cast<float4*>(smem_raw);                 // [QN][EC][SA] fp32 products
  float* s_w = reinterpret_cast<float*>(prod + QN * kPmEC * SA);      // [EC]
  int32_t* s_rp = reinterpret_cast<int32_t*>(s_w + kPmEC);            // [TR + 1] local entry offsets
  int32_t* s_dst = s_rp + kPmTR + 1;                                  // [TR]

  // ---- which tile / job ----
  const int s = blockIdx.y;
  const int v_lo = min(max(a.job_begin - s * a.vpf, 0), a.vpf);
  const int v_hi = min(max(a.job_begin + a.job_count - s * a.vpf, 0), a.vpf);
  const int t = __ldg(d_view_ptr + v_lo) + blockIdx.x;
  pdl_launch_dependents();
  if (t >= __ldg(d_view_ptr + v_hi)) {
    pm_merge_end<PEER>(a);                         // a CTA without a tile still arrives for the publication
    return;
  }
  __shared__ PmReplicas rep;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e_base = t * kPmEC;

  // ---- level-1 loads: nothing below depends on another global load ----
  int k[2];
  float w[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    k[h] = __ldg(d_ent_k + e_base + h * 32 + lane);
    w[h] = __ldg(d_ent_w + e_base + h * 32 + lane);
  }
  const int4 td = __ldg(d_tile + t);
  const int r0 = td.x, view = td.y, cnt = td.z;
  const int fb = d_fbegin != nullptr ? __ldg(d_fbegin + s) % a.F : 0;
  bool liveA[2], liveB;
  int fB;
  {
    const int fo_b = threadIdx.x & (FL - 1);
    fB = fb + fo_b;
    fB = fB >= a.F ? fB - a.F : fB;
    liveB = d_ftab == nullptr || __ldg(d_ftab + fB) == s * FL + fo_b;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int fo = warp + j * kWarps;
      int f = fb + fo;
      f = f >= a.F ? f - a.F : f;
      liveA[j] = d_ftab == nullptr || __ldg(d_ftab + f) == s * FL + fo;
    }
  }
  if (threadIdx.x < cnt) {
    s_rp[threadIdx.x] = __ldg(d_row_start + r0 + threadIdx.x) - e_base;
    s_dst[threadIdx.x] = __ldg(d_row_dst + r0 + threadIdx.x);
  } else if (threadIdx.x == cnt) {
    s_rp[cnt] = td.w;
  }
  if (warp == 0) { s_w[lane] = w[0]; s_w[32 + lane] = w[1]; }

  // the plan data above is static; payloads (and, across GPUs, the counters) come from earlier kernels
  pdl_wait_predecessors();
  const PmPeerFlag pflag = pm_merge_begin<PEER>(a, rep);

  // ---- phase A: products (level-2 loads) ----
  const int plane_stride = FL * (HWC > 0 ? HWC : a.hw);
  const int job_off = (s * a.vpf + view - a.job_begin) * kPmPlanes * plane_stride;
  if (DD) {
    static_assert(!DD || QN == 2, "the fused denoise epilogue produces exactly x_prev and pred_x0");
    const bool has_eu = a.pay[2] != nullptr, has_nz = a.pay[3] != nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      T x[2][4], ec[2][4], eu[2][4], nz[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (k[h] >= 0 && liveA[j]) {
          const int o = job_off + (warp + j * kWarps) * (HWC > 0 ? HWC : a.hw) + k[h];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            x[j][p] = ldg(static_cast<const T*>(a.pay[0]) + o + p * plane_stride);
            ec[j][p] = ldg(static_cast<const T*>(a.pay[1]) + o + p * plane_stride);
            if (has_eu) eu[j][p] = ldg(static_cast<const T*>(a.pay[2]) + o + p * plane_stride);
            if (has_nz) nz[j][p] = ldg(static_cast<const T*>(a.pay[3]) + o + p * plane_stride);
          }
        }
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (k[h] >= 0 && liveA[j]) {
          float xp[4], x0[4];
          pm_ddim<T>(a.dd, x[j], ec[j], eu[j], nz[j], has_eu, has_nz, xp, x0);
          if (a.dd.prev != nullptr)
            pm_prev_mix<T>(a.dd, job_off + (warp + j * kWarps) * (HWC > 0 ? HWC : a.hw) + k[h], job_off + k[h], plane_stride,
                           HWC > 0 ? HWC : a.hw, xp);
          // the reference merges the ROUNDED payload tensors: round to T before the weighting
          const int e = h * 32 + lane, fo = warp + j * kWarps;
          prod[(0 * kPmEC + e) * SA + fo] = make_float4(
              __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(xp[0])), w[h]), __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(xp[1])), w[h]),
              __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(xp[2])), w[h]), __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(xp[3])), w[h]));
          prod[(1 * kPmEC + e) * SA + fo] = make_float4(
              __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(x0[0])), w[h]), __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(x0[1])), w[h]),
              __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(x0[2])), w[h]), __fmul_rn(Elem<T>::to_f(Elem<T>::from_f(x0[3])), w[h]));
        }
    }
  } else {
  // HWC path: branch-free, one base pointer per (half, payload) and immediates for everything else.  Padding entries
  // (k < 0, w = 0) read view pixel 0 instead of being predicated off: their products are never consumed (phase B
  // only walks the real entries of a row); dead window frames of a frame ring are skipped per warp (uniform).
  const unsigned koff[2] = {static_cast<unsigned>(max(k[0], 0)), static_cast<unsigned>(max(k[1], 0))};
  T x[2][QN][2][4];                               // [half][payload][frame j][plane]
  if (HWC > 0) {
    // one base pointer per (half, payload): (job, plane 0, frame `warp`, pixel k); the rest are immediates
#pragma unroll
    for (int q = 0; q < QN; ++q)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const T* __restrict__ b = static_cast<const T*>(a.pay[q]) + job_off + warp * HWC + koff[h];
#pragma unroll
        for (int j = 0; j < 2; ++j)
          if (liveA[j]) {
#pragma unroll
            for (int p = 0; p < 4; ++p) x[h][q][j][p] = ldg(b + (j * kWarps + p * FL) * HWC);
          }
      }
  } else {
    // generic window size: predicated per entry (measured faster than the branch-free form at runtime strides)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int q = 0; q < QN; ++q)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (k[h] >= 0 && liveA[j]) {
            const T* __restrict__ src = static_cast<const T*>(a.pay[q]) + job_off + (warp + j * kWarps) * a.hw + k[h];
#pragma unroll
            for (int p = 0; p < 4; ++p) x[h][q][j][p] = ldg(src + p * plane_stride);
          }
        }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int q = 0; q < QN; ++q)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if ((HWC > 0 || k[h] >= 0) && liveA[j])
          prod[(q * kPmEC + h * 32 + lane) * SA + warp + j * kWarps] =
              make_float4(__fmul_rn(Elem<T>::to_f(x[h][q][j][0]), w[h]), __fmul_rn(Elem<T>::to_f(x[h][q][j][1]), w[h]),
                          __fmul_rn(Elem<T>::to_f(x[h][q][j][2]), w[h]), __fmul_rn(Elem<T>::to_f(x[h][q][j][3]), w[h]));
  }
  pm_merge_wait<PEER>(a, pflag);
  __syncthreads();

  // ---- phase B: ordered accumulation, normalise, full-row stores ----
  // Exchange mode with whole rows (no frame ring): results are staged in shared memory (over the
  // consumed products) and every 256-B row leaves through one TMA bulk copy per replica.
  const bool row_bulk = PEER && a.peer->mc_pm[0] == 0 && sizeof(A) == 16 && d_ftab == nullptr && a.mask_mode != DS_MASK_FULL;   // fp32: 16-byte aligned rows
  const int fo_b = threadIdx.x & (FL - 1), row_b = threadIdx.x / FL;
  constexpr int NSW = kPmTR / (kBlock / FL);
  float4 res[NSW][QN];
  bool ok[NSW];
#pragma unroll
  for (int sw = 0; sw < NSW; ++sw) {
    const int i = sw * (kBlock / FL) + row_b;
    ok[sw] = false;
    if (i >= cnt) continue;
    const int lo = s_rp[i], hi = s_rp[i + 1];
    float ws = 0.f;
    float acc[QN][4];
#pragma unroll
    for (int q = 0; q < QN; ++q)
#pragma unroll
      for (int p = 0; p < 4; ++p) acc[q][p] = 0.f;
    for (int e = lo; e < hi; ++e) {
      ws = __fadd_rn(ws, s_w[e]);
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        const float4 pr = prod[(q * kPmEC + e) * SA + fo_b];
        acc[q][0] = __fadd_rn(acc[q][0], pr.x);
        acc[q][1] = __fadd_rn(acc[q][1], pr.y);
        acc[q][2] = __fadd_rn(acc[q][2], pr.z);
        acc[q][3] = __fadd_rn(acc[q][3], pr.w);
      }
    }
    if (!(ws > 0.f)) continue;
    ok[sw] = true;
#pragma unroll
    for (int q = 0; q < QN; ++q)
      res[sw][q] = make_float4(__fdiv_rn(acc[q][0], ws), __fdiv_rn(acc[q][1], ws), __fdiv_rn(acc[q][2], ws),
                               __fdiv_rn(acc[q][3], ws));
    const int dst = s_dst[i];
    if (fo_b == 0 && a.mask_mode == DS_MASK_PLANE) pm_store_mask_plane<T, PEER>(a, rep, dst);
    if (!row_bulk && liveB) {
      const size_t o = static_cast<size_t>(dst) * a.F + fB;
#pragma unroll
      for (int q = 0; q < QN; ++q) {
        const float r[4] = {res[sw][q].x, res[sw][q].y, res[sw][q].z, res[sw][q].w};
        pm_store<A, PEER>(a, rep, q, o, Atom<T>::pack(r));
      }
      if (a.mask_mode == DS_MASK_FULL) pm_store_mask_atom<A, PEER>(a, rep, o, atom_ones<T>());
    }
  }
  if (row_bulk) {
    __syncthreads();                               // every product has been consumed
    A* rows = reinterpret_cast<A*>(prod);          // [QN][TR][SA] result atoms (TR <= EC: fits over the products)
    int32_t* s_ok = s_rp;                          // row offsets are no longer needed
#pragma unroll
    for (int sw = 0; sw < NSW; ++sw) {
      const int i = sw * (kBlock / FL) + row_b;
      if (i < cnt && fo_b == 0) s_ok[i] = ok[sw] ? 1 : 0;
      if (ok[sw]) {
#pragma unroll
        for (int q = 0; q < QN; ++q) {
          const float r[4] = {res[sw][q].x, res[sw][q].y, res[sw][q].z, res[sw][q].w};
          rows[(q * kPmTR + i) * SA + fo_b] = Atom<T>::pack(r);
        }
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    const uint32_t row_bytes = static_cast<uint32_t>(FL * sizeof(A));
    const int n_rows = QN * cnt;
    bool issued = false;
    for (int u = threadIdx.x; u < n_rows * rep.world; u += kBlock) {
      const int r = u / n_rows, qi = u - r * n_rows;
      const int q = qi / cnt, i = qi - q * cnt;
      if (s_ok[i]) {
        bulk_store(reinterpret_cast<A*>(rep.ptr[r][q]) + static_cast<size_t>(s_dst[i]) * a.F,
                   rows + (q * kPmTR + i) * SA, row_bytes);
        issued = true;
      }
    }
    if (issued) bulk_commit_and_wait_read();
  }
  pm_merge_end<PEER>(a);
}

// Generic (runtime f_len, any row length) weighted merge.  CTA = 32 panorama rows; each lane
// owns a row and walks its CSR entries (tap-major, k ascending -- index_add_'s order):
//   acc = acc + (payload * w);  ws = ws + w;   row = acc / ws     (each op rounded on its own)
template <typename T, bool PEER>
__global__ void __launch_bounds__(kBlock)
pm_merge_bilinear_kernel(const PmMergeArgs a, const int32_t* __restrict__ d_view_ptr,
                         const int32_t* __restrict__ d_row_dst, const int32_t* __restrict__ d_row_ptr,
                         const int32_t* __restrict__ d_ent_src, const float* __restrict__ d_ent_w,
                         const int32_t* __restrict__ d_ftab, const int32_t* __restrict__ d_fbegin) {
  using A = typename Atom<T>::type;
  pdl_launch_dependents();
  __shared__ PmReplicas rep;
  PmRange rg;
  if (pm_range(a, d_view_ptr, d_fbegin, rg)) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int SA = a.f_len | 1;
    A* tile = reinterpret_cast<A*>(smem_raw);                          // [Q][ET][SA]
    int32_t* s_dst = reinterpret_cast<int32_t*>(tile + (a.Q > 0 ? a.Q : 1) * kPmET * SA);   // -1: row not written

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int plane_stride = a.f_len * a.hw;
    const int n_pairs = a.Q * a.f_len;
    const int cnt = rg.hi - rg.lo;
    const bool have = lane < cnt;
    int e0 = 0, e1 = 0;
    if (have) {
      e0 = __ldg(d_row_ptr + rg.lo + lane);
      e1 = __ldg(d_row_ptr + rg.lo + lane + 1);
    }
    float ws = 0.f;
    for (int e = e0; e < e1; ++e) ws = __fadd_rn(ws, __ldg(d_ent_w + e));
    const bool wr = have && ws > 0.f;
    if (warp == 0) {
      const int dst = wr ? __ldg(d_row_dst + rg.lo + lane) : -1;
      s_dst[lane] = dst;
    }
    pdl_wait_predecessors();                       // plan data above is static, payloads are not
    const PmPeerFlag pflag = pm_merge_begin<PEER>(a, rep);
    for (int pr = warp; pr < n_pairs; pr += kWarps) {
      const int q = pr / a.f_len, fo = pr - q * a.f_len;
      int f;
      if (!wr || !pm_frame_live(a, d_ftab, rg, fo, a.f_len, f)) continue;
      const T* __restrict__ src = static_cast<const T*>(pm_sel(a.pay, q)) + fo * a.hw;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int e = e0; e < e1; ++e) {
        const int off = pm_job_offset(a, rg.s, __ldg(d_ent_src + e), plane_stride);
        const float w = __ldg(d_ent_w + e);
#pragma unroll
        for (int p = 0; p < 4; ++p)
          acc[p] = __fadd_rn(acc[p], __fmul_rn(Elem<T>::to_f(ldg(src + off + p * plane_stride)), w));
      }
      float r[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) r[p] = __fdiv_rn(acc[p], ws);
      tile[(q * kPmET + lane) * SA + fo] = Atom<T>::pack(r);
    }
    pm_merge_wait<PEER>(a, pflag);
    __syncthreads();
    if (a.mask_mode == DS_MASK_PLANE && static_cast<int>(threadIdx.x) < cnt && s_dst[threadIdx.x] >= 0)
      pm_store_mask_plane<T, PEER>(a, rep, s_dst[threadIdx.x]);
    const int per_q = cnt * a.f_len;
    const int n_units = (a.Q > 0 ? a.Q : 1) * per_q;
    for (int u = threadIdx.x; u < n_units; u += kBlock) {
      const int q = u / per_q, r = u - q * per_q;
      const int i = r / a.f_len, fo = r - i * a.f_len;
      int f;
      if (s_dst[i] >= 0 && pm_frame_live(a, d_ftab, rg, fo, a.f_len, f)) {
        const size_t o = static_cast<size_t>(s_dst[i]) * a.F + f;
        if (a.Q > 0) pm_store<A, PEER>(a, rep, q, o, tile[(q * kPmET + i) * SA + fo]);
        if (q == 0 && a.mask_mode == DS_MASK_FULL) pm_store_mask_atom<A, PEER>(a, rep, o, atom_ones<T>());
      }
    }
  }
  pm_merge_end<PEER>(a);
}

static bool pm_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int pm_fill_merge_args(PmMergeArgs& a, void* const* h_pms, const void* const* h_payloads, int Q,
                              void* mask, int mask_mode, int n_ent, const int32_t* d_ftab,
                              const int32_t* d_fbegin, int n_fwin, int vpf, int job_begin, int job_count,
                              int P, int F, int HW, int f_len, int hw, void* d_peer) {
  DS_REQUIRE(Q >= 0 && Q <= DS_MAX_PAYLOADS, DS_ERR_COUNT);
  DS_REQUIRE(Q == 0 || (h_pms != nullptr && h_payloads != nullptr), DS_ERR_NULL);
  DS_REQUIRE(P == kPmPlanes, DS_ERR_DIM);
  DS_REQUIRE(n_ent >= 0 && n_fwin > 0 && vpf > 0 && F > 0 && HW > 0 && f_len > 0 && hw > 0, DS_ERR_DIM);
  DS_REQUIRE(f_len <= F && n_fwin <= 65535, DS_ERR_DIM);
  DS_REQUIRE(mask_mode == DS_MASK_NONE || mask_mode == DS_MASK_PLANE || mask_mode == DS_MASK_FULL, DS_ERR_DIM);
  DS_REQUIRE(mask_mode == DS_MASK_NONE || mask != nullptr, DS_ERR_NULL);
  DS_REQUIRE(Q > 0 || mask_mode != DS_MASK_NONE, DS_ERR_COUNT);
  DS_REQUIRE(job_begin >= 0 && job_begin < (1 << 30) && job_count >= 0 && job_count <= (1 << 30), DS_ERR_DIM);
  DS_REQUIRE(job_count > 0 || d_peer != nullptr, DS_ERR_DIM);   // an idle rank still takes part in the exchange
  // payload offsets are 32-bit inside the kernels
  const long long all_jobs = static_cast<long long>(n_fwin) * vpf;
  const long long jobs_here = job_count < all_jobs ? job_count : all_jobs;
  DS_REQUIRE(jobs_here * P * f_len * hw < (1ll << 31), DS_ERR_DIM);
  // frame ring: both tables or neither; without them the single frame window is the identity
  DS_REQUIRE((d_ftab == nullptr) == (d_fbegin == nullptr), DS_ERR_NULL);
  DS_REQUIRE(d_ftab != nullptr || (f_len == F && n_fwin == 1), DS_ERR_DIM);
  DS_REQUIRE(mask_mode != DS_MASK_PLANE || d_ftab == nullptr, DS_ERR_DIM);
  for (int q = 0; q < DS_MAX_PAYLOADS; ++q) {
    a.pm[q] = q < Q ? h_pms[q] : nullptr;
    a.pay[q] = q < Q ? h_payloads[q] : nullptr;
    if (q < Q) DS_REQUIRE(a.pm[q] != nullptr && (a.pay[q] != nullptr || job_count == 0), DS_ERR_NULL);
    if (q < Q) DS_REQUIRE(pm_aligned16(a.pm[q]), DS_ERR_ALIGN);
  }
  if (mask_mode == DS_MASK_FULL) DS_REQUIRE(pm_aligned16(mask), DS_ERR_ALIGN);
  a.mask = mask; a.mask_mode = mask_mode;
  a.Q = Q; a.n_ent = n_ent;
  a.vpf = vpf; a.job_begin = job_begin; a.job_count = job_count;
  a.F = F; a.HW = HW; a.f_len = f_len; a.hw = hw;
  a.peer = static_cast<ds_peer_table_t*>(d_peer);
  a.publish = d_peer != nullptr ? g_merge_publish : 0;
  a.peer_timeout_ns = g_peer_timeout_ns;
  return DS_OK;
}

// ctas_per_sm: 0 = occupancy limit, > 0 = cap (persistent kernels); < 0 = one CTA per item.
template <typename K, typename... Args>
static int pm_launch(K kernel, int pdl_bit, long long items, int ctas_per_sm, unsigned grid_y, size_t smem, cudaStream_t st, Args... args) {
  DS_REQUIRE(smem <= 200 * 1024, DS_ERR_DIM);
  if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  long long gx = items;
  if (ctas_per_sm >= 0) {
    int occ = 0, dev = 0, sms = kSMs;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, smem) != cudaSuccess || occ < 1) occ = 1;
    if (ctas_per_sm > 0 && ctas_per_sm < occ) occ = ctas_per_sm;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long slots = static_cast<long long>(sms) * occ;
    gx = items < slots ? items : slots;
  }
  if (gx < 1) gx = 1;
  DS_REQUIRE(gx < (1ll << 31), DS_ERR_DIM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(gx), grid_y);
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = (g_pdl & pdl_bit) ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  return e == cudaSuccess ? check_launch() : static_cast<int>(e);
}

}  // namespace ds

extern "C" int ds_pm_gather(const void* pm, void* out, const int32_t* d_idx, const float* d_wt, int taps,
                            const int32_t* d_jobs, int h_f0, int n_jobs, int P, int F, int HW, int f_len,
                            int hw, void* d_peer, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(pm != nullptr && (out != nullptr || n_jobs == 0), DS_ERR_NULL);
  DS_REQUIRE(taps == 1 || taps == 4, DS_ERR_DIM);
  DS_REQUIRE(taps == 1 || (d_idx != nullptr && d_wt != nullptr), DS_ERR_NULL);
  DS_REQUIRE(P == kPmPlanes, DS_ERR_DIM);
  DS_REQUIRE(n_jobs >= 0 && F > 0 && HW > 0 && f_len > 0 && f_len <= F && hw > 0 && h_f0 >= 0, DS_ERR_DIM);
  DS_REQUIRE(n_jobs > 0 || d_peer != nullptr, DS_ERR_DIM);        // an idle rank still takes part in the exchange
  DS_REQUIRE(d_idx != nullptr || hw <= HW, DS_ERR_DIM);
  DS_REQUIRE(pm_aligned16(pm) && pm_aligned16(out), DS_ERR_ALIGN);
  const int tpj = static_cast<int>(div_up(hw, kPmKT));
  const long long n_tiles = static_cast<long long>(n_jobs) * tpj;
  DS_REQUIRE(n_tiles < (1ll << 31) && static_cast<long long>(n_jobs) * P * f_len * hw < (1ll << 40), DS_ERR_DIM);
  DS_REQUIRE(static_cast<long long>(n_jobs) * taps * hw < (1ll << 31), DS_ERR_DIM);
  PmGatherArgs a{F, HW, f_len, hw, static_cast<int>(n_tiles), tpj, d_peer != nullptr ? g_merge_publish : 0};
  const bool pipe = (g_gather_pipe & (taps == 1 ? 1 : 2)) != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DS_DISPATCH_DTYPE(dtype, {
    constexpr int VE = 16 / sizeof(ET);
    const size_t smem = static_cast<size_t>(kPmPlanes) * f_len * (kPmKT + VE) * sizeof(ET) +
                        static_cast<size_t>(taps) * kPmKT * 8 * (pipe ? 2 : 1);
    using AT = typename Atom<ET>::type;
    const AT* pmp = static_cast<const AT*>(pm);
    ET* op = static_cast<ET*>(out);
#define DS_PM_GATHER(TAPS_, FL_, MINB_, PIPE_, PDL_)                                                              \
  return pm_launch(pm_gather_kernel<ET, TAPS_, FL_, MINB_, PIPE_>, PDL_, n_tiles, g_gather_ctas_per_sm, 1, smem, st, pmp, \
                   op, d_idx, d_wt, d_jobs, h_f0, a, static_cast<ds_peer_table_t*>(d_peer), g_peer_timeout_ns)
#define DS_PM_GATHER_HW(TAPS_, MINB_, PIPE_, PDL_)                                                                \
  return pm_launch(pm_gather_kernel<ET, TAPS_, 16, MINB_, PIPE_, 2560>, PDL_, n_tiles, g_gather_ctas_per_sm, 1, smem, st, \
                   pmp, op, d_idx, d_wt, d_jobs, h_f0, a, static_cast<ds_peer_table_t*>(d_peer), g_peer_timeout_ns)
    // 40 x 64 windows of 16 frames: the 4-tap gather gains 9 % from the specialisation (13.4 vs 14.7 us at cfg2); the
    // copy-only gather does not (it spills at its 32-register budget) and keeps the generic kernel
    const bool hwc = g_hw_const && hw == 2560 && f_len == 16 && d_idx != nullptr;
    if (hwc && taps == 4) { if (pipe) DS_PM_GATHER_HW(4, 4, true, 2); DS_PM_GATHER_HW(4, 4, false, 2); }
    if (taps == 1) {
      if (f_len == 16) { if (pipe) DS_PM_GATHER(1, 16, 8, true, 1); DS_PM_GATHER(1, 16, 8, false, 1); }
      if (pipe) DS_PM_GATHER(1, 0, 4, true, 1);
      DS_PM_GATHER(1, 0, 4, false, 1);
    }
    if (f_len == 16) { if (pipe) DS_PM_GATHER(4, 16, 4, true, 2); DS_PM_GATHER(4, 16, 4, false, 2); }
    if (pipe) DS_PM_GATHER(4, 0, 4, true, 2);
    DS_PM_GATHER(4, 0, 4, false, 2);
  });
  return DS_OK;
}

extern "C" int ds_pm_gather_renoise(const void* pm, void* out, void* out_raw, void* out_mask, const void* eps,
                                    const void* mask, int mask_mode, const int32_t* d_idx, const int32_t* d_midx,
                                    const float* h_coef, const int32_t* d_jobs, int h_f0, int n_jobs, int P, int F,
                                    int HW, int f_len, int hw, void* d_peer, int dtype, void* stream) {
  using namespace ds;
  DS_REQUIRE(pm != nullptr && eps != nullptr && mask != nullptr && d_idx != nullptr && d_midx != nullptr &&
             h_coef != nullptr && (out != nullptr || n_jobs == 0), DS_ERR_NULL);
  DS_REQUIRE(mask_mode == DS_MASK_PLANE || mask_mode == DS_MASK_FULL, DS_ERR_DIM);
  DS_REQUIRE(P == kPmPlanes, DS_ERR_DIM);
  DS_REQUIRE(n_jobs >= 0 && F > 0 && HW > 0 && f_len > 0 && f_len <= F && hw > 0 && h_f0 >= 0, DS_ERR_DIM);
  DS_REQUIRE(n_jobs > 0 || d_peer != nullptr, DS_ERR_DIM);
  DS_REQUIRE(pm_aligned16(pm) && (mask_mode != DS_MASK_FULL || pm_aligned16(mask)), DS_ERR_ALIGN);
  const int tpj = static_cast<int>(div_up(hw, kPmKT));
  const long long n_tiles = static_cast<long long>(n_jobs) * tpj;
  DS_REQUIRE(n_tiles < (1ll << 31) && static_cast<long long>(n_jobs) * P * f_len * hw < (1ll << 40), DS_ERR_DIM);
  PmGatherArgs a{F, HW, f_len, hw, static_cast<int>(n_tiles), tpj, d_peer != nullptr ? g_merge_publish : 0};
  PmRenoiseArgs e{eps, mask, out_raw, out_mask, h_coef[0], h_coef[1], h_coef[2], h_coef[3]};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DS_DISPATCH_DTYPE(dtype, {
    constexpr int VE = 16 / sizeof(ET);
    const size_t smem = static_cast<size_t>(kPmPlanes) * f_len * (kPmKT + VE) * sizeof(ET) + 2 * kPmKT * sizeof(int32_t);
    using AT = typename Atom<ET>::type;
    if (mask_mode == DS_MASK_PLANE)
      return pm_launch(pm_gather_renoise_kernel<ET, DS_MASK_PLANE>, 0, n_tiles, 0, 1, smem, st, static_cast<const AT*>(pm),
                       static_cast<ET*>(out), d_idx, d_midx, d_jobs, h_f0, a, e, static_cast<ds_peer_table_t*>(d_peer),
                       g_peer_timeout_ns);
    return pm_launch(pm_gather_renoise_kernel<ET, DS_MASK_FULL>, 0, n_tiles, 0, 1, smem, st, static_cast<const AT*>(pm),
                     static_cast<ET*>(out), d_idx, d_midx, d_jobs, h_f0, a, e, static_cast<ds_peer_table_t*>(d_peer),
                     g_peer_timeout_ns);
  });
  return DS_OK;
}

extern "C" int ds_pm_merge_nearest(void* const* h_pms, const void* const* h_payloads, int Q,
                                   void* mask, int mask_mode, const int32_t* d_view_ptr,
                                   const int32_t* d_ent_src, const int32_t* d_ent_dst, int n_ent,
                                   const int32_t* d_ftab, const int32_t* d_fbegin, int n_fwin,
                                   int views_per_fwin, int job_begin, int job_count,
                                   int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype, void* stream) {
  using namespace ds;
  PmMergeArgs a;
  int rc = pm_fill_merge_args(a, h_pms, h_payloads, Q, mask, mask_mode, n_ent, d_ftab, d_fbegin, n_fwin,
                              views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw, d_peer);
  if (rc != DS_OK) return rc;
  DS_REQUIRE((d_view_ptr == nullptr) == (d_ent_src == nullptr) && (d_view_ptr == nullptr) == (d_ent_dst == nullptr), DS_ERR_NULL);
  if (d_view_ptr == nullptr)                     // identity plan: entry i = (k = i, dst = i) of job 0
    DS_REQUIRE(n_ent <= HW && n_ent <= hw && views_per_fwin == 1 && n_fwin == 1 && job_begin == 0, DS_ERR_DIM);
  if (n_ent == 0 && d_peer == nullptr) return DS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long items = div_up(n_ent, kPmET);
  DS_DISPATCH_DTYPE(dtype, {
    using AT = typename Atom<ET>::type;
    const size_t smem = static_cast<size_t>(Q > 0 ? Q : 1) * kPmET * (f_len | 1) * sizeof(AT) + kPmET * sizeof(int32_t);
    if (f_len == 16) {
      if (d_peer != nullptr) return pm_launch(pm_merge_nearest_kernel<ET, 16, false, true>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin);
      // (38 registers: 6 CTAs fit an SM instead of 5.  Holding it at 5 by padding the shared-memory request was measured
      // and is far WORSE -- 24.7 vs 15.2 us -- because the larger carve-out shrinks L1, which the scattered 4-byte payload
      // loads depend on: profiles/r02f_kbench_hw_const.jsonl)
      if (hw == 2560 && Q == 2 && g_hw_const && d_view_ptr != nullptr)
        return pm_launch(pm_merge_nearest_kernel<ET, 16, false, false, 2560>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin);
      return pm_launch(pm_merge_nearest_kernel<ET, 16, false, false>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin);
    }
    { if (d_peer != nullptr) return pm_launch(pm_merge_nearest_kernel<ET, 0, false, true>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin); return pm_launch(pm_merge_nearest_kernel<ET, 0, false, false>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin); }
  });
  return DS_OK;
}

extern "C" int ds_pm_merge_bilinear(void* const* h_pms, const void* const* h_payloads, int Q,
                                    void* mask, int mask_mode, const int32_t* d_view_ptr,
                                    const int32_t* d_row_dst, const int32_t* d_row_ptr, int n_rows,
                                    const int32_t* d_ent_src, const float* d_ent_w,
                                    const int32_t* d_ftab, const int32_t* d_fbegin, int n_fwin,
                                    int views_per_fwin, int job_begin, int job_count,
                                    int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype, void* stream) {
  using namespace ds;
  PmMergeArgs a;
  int rc = pm_fill_merge_args(a, h_pms, h_payloads, Q, mask, mask_mode, n_rows, d_ftab, d_fbegin, n_fwin,
                              views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw, d_peer);
  if (rc != DS_OK) return rc;
  DS_REQUIRE(d_view_ptr != nullptr && d_row_dst != nullptr && d_row_ptr != nullptr && d_ent_src != nullptr && d_ent_w != nullptr, DS_ERR_NULL);
  if (n_rows == 0 && d_peer == nullptr) return DS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long items = div_up(n_rows, kPmET);
  DS_DISPATCH_DTYPE(dtype, {
    const size_t smem = static_cast<size_t>(Q > 0 ? Q : 1) * kPmET * (f_len | 1) * sizeof(typename Atom<ET>::type) + kPmET * sizeof(int32_t);
    { if (d_peer != nullptr) return pm_launch(pm_merge_bilinear_kernel<ET, true>, 8, items, -1, n_fwin, smem, st, a, d_view_ptr, d_row_dst, d_row_ptr, d_ent_src, d_ent_w, d_ftab, d_fbegin); return pm_launch(pm_merge_bilinear_kernel<ET, false>, 8, items, -1, n_fwin, smem, st, a, d_view_ptr, d_row_dst, d_row_ptr, d_ent_src, d_ent_w, d_ftab, d_fbegin); }
  });
  return DS_OK;
}

extern "C" int ds_pm_merge_bilinear_tiled(void* const* h_pms, const void* const* h_payloads, int Q,
                                          void* mask, int mask_mode, const int32_t* d_view_ptr,
                                          const int32_t* d_tile, int n_tiles, const int32_t* d_row_dst,
                                          const int32_t* d_row_start, int n_rows, const int32_t* d_ent_k,
                                          const float* d_ent_w, const int32_t* d_ftab, const int32_t* d_fbegin,
                                          int n_fwin, int views_per_fwin, int job_begin, int job_count,
                                          int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype, void* stream) {
  using namespace ds;
  PmMergeArgs a;
  int rc = pm_fill_merge_args(a, h_pms, h_payloads, Q, mask, mask_mode, n_rows, d_ftab, d_fbegin, n_fwin,
                              views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw, d_peer);
  if (rc != DS_OK) return rc;
  DS_REQUIRE(d_view_ptr != nullptr && d_tile != nullptr && d_row_dst != nullptr && d_row_start != nullptr &&
             d_ent_k != nullptr && d_ent_w != nullptr, DS_ERR_NULL);
  DS_REQUIRE(pm_aligned16(d_tile), DS_ERR_ALIGN);
  const int4* d_tile4 = reinterpret_cast<const int4*>(d_tile);
  DS_REQUIRE(f_len == 16 && (Q == 1 || Q == 2 || Q == 4) && n_tiles >= 0, DS_ERR_DIM);   // other shapes: ds_pm_merge_bilinear
  if (n_tiles == 0 && d_peer == nullptr) return DS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(Q) * kPmEC * 17 * sizeof(float4) + kPmEC * sizeof(float) + (2 * kPmTR + 1) * sizeof(int32_t);
  DS_DISPATCH_DTYPE(dtype, {
    if (Q == 1) { if (d_peer != nullptr) return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 1, false, true>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 1, false, false>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); }
    if (Q == 2) { if (d_peer != nullptr) return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 2, false, true>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); if (hw == 2560 && g_hw_const) return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 2, false, false, 2560>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 2, false, false>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); }
    { if (d_peer != nullptr) return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 4, false, true>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 4, false, false>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); }
  });
  return DS_OK;
}

// Shared argument handling of the two fused-denoise entry points: panoramas 0 / 1 receive
// x_prev / pred_x0; payload slots carry x, eps_cond, eps_uncond (or NULL), noise (or NULL).
static int pm_fill_ddim_args(ds::PmMergeArgs& a, void* const* h_pms, const void* x, const void* eps_cond,
                             const void* eps_uncond, const void* noise, const void* prev, const void* mask_view,
                             int mask_view_full, const float* h_coef, void* mask, int mask_mode,
                             int n_ent, const int32_t* d_ftab, const int32_t* d_fbegin, int n_fwin, int vpf,
                             int job_begin, int job_count, int P, int F, int HW, int f_len, int hw, void* d_peer) {
  using namespace ds;
  DS_REQUIRE(h_coef != nullptr && x != nullptr && eps_cond != nullptr, DS_ERR_NULL);
  const void* two[2] = {x, eps_cond};
  int rc = pm_fill_merge_args(a, h_pms, two, 2, mask, mask_mode, n_ent, d_ftab, d_fbegin, n_fwin, vpf, job_begin,
                              job_count, P, F, HW, f_len, hw, d_peer);
  if (rc != DS_OK) return rc;
  a.pay[2] = eps_uncond;
  a.pay[3] = noise;
  DS_REQUIRE((prev == nullptr) == (mask_view == nullptr), DS_ERR_NULL);
  a.dd = PmDdim{h_coef[0], h_coef[1], h_coef[2], h_coef[3], h_coef[4], h_coef[5], prev, mask_view,
                mask_view_full ? 1 : 0, h_coef[6], h_coef[7]};
  DS_REQUIRE(a.dd.sa != 0.f, DS_ERR_DIM);
  return DS_OK;
}

extern "C" int ds_pm_merge_nearest_ddim(void* const* h_pms, const void* x, const void* eps_cond, const void* eps_uncond,
                                        const void* noise, const void* prev, const void* mask_view, int mask_view_full,
                                        const float* h_coef, void* mask, int mask_mode,
                                        const int32_t* d_view_ptr, const int32_t* d_ent_src,
                                        const int32_t* d_ent_dst, int n_ent, const int32_t* d_ftab,
                                        const int32_t* d_fbegin, int n_fwin, int views_per_fwin, int job_begin,
                                        int job_count, int P, int F, int HW, int f_len, int hw, void* d_peer,
                                        int dtype, void* stream) {
  using namespace ds;
  PmMergeArgs a;
  int rc = pm_fill_ddim_args(a, h_pms, x, eps_cond, eps_uncond, noise, prev, mask_view, mask_view_full, h_coef, mask, mask_mode, n_ent, d_ftab, d_fbegin,
                             n_fwin, views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw, d_peer);
  if (rc != DS_OK) return rc;
  DS_REQUIRE(d_view_ptr != nullptr && d_ent_src != nullptr && d_ent_dst != nullptr, DS_ERR_NULL);
  if (n_ent == 0 && d_peer == nullptr) return DS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long items = div_up(n_ent, kPmET);
  DS_DISPATCH_DTYPE(dtype, {
    using AT = typename Atom<ET>::type;
    const size_t smem = 2 * static_cast<size_t>(kPmET) * (f_len | 1) * sizeof(AT) + kPmET * sizeof(int32_t);
    if (f_len == 16) {
      if (d_peer != nullptr) return pm_launch(pm_merge_nearest_kernel<ET, 16, true, true>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin);
      if (hw == 2560 && g_hw_const) return pm_launch(pm_merge_nearest_kernel<ET, 16, true, false, 2560>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin);
      return pm_launch(pm_merge_nearest_kernel<ET, 16, true, false>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin);
    }
    { if (d_peer != nullptr) return pm_launch(pm_merge_nearest_kernel<ET, 0, true, true>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin); return pm_launch(pm_merge_nearest_kernel<ET, 0, true, false>, 4, items, -1, n_fwin, smem, st, a, d_view_ptr, d_ent_src, d_ent_dst, d_ftab, d_fbegin); }
  });
  return DS_OK;
}

extern "C" int ds_pm_merge_bilinear_tiled_ddim(void* const* h_pms, const void* x, const void* eps_cond,
                                               const void* eps_uncond, const void* noise, const void* prev,
                                               const void* mask_view, int mask_view_full, const float* h_coef,
                                               void* mask, int mask_mode, const int32_t* d_view_ptr,
                                               const int32_t* d_tile, int n_tiles, const int32_t* d_row_dst,
                                               const int32_t* d_row_start, int n_rows, const int32_t* d_ent_k,
                                               const float* d_ent_w, const int32_t* d_ftab, const int32_t* d_fbegin,
                                               int n_fwin, int views_per_fwin, int job_begin, int job_count,
                                               int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype,
                                               void* stream) {
  using namespace ds;
  PmMergeArgs a;
  int rc = pm_fill_ddim_args(a, h_pms, x, eps_cond, eps_uncond, noise, prev, mask_view, mask_view_full, h_coef, mask, mask_mode, n_rows, d_ftab, d_fbegin,
                             n_fwin, views_per_fwin, job_begin, job_count, P, F, HW, f_len, hw, d_peer);
  if (rc != DS_OK) return rc;
  DS_REQUIRE(d_view_ptr != nullptr && d_tile != nullptr && d_row_dst != nullptr && d_row_start != nullptr &&
             d_ent_k != nullptr && d_ent_w != nullptr, DS_ERR_NULL);
  DS_REQUIRE(pm_aligned16(d_tile) && f_len == 16 && n_tiles >= 0, DS_ERR_DIM);
  if (n_tiles == 0 && d_peer == nullptr) return DS_OK;
  const int4* d_tile4 = reinterpret_cast<const int4*>(d_tile);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = 2 * static_cast<size_t>(kPmEC) * 17 * sizeof(float4) + kPmEC * sizeof(float) + (2 * kPmTR + 1) * sizeof(int32_t);
  DS_DISPATCH_DTYPE(dtype, {
    { if (d_peer != nullptr) return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 2, true, true>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); if (hw == 2560 && g_hw_const) return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 2, true, false, 2560>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); return pm_launch(pm_merge_weighted_tiled_kernel<ET, 16, 2, true, false>, 8, n_tiles, -1, n_fwin, smem, st, a, d_view_ptr, d_tile4, d_row_dst, d_row_start, d_ent_k, d_ent_w, d_ftab, d_fbegin); }
  });
  return DS_OK;
}

extern "C" int ds_pm_peer_sync(void* d_peer, int phase, int wait, void* stream) {
  DS_REQUIRE(d_peer != nullptr, DS_ERR_NULL);
  DS_REQUIRE(phase == DS_PEER_AFTER_GATHER || phase == DS_PEER_AFTER_MERGE || phase == DS_PEER_WAIT_MERGES, DS_ERR_DIM);
  if (phase == DS_PEER_WAIT_MERGES) wait = 1;
  ds::pm_peer_sync_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<ds_peer_table_t*>(d_peer), phase,
                                                                           wait, ds::g_peer_timeout_ns);
  return ds::check_launch();
}

extern "C" int ds_peer_table_size(void) { return static_cast<int>(sizeof(ds_peer_table_t)); }

extern "C" int ds_tune(int knob, long long value) {
  switch (knob) {
    case DS_TUNE_PM_GATHER_CTAS_PER_SM: ds::g_gather_ctas_per_sm = static_cast<int>(value); return DS_OK;
    case DS_TUNE_PEER_TIMEOUT_NS: ds::g_peer_timeout_ns = value; return DS_OK;
    case DS_TUNE_PDL: ds::g_pdl = static_cast<int>(value); return DS_OK;
    case DS_TUNE_PM_GATHER_PIPELINE: ds::g_gather_pipe = static_cast<int>(value); return DS_OK;
    case DS_TUNE_PM_HW_CONST: ds::g_hw_const = value != 0 ? 1 : 0; return DS_OK;
    case DS_TUNE_PEER_MERGE_PUBLISH: ds::g_merge_publish = value != 0 ? 1 : 0; return DS_OK;
    default: return DS_ERR_DIM;
  }
}
