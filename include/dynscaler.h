/*
 * dynscaler.h -- C ABI of libdynscaler_sm100.so: the B200 (sm_100a) implementation of
 * DynamicScaler's per-denoising-step panorama tiling path.
 *
 * This is the drop-in boundary ("interface level 2").  The reference is
 * pure Python/PyTorch and has no FFI; each entry point below names the reference
 * function (file:line under the reference checkout) whose tensor arithmetic it replaces.
 * The Python shims in dynamicscaler_b200/ keep the reference's class/method signatures
 * ("interface level 1") and call these through ctypes with tensor.data_ptr().
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer unless its name starts with h_ (host).
 *   - The library allocates nothing, keeps no state besides the ds_tune knobs, and is stream-ordered and
 *     re-entrant: every call only enqueues kernels on `stream` (a cudaStream_t).
 *   - Return value: 0 = ok; < 0 = argument error (DS_ERR_*); > 0 = a cudaError_t.
 *   - dtype: DS_F32, DS_BF16 or DS_F16 (the reference's use_fp16 mode) for every latent/payload/window buffer of one call.
 *   - A panorama latent is addressed as planes [P][F][H*W] (P = B*C, frame-major, W
 *     fastest) -- the reference's own [B,C,F,H,W] memory order
 *     (shift_window_utils.py:42).  A window/view batch is [n_jobs][P][f_len][h*w].
 *   - k = y*w + x is a view pixel, p = r*W + c a panorama pixel.  Index maps are int32.
 *   - "job" j reads/writes frames (frame_begin_j + fo) mod F, fo in [0, f_len): the
 *     frame-axis ring of shift_window_utils.py:97,181 / ring_panorama_tensor_utils.py:61.
 */
#ifndef DYNSCALER_H_
#define DYNSCALER_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS_ABI_VERSION 7

enum { DS_F32 = 0, DS_BF16 = 1, DS_F16 = 2 };

enum {
  DS_OK = 0,
  DS_ERR_NULL = -1,      /* a required pointer is NULL          */
  DS_ERR_DIM = -2,       /* a dimension is <= 0 or inconsistent */
  DS_ERR_DTYPE = -3,     /* unknown dtype                       */
  DS_ERR_WINDOW = -4,    /* window outside 0 <= begin < end <= 2*size, or larger than the axis on a write */
  DS_ERR_COUNT = -5,     /* too many payloads (> DS_MAX_PAYLOADS) */
  DS_ERR_ALIGN = -6      /* vectorised variant requested on unaligned data */
};

#define DS_MAX_PAYLOADS 4

/* mask_mode of the merge calls */
enum { DS_MASK_NONE = 0, DS_MASK_PLANE = 1 /* one [H*W] plane */, DS_MASK_FULL = 2 /* [mask_planes][F][H*W] */ };

/* Process-wide knobs (tools/kbench.py; results never depend on them, only speed / patience). */
enum {
  DS_TUNE_PM_GATHER_CTAS_PER_SM = 0,   /* persistent gather: CTAs per SM, 0 = occupancy limit */
  DS_TUNE_PEER_TIMEOUT_NS = 1,         /* cross-GPU waits give up (and set ds_peer_table_t.error) after this */
  DS_TUNE_PDL = 2,                     /* programmatic dependent launch, bit mask: 1 copy-only gather, 2 4-tap gather,
                                          4 nearest merge, 8 weighted merges (default: 11) */
  DS_TUNE_PM_GATHER_PIPELINE = 3       /* software-pipelined gather tile loop (next tile's map entries prefetched with
                                          cp.async, deferred bulk-copy wait), bit mask: 1 copy-only gather, 2 4-tap gather */
  , DS_TUNE_PM_HW_CONST = 4            /* 1 (default): 40 x 64 windows (the reference's only window size) run kernels
                                          specialised on hw = 2560 (payload offsets become load immediates); 0: generic kernels */
                                       /* (5: retired -- the 5-CTA/SM weighted merge and, formerly 4, the 2-D gather tiles, both
                                          measured slower, profiles/r02a_kbench_*.jsonl) */
  , DS_TUNE_PEER_MERGE_PUBLISH = 6     /* exchange-mode ordering protocol: 0 = host-side (ds_pm_peer_sync after every pixel-major
                                          gather / merge), 1 = deferred (kernels count themselves, the next kernel publishes) */
};
int ds_tune(int knob, long long value);

int ds_abi_version(void);
/* Static string for a return code of this library (never NULL). */
const char* ds_error_string(int code);

/* ---------------------------------------------------------------- ring windows ---- */

/* RingLatent.get_window_latent (utils/shift_window_utils.py:48-114) for n windows at once:
 *   out[j][pl][fo][y][x] = pano[pl][(f0_j+fo) mod F][(top_j+y) mod H][(left_j+x) mod W]
 * d_wins: [n][3] int32 (left, top, frame_begin).  If d_wins is NULL, n must be 1 and the
 * window is (h_left, h_top, h_f0) (no device descriptor needed for the per-call API).
 * Also serves RingImageTensor.get_window_tensor (:232-263) with P=3, F=f_len=1. */
int ds_ring_gather(const void* pano, void* out, const int32_t* d_wins, int n,
                   int h_left, int h_top, int h_f0,
                   int P, int F, int H, int W, int f_len, int h, int w, int dtype, void* stream);

/* RingLatent.set_window_latent (utils/shift_window_utils.py:116-206), one window:
 *   pano[pl][(f0+fo) mod F][(top+y) mod H][(left+x) mod W] = in[pl][fo][y][x]
 * requires f_len <= F, h <= H, w <= W (":145-147 warp should not occur"). */
int ds_ring_scatter(void* pano, const void* in, int left, int top, int f0,
                    int P, int F, int H, int W, int f_len, int h, int w, int dtype, void* stream);

/* Merge plan of an ordered list of planar ring windows (the sequential
 * set_window_latent loop of pipeline/t2v_sphere_panorama_pipeline.py:609-632 collapsed
 * to "last window in order wins"): for every panorama pixel p covered by any window,
 *   plan_src[p] = j*h*w + y*w + x  of the LAST window j covering it, else -1.
 * d_wins: [n][2] int32 (left, top).  plan_src: [H*W] int32 out. */
int ds_ring_merge_plan(int32_t* plan_src, const int32_t* d_wins, int n,
                       int H, int W, int h, int w, void* stream);

/* ------------------------------------------------- sphere (equirect <-> view) ---- */

/* PanoramaTensor.get_view_tensor_no_interpolate (utils/panorama_tensor_utils.py:53-70,
 * 185-202) and its frame-ring twin (utils/ring_panorama_tensor_utils.py:59-78), batched:
 *   out[j][pl][fo][k] = pano[pl][(f0_j+fo) mod F][ idx[map_j][k] ]
 * d_idx:  [n_maps][hw] int32 panorama pixel per view pixel (column wrapped, row clamped).
 * d_jobs: [n_jobs][2] int32 (map row, frame_begin), or NULL => job j = (j, h_f0). */
int ds_gather_nearest(const void* pano, void* out, const int32_t* d_idx, const int32_t* d_jobs,
                      int h_f0, int n_jobs, int P, int F, int HW, int f_len, int hw, int dtype,
                      void* stream);

/* PanoramaTensor.get_view_tensor_interpolate (utils/panorama_tensor_utils.py:28-51;
 * ring twin ring_panorama_tensor_utils.py:31-57): F.grid_sample(bilinear, border).
 *   out = fma(t3,w3, fma(t2,w2, fma(t1,w1, t0*w0)))   (ATen's CPU summation order)
 * d_idx4: [n_maps][4][hw] int32 tap pixels (nw, ne, sw, se; clamped in bounds),
 * d_wt4:  [n_maps][4][hw] float32 tap weights. */
int ds_gather_bilinear(const void* pano, void* out, const int32_t* d_idx4, const float* d_wt4,
                       const int32_t* d_jobs, int h_f0, int n_jobs, int P, int F, int HW, int f_len,
                       int hw, int dtype, void* stream);

/* Nearest merge in gather form.  Replaces the forward scatters
 * PanoramaTensor.set_view_tensor_no_interpolation (utils/panorama_tensor_utils.py:154-183),
 * its ring twin (ring_panorama_tensor_utils.py:170-198), set_view_tensor (:72-96) and --
 * with a plan from ds_ring_merge_plan -- RingLatent.set_window_latent, for one view or
 * for a whole ordered list of views/windows ("last writer wins"):
 *   for t < T, q < Q, every plane pl and panorama frame f with d_ftab[f] >= 0:
 *     panos[q][pl][f][dst[t]] = payloads[q][(s*views_per_fwin + v)][pl][fo][k]
 *   where src[t] = v*hw + k and d_ftab[f] = s*f_len + fo.
 * d_dst/d_src: [T] int32.  If d_dst is NULL the plan is dense: T = HW, dst[t] = t and
 *   entries with src[t] < 0 are skipped.
 * d_ftab: [F] int32 or NULL (=> identity: f_len == F, s = 0, fo = f).
 * mask: optional panorama that receives 1.0 at every written location
 *   (DS_MASK_PLANE: mask[dst]; DS_MASK_FULL: mask[mpl][f][dst] for mpl < mask_planes).
 * h_panos / h_payloads: HOST arrays of Q device pointers.
 * vec: 1, or 4 when every dst/src entry addresses 4 consecutive, 16-byte aligned
 *   elements (planar ring windows on multiples of 4); then T counts groups of 4.
 * job_begin/job_count: window sharding.  Only entries whose job
 *   s*views_per_fwin + v lies in [job_begin, job_begin + job_count) are applied and the
 *   payloads hold just those jobs ([job_count][P][f_len][hw]); everything else is left
 *   untouched (mask included), so per-rank outputs have disjoint support. */
int ds_merge_nearest(void* const* h_panos, const void* const* h_payloads, int Q,
                     void* mask, int mask_mode, int mask_planes,
                     const int32_t* d_dst, const int32_t* d_src, int T, const int32_t* d_ftab,
                     int views_per_fwin, int job_begin, int job_count,
                     int P, int F, int HW, int f_len, int hw, int vec, int dtype, void* stream);

/* ds_merge_nearest in exchange mode (windows sharded over ranks; planar ring path of
 * pipeline/t2v_sphere_panorama_pipeline.py:537-632): d_peer is the local ds_peer_table_t (below) whose
 * pm[r][q] / mask[r] list every rank's REFERENCE-ORDER replica; every element this rank's jobs win is
 * stored into all of them over NVLink (peer st.global, 16-byte vectors when vec == 4), so the sharded
 * planar merge needs no zeroed sum buffer, all-reduce or select.  h_panos / mask are only validated.
 * The host orders the phases with ds_pm_peer_sync(wait = 1) after the gather and after the merge. */
int ds_merge_nearest_peer(void* const* h_panos, const void* const* h_payloads, int Q,
                          void* mask, int mask_mode, int mask_planes,
                          const int32_t* d_dst, const int32_t* d_src, int T,
                          const int32_t* d_ftab, int views_per_fwin, int job_begin, int job_count,
                          int P, int F, int HW, int f_len, int hw, int vec, void* d_peer, int dtype, void* stream);

/* ds_merge_nearest fused with the denoise epilogue on REFERENCE-ORDER tensors -- the write side of the
 * planar shift-window loop, pipeline/t2v_normal_pipeline.py:526-557 (``_basic_denoise_one_step`` :572-617
 * followed by the two ``set_window_latent`` calls), and of any nearest plan:
 *     e = eps_uncond + g*(eps_cond - eps_uncond)            (:603; eps_uncond NULL: e = eps_cond)
 *     pred_x0 = (x - s1m*e)/sa;  x_prev = sp*pred_x0 + sd*e [+ sig*noise]     (pipeline/scheduler.py:84-94)
 * h_panos[0] receives x_prev, h_panos[1] pred_x0, for the winning elements only.  x / eps_* / noise:
 * [n_jobs][P][f_len][hw] like the payloads of ds_merge_nearest; h_coef = {sa, s1m, sp, sd, sig, g, r, 1-r}
 * (float32, formed like the reference's torch.full tensors).  prev != NULL: the merge_prev_denoised_ratio
 * mix of the planar I2V loop (pipeline/i2v_sphere_panorama_pipeline.py:938-943) between ddim_step and the
 * latent set: x_prev <- mix_latents_with_mask(x_prev, prev, mask_view, r), prev / mask_view in window layout
 * like x (the window as gathered, :866, and the denoised-mask window read at gather time, :868-873).
 * d_peer NULL: local stores; else the exchange mode of ds_merge_nearest_peer.  Everything else as
 * ds_merge_nearest. */
int ds_merge_nearest_ddim(void* const* h_panos, const void* x, const void* eps_cond, const void* eps_uncond,
                          const void* noise, const void* prev, const void* mask_view, const float* h_coef,
                          void* mask, int mask_mode, int mask_planes,
                          const int32_t* d_dst, const int32_t* d_src, int T, const int32_t* d_ftab,
                          int views_per_fwin, int job_begin, int job_count, int P, int F, int HW, int f_len,
                          int hw, int vec, void* d_peer, int dtype, void* stream);

/* Weighted scatter-accumulate + per-pixel weight normalisation in gather form:
 * PanoramaTensor.set_view_tensor_bilinear (utils/panorama_tensor_utils.py:98-152; ring twin
 * ring_panorama_tensor_utils.py:107-166).  Row t owns panorama pixel d_row_dst[t] and
 * the CSR entries e in [d_row_ptr[t], d_row_ptr[t+1]), listed tap-major then k ascending
 * (the order index_add_ applies them):
 *     acc = 0; ws = 0;  for e: acc = acc + (payload[src_e] * w_e);  ws = ws + w_e
 *     if ws > 0: pano[...][dst] = acc / ws            (mul, add, div individually rounded)
 * Frame/plane/payload addressing, the mask and job_begin/job_count as in
 * ds_merge_nearest (all entries of one row come from one view). */
int ds_merge_bilinear(void* const* h_panos, const void* const* h_payloads, int Q,
                      void* mask, int mask_mode, int mask_planes,
                      const int32_t* d_row_dst, const int32_t* d_row_ptr, int T,
                      const int32_t* d_ent_src, const float* d_ent_w, const int32_t* d_ftab,
                      int views_per_fwin, int job_begin, int job_count,
                      int P, int F, int HW, int f_len, int hw, int dtype, void* stream);

/* ------------------------------------------- pixel-major resident panorama path ---- */
/* Multi-GPU exchange (windows sharded over ranks).  Every rank keeps a replica of
 * the pixel-major panoramas, the mask and one ds_peer_table_t in PEER-MAPPED device memory
 * (e.g. torch symmetric memory, cudaIpc or cuMem fabric handles: the host maps them, this
 * library only dereferences the addresses).  Passing the LOCAL table as d_peer switches the
 * merge calls below to exchange mode: every surviving row is stored into ALL replicas
 * -- one multimem.st to the NVLS multicast alias (mc_pm[q], mc_mask) when the host provides it,
 * else one peer store / TMA bulk copy per replica (pm[r][q], mask[r]) -- so there is no sum
 * buffer, all-reduce or select.
 * Ordering across GPUs is two monotonic counters per peer: a rank's gather count and merge count, each
 * published into every peer's table with st.release.sys.  Host-side protocol (default): the host enqueues
 * ds_pm_peer_sync (one 32-thread kernel) after every gather (DS_PEER_AFTER_GATHER) and every merge
 * (DS_PEER_AFTER_MERGE); it counts the launch as issued, publishes the count and, if wait != 0, also waits
 * for the peers' counts.  Deferred protocol (DS_TUNE_PEER_MERGE_PUBLISH = 1): the pixel-major kernels count
 * themselves as issued when they finish and the NEXT kernel on the state publishes the count in its
 * prologue, so a step has no ordering launches at all (the host must then NOT enqueue the sync, except
 * DS_PEER_WAIT_MERGES before it touches a replica itself).  With d_peer the data kernels do the waiting
 * themselves (ld.acquire.sys, overlapped with their first loads): a gather starts reading only when every
 * peer has published as many merges as this rank has issued (replicas complete), a merge starts storing
 * only when every peer has published as many gathers (replicas may be overwritten).
 * All ranks must issue the same gather / merge sequence on a state.  A wait that exceeds
 * DS_TUNE_PEER_TIMEOUT_NS sets `error`.  The host zero-fills the table, then writes rank,
 * world, pm, mask and ctl (the address of every rank's table, own included) and makes sure
 * (host barrier) that every rank has done so before the first call. */
#define DS_MAX_PEERS 8
typedef struct ds_peer_table {
  int32_t merges_done[DS_MAX_PEERS];            /* [p]: written by peer p */
  int32_t gathers_done[DS_MAX_PEERS];           /* [p]: written by peer p */
  int32_t my_merges, my_gathers;                /* this rank's counts as PUBLISHED to the peers */
  int32_t issued_gathers, issued_merges;        /* gather / merge launches of this rank that have completed (>= published) */
  int32_t error, rank, world, reserved;
  uint64_t pm[DS_MAX_PEERS][DS_MAX_PAYLOADS];   /* replica r, panorama q */
  uint64_t mask[DS_MAX_PEERS];                  /* replica r of the mask panorama (or 0) */
  uint64_t ctl[DS_MAX_PEERS];                   /* rank r's ds_peer_table_t */
  uint64_t mc_pm[DS_MAX_PAYLOADS];              /* NVLS multicast alias of panorama q over ALL replicas, or 0 */
  uint64_t mc_mask;                             /* multicast alias of the mask, or 0 */
} ds_peer_table_t;
int ds_peer_table_size(void);                   /* sizeof(ds_peer_table_t), for bindings */
enum { DS_PEER_AFTER_GATHER = 0, DS_PEER_AFTER_MERGE = 1,
       DS_PEER_WAIT_MERGES = 2 /* counts nothing: publishes what is pending and waits until every peer has pushed as many
                                  merges as this rank has issued -- before the host touches its replica outside the
                                  kernels above (e.g. zeroing the mask panorama at the start of a step) */ };
int ds_pm_peer_sync(void* d_peer, int phase, int wait, void* stream);

/* The batched step driver keeps the panorama of a stage resident as pm[pixel][frame][plane]
 * (H*W pixels, F frames, P == 4 planes; element (pl, f, p) of the reference's [P][F][H*W]
 * order lives at (p*F + f)*4 + pl).  Window tensors keep the reference's
 * [job][plane][frame][h*w] order.  Same reference functions as the plane-major entry
 * points above; only the resident layout differs. */

/* get_view_tensor_no_interpolate (taps == 1; panorama_tensor_utils.py:53-70,185-202) or
 * get_view_tensor_interpolate (taps == 4; :28-51) for all jobs of a step:
 *   out[j][pl][fo][k] = sum_t wt[map_j][t][k] * pm[idx[map_j][t][k]][(f0_j+fo) mod F][pl]
 * d_idx: [n_maps][taps][hw] int32; d_wt: [n_maps][4][hw] float32 (taps == 4 only).
 * d_idx == NULL (taps == 1): identity map, i.e. the unpack pm -> [P][F][hw], hw <= H*W.
 * d_jobs as in ds_gather_nearest.  Window rows that are 16-byte multiples (hw % (16 / elem_size) == 0)
 * leave through TMA bulk copies, other sizes through plain stores.
 * d_peer: NULL, or the local ds_peer_table_t: the kernel first waits for the peers' merges. */
int ds_pm_gather(const void* pm, void* out, const int32_t* d_idx, const float* d_wt, int taps,
                 const int32_t* d_jobs, int h_f0, int n_jobs, int P, int F, int HW, int f_len,
                 int hw, void* d_peer, int dtype, void* stream);

/* Gather fused with the re-noise + mask mix that follows it in the reference's loops (SURVEY §8f-1;
 * pipeline/t2v_sphere_panorama_pipeline.py:194-221, pipeline/i2v_sphere_panorama_pipeline.py:330-365):
 *   view   = get_view_tensor_no_interpolate(latent panorama)         panorama_tensor_utils.py:53-70
 *   m      = get_view_tensor_no_interpolate(denoised-mask panorama)
 *   noised = c * view + s * eps                                       pipeline/scheduler.py:98-110
 *   out    = mix_latents_with_mask(view, noised, m, ratio)           utils/tensor_utils.py:19-39
 * for every job, each elementwise op rounded on its own like the reference's tensor ops.  eps is
 * INJECTED (same layout as out; the reference draws randn_like inside re_noise).
 * mask: DS_MASK_PLANE -> the [H*W] plane of the T2V loop (t2v_...:159); DS_MASK_FULL -> a
 * pixel-major panorama (the latent-shaped I2V mask, i2v_...:220).  d_idx / d_midx: nearest read maps
 * [n_maps][hw] of the latent and of the mask read (they differ only when phi_fov_dict gives the
 * mask another fov, t2v_...:194,206).  h_coef: HOST float[4] = { c, s, ratio, 1 - ratio }.
 * out_raw (or NULL): the unmixed view (view_latent_prev_denoise, i2v_...:345); out_mask (or NULL):
 * m, as [jobs][hw] (DS_MASK_PLANE) or in window layout (DS_MASK_FULL) -- both feed the merge
 * prologue of ds_pm_merge_*_ddim.  d_jobs, h_f0, d_peer as in ds_pm_gather. */
int ds_pm_gather_renoise(const void* pm, void* out, void* out_raw, void* out_mask, const void* eps,
                         const void* mask, int mask_mode, const int32_t* d_idx, const int32_t* d_midx,
                         const float* h_coef, const int32_t* d_jobs, int h_f0, int n_jobs, int P, int F,
                         int HW, int f_len, int hw, void* d_peer, int dtype, void* stream);

/* Ordered sequence of set_view_tensor_no_interpolation calls (panorama_tensor_utils.py:
 * 154-183; ring twin ring_panorama_tensor_utils.py:170-198), scatter form over the surviving
 * view pixels only ("last writer wins").  Entry e says: view pixel
 * d_ent_src[e] = v*hw + k wins panorama pixel d_ent_dst[e].  Entries are sorted by view and
 * d_view_ptr[v] (length views_per_fwin + 1) is the first entry of view v:
 *   pms[q][dst][(fb_s+fo) mod F][pl] = payloads[q][s*views_per_fwin + v - job_begin][pl][fo][k]
 * for every frame window s < n_fwin whose job is in [job_begin, job_begin + job_count) and
 * every fo with d_ftab[(fb_s+fo) mod F] == s*f_len + fo (d_ftab/d_fbegin NULL: one window,
 * identity).  All three plan arrays NULL: identity plan over n_ent pixels, i.e. the pack
 * [P][F][hw] -> pm.
 * mask: DS_MASK_PLANE -> mask[dst] = 1 ([H*W] plane); DS_MASK_FULL -> a pixel-major
 * panorama that receives 1 in every written atom.
 * d_peer: NULL, or the local ds_peer_table_t: rows go to every replica listed there (h_pms /
 * mask are then only validated) and job_count may be 0. */
int ds_pm_merge_nearest(void* const* h_pms, const void* const* h_payloads, int Q,
                        void* mask, int mask_mode, const int32_t* d_view_ptr,
                        const int32_t* d_ent_src, const int32_t* d_ent_dst, int n_ent,
                        const int32_t* d_ftab, const int32_t* d_fbegin, int n_fwin,
                        int views_per_fwin, int job_begin, int job_count,
                        int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype, void* stream);

/* Ordered sequence of set_view_tensor_bilinear calls (panorama_tensor_utils.py:98-152; ring
 * twin ring_panorama_tensor_utils.py:107-166).  Row r is panorama pixel d_row_dst[r], owned
 * by one view, with CSR entries [d_row_ptr[r], d_row_ptr[r+1]) of (view pixel d_ent_src =
 * v*hw + k, weight d_ent_w) in index_add_ order (tap-major, k ascending); rows are sorted by
 * view and d_view_ptr[v] is the first row of view v:
 *   acc = acc + (payload * w); ws = ws + w; pm[...][dst] = acc / ws  if ws > 0
 * Frame windows, sharding and mask as in ds_pm_merge_nearest. */
int ds_pm_merge_bilinear(void* const* h_pms, const void* const* h_payloads, int Q,
                         void* mask, int mask_mode, const int32_t* d_view_ptr,
                         const int32_t* d_row_dst, const int32_t* d_row_ptr, int n_rows,
                         const int32_t* d_ent_src, const float* d_ent_w,
                         const int32_t* d_ftab, const int32_t* d_fbegin, int n_fwin,
                         int views_per_fwin, int job_begin, int job_count,
                         int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype, void* stream);

/* ds_pm_merge_bilinear on a TILED plan (f_len == 16, Q in {1, 2, 4}; fast path).  The host pads
 * the CSR entry list so that tile t owns entries [64t, 64t + 64) (d_ent_k < 0 = padding; all of
 * one view, k relative to that view) and at most 32 whole rows:
 * d_tile[t] = (first row, view, rows, entries) as 4 x int32 (16-byte aligned);
 * d_view_ptr[v] = first tile of view v; d_row_start[r] = first padded entry of row r (the rows
 * of a tile are contiguous in the padded arrays).  Same arithmetic and result as
 * ds_pm_merge_bilinear. */
int ds_pm_merge_bilinear_tiled(void* const* h_pms, const void* const* h_payloads, int Q,
                               void* mask, int mask_mode, const int32_t* d_view_ptr,
                               const int32_t* d_tile, int n_tiles, const int32_t* d_row_dst,
                               const int32_t* d_row_start, int n_rows, const int32_t* d_ent_k,
                               const float* d_ent_w, const int32_t* d_ftab, const int32_t* d_fbegin,
                               int n_fwin, int views_per_fwin, int job_begin, int job_count,
                               int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype, void* stream);

/* ------------------------------------------------------------------ elementwise ---- */

/* mix_latents_with_mask (utils/tensor_utils.py:19-39):
 *   mixed = l1*(1-r) + add*r;  out = l1*(1-m) + mixed*m     (no FMA contraction)
 * one_minus_ratio is passed separately because the reference forms (1 - mix_ratio) in
 * Python double precision before the tensor multiply.
 * mask_hw > 0: mask is one [mask_hw] plane broadcast over n/mask_hw planes (the [1,h,w]
 * case, :21-23); mask_hw == 0: mask has n elements (the 5-D case, :24-25). */
int ds_mix_mask(const void* l1, const void* add, const void* mask, void* out,
                float ratio, float one_minus_ratio, long long n, int mask_hw, int dtype, void* stream);

/* Fused denoise epilogue + merge.  Instead of reading the payloads x_prev / pred_x0
 * the merge computes them, for the surviving elements only, from the window latent x
 * ([jobs][P][f_len][hw], what ds_pm_gather produced) and the UNet outputs, with the reference's
 * elementwise arithmetic (each op rounded separately):
 *   e       = eps_uncond + g * (eps_cond - eps_uncond)   (eps_uncond NULL: e = eps_cond)
 *                                                   pipeline/t2v_sphere_panorama_pipeline.py:256-261
 *   pred_x0 = (x - s1m * e) / sa                    pipeline/scheduler.py:84
 *   x_prev  = (sp * pred_x0 + sd * e) + sig * noise :86-89   (noise NULL: the term is skipped)
 * Optional merge prologue, the merge_prev_denoised_ratio mix of the I2V loops
 * (pipeline/i2v_sphere_panorama_pipeline.py:430-435, :938-943) applied to x_prev before it is merged:
 *   x_prev = mix_latents_with_mask(x_prev, prev, mask_view, r)       utils/tensor_utils.py:19-39
 * prev: the window as gathered (view_latent_prev_denoise, :345), same layout as x; mask_view: the
 * denoised-mask view read at gather time ([jobs][hw] if mask_view_full == 0, else window layout);
 * both NULL: no mix.
 * h_coef: HOST float[8] = { sa = sqrt(alpha_t), s1m = sqrt(1 - alpha_t), sp = sqrt(alpha_prev),
 *   sd = sqrt(1 - alpha_prev - sigma^2), sig = sigma_t, g = guidance scale (each computed in float32
 *   like the reference's torch.full(...) tensors), r = merge_prev ratio, 1 - r (formed in double
 *   like the reference) }.
 * h_pms: exactly 2 panoramas: [0] <- x_prev (latent), [1] <- pred_x0 (denoised).
 * Plans, frame windows, sharding, mask and d_peer as in ds_pm_merge_nearest /
 * ds_pm_merge_bilinear_tiled. */
int ds_pm_merge_nearest_ddim(void* const* h_pms, const void* x, const void* eps_cond, const void* eps_uncond,
                             const void* noise, const void* prev, const void* mask_view, int mask_view_full,
                             const float* h_coef, void* mask, int mask_mode,
                             const int32_t* d_view_ptr, const int32_t* d_ent_src,
                             const int32_t* d_ent_dst, int n_ent, const int32_t* d_ftab,
                             const int32_t* d_fbegin, int n_fwin, int views_per_fwin, int job_begin,
                             int job_count, int P, int F, int HW, int f_len, int hw, void* d_peer,
                             int dtype, void* stream);
int ds_pm_merge_bilinear_tiled_ddim(void* const* h_pms, const void* x, const void* eps_cond,
                                    const void* eps_uncond, const void* noise, const void* prev,
                                    const void* mask_view, int mask_view_full, const float* h_coef,
                                    void* mask, int mask_mode, const int32_t* d_view_ptr,
                                    const int32_t* d_tile, int n_tiles, const int32_t* d_row_dst,
                                    const int32_t* d_row_start, int n_rows, const int32_t* d_ent_k,
                                    const float* d_ent_w, const int32_t* d_ftab, const int32_t* d_fbegin,
                                    int n_fwin, int views_per_fwin, int job_begin, int job_count,
                                    int P, int F, int HW, int f_len, int hw, void* d_peer, int dtype,
                                    void* stream);

/* resize_video_latent (utils/diffusion_utils.py:21-33): F.interpolate(mode, align_corners) of every
 * (b, c, f) plane, in [planes][h][w] -> out [planes][H][W].  mode 'nearest' ignores
 * align_corners (the reference forces it to None, :22-23). */
enum { DS_RESIZE_NEAREST = 0, DS_RESIZE_BILINEAR = 1, DS_RESIZE_BICUBIC = 2 };
int ds_resize_planes(const void* in, void* out, long long planes, int h, int w, int H, int W,
                     int mode, int align_corners, int dtype, void* stream);

/* ------------------------------------------ pixel-side neighbours ---- */
/* One tile of tiled_vae_encode_tensor_simple (pipeline/i2v_sphere_panorama_pipeline.py:505-562):
 *   img_latent[:, :, :, dst_top:dst_top+ch, dst_left:dst_left+cw] += tile[:, :, :, top_cut:top_cut+ch, left_cut:left_cut+cw]
 *   count[:, :, :, same region] += 1                                                     (:550-554)
 * acc: float32 [planes][H][W] (the reference accumulates in float32 whatever the tile dtype, :521);
 * count: float32 [planes / planes_per_count][H][W] (the reference's [B,1,1,H,W]: planes_per_count = 4*F);
 * tile: [planes][th][tw] of `dtype`.  Calls on one stream are applied in order. */
int ds_tile_accumulate(float* acc, float* count, const void* tile, long long planes, long long planes_per_count,
                       int H, int W, int th, int tw, int top_cut, int left_cut, int ch, int cw,
                       int dst_top, int dst_left, int dtype, void* stream);
/* count = clamp(count, min=1); out = acc / count   (:556-558), float32, IEEE division. */
int ds_tile_normalise(const float* acc, const float* count, float* out, long long planes,
                      long long planes_per_count, long long HW, void* stream);

/* paste_on_static (pipeline/i2v_sphere_panorama_pipeline.py:245-254): the frame-0 latent x0 [planes][HW]
 * repeated over F frames (torch.cat([x0] * F, dim=2), :248) and re-noised with injected epsilon
 * (pipeline/scheduler.py:98-110): out[pl][f][p] = c * x0[pl][p] + s * eps[pl][f][p], each op rounded
 * on its own.  eps == NULL: the plain repeat. */
int ds_repeat_renoise(const void* x0, const void* eps, void* out, float c, float s, long long planes, int F,
                      long long HW, int dtype, void* stream);

/* The read side of the planar I2V loop after its two window gathers (pipeline/i2v_sphere_panorama_pipeline.py:
 * 875-884), one pass:  noised = c*x + s*eps  (re_noise, pipeline/scheduler.py:98-110, epsilon injected; c, s
 * cast to the tensor dtype like the reference's 0-dim tensors);  out = mix_latents_with_mask(x, noised, mask,
 * ratio) (utils/tensor_utils.py:19-39).  x / eps / mask / out: n elements each; every op rounded on its own.
 * mask_plane_hw > 0: the T2V planar loop's variant (pipeline/t2v_sphere_panorama_pipeline.py:552-556,
 * ``window_denoised_mask[0, 0, [0]]``): each window of window_elems elements is mixed with the first h*w =
 * mask_plane_hw elements (batch 0, channel 0, frame 0) of its own mask window, broadcast over C and F. */
int ds_renoise_mix(const void* x, const void* eps, const void* mask, void* out, float c, float s, float ratio,
                   float one_minus_ratio, long long n, long long window_elems, int mask_plane_hw, int dtype,
                   void* stream);

/* Final select of the sharded merge:
 *   pano[i] = weight[(i / weight_group) mod weight_period] > 0 ? sum[i] : pano[i]   for i < n.
 * weight has the panorama's dtype.  Reference layout: weight_group = 1 and weight_period =
 * H*W (one plane), F*H*W (one plane per frame) or n.  Pixel-major layout: weight_group =
 * F*P with weight_period = H*W (one weight per pixel), or group 1 / period n. */
int ds_select_merged(void* pano, const void* sum, const void* weight,
                     long long n, long long weight_period, long long weight_group, int dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DYNSCALER_H_ */
