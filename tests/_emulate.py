"""numpy emulation of the kernels' contracts (include/dynscaler.h), used by the CPU test
suite to validate the host-built maps / plans without a GPU.  Test-only."""
import numpy as np

F32 = np.float32


def gather_nearest(pano, idx, jobs, f_len):
    """pano [P, F, HW]; idx [n_maps, hw]; jobs [(map, f0)] -> [n_jobs, P, f_len, hw]"""
    P, F, HW = pano.shape
    out = np.empty((len(jobs), P, f_len, idx.shape[1]), pano.dtype)
    for j, (m, f0) in enumerate(jobs):
        fi = (f0 + np.arange(f_len)) % F
        out[j] = pano[:, fi][:, :, idx[m]]
    return out


def gather_bilinear(pano, idx4, wt4, jobs, f_len):
    P, F, HW = pano.shape
    out = np.empty((len(jobs), P, f_len, idx4.shape[2]), pano.dtype)
    fma = lambda a, b, c: (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)
    for j, (m, f0) in enumerate(jobs):
        fi = (f0 + np.arange(f_len)) % F
        src = pano[:, fi]
        r = src[:, :, idx4[m, 0]] * wt4[m, 0]
        for t in (1, 2, 3):
            r = fma(src[:, :, idx4[m, t]], wt4[m, t][None, None], r)
        out[j] = r
    return out


def _sources(ftab, F, f_len, vpf, v, job_begin, job_count):
    """per panorama frame f: (local job, fo) or None"""
    res = []
    for f in range(F):
        sfo = f if ftab is None else int(ftab[f])
        if sfo < 0:
            res.append(None)
            continue
        s, fo = divmod(sfo, f_len)
        job = s * vpf + v - job_begin
        res.append((job, fo) if 0 <= job < job_count else None)
    return res


def merge_nearest(panos, payloads, dst, src, ftab, vpf, f_len, hw, mask=None, mask_mode=0,
                  job_begin=0, job_count=1 << 30):
    """panos[q] [P, F, HW] in place; payloads[q] [n_local_jobs, P, f_len, hw]"""
    P, F, HW = panos[0].shape if panos else (0, len(ftab) if ftab is not None else 0, 0)
    if dst is None:
        keep = src >= 0
        dst, src = np.nonzero(keep)[0], src[keep]
    v_all, k_all = src // hw, src % hw
    for v in np.unique(v_all):
        sel = v_all == v
        d, k = dst[sel], k_all[sel]
        for f, sf in enumerate(_sources(ftab, F, f_len, vpf, int(v), job_begin, job_count)):
            if sf is None:
                continue
            job, fo = sf
            for q in range(len(panos)):
                panos[q][:, f, d] = payloads[q][job][:, fo][:, k]
            if mask_mode == 2:
                mask[:, f, d] = 1
            if mask_mode == 1 and f == 0:
                mask[d] = 1


def merge_bilinear(panos, payloads, plan, ftab, vpf, f_len, hw, mask=None, mask_mode=0,
                   job_begin=0, job_count=1 << 30):
    row_dst, row_ptr, ent_src, ent_w = plan
    P, F, HW = panos[0].shape
    for t in range(row_dst.shape[0]):
        e0, e1 = row_ptr[t], row_ptr[t + 1]
        if e1 <= e0:
            continue
        v = int(ent_src[e0]) // hw
        ks = ent_src[e0:e1] - v * hw
        ws_ = F32(0)
        for w in ent_w[e0:e1]:
            ws_ = F32(ws_ + w)
        if not ws_ > 0:
            continue
        for f, sf in enumerate(_sources(ftab, F, f_len, vpf, v, job_begin, job_count)):
            if sf is None:
                continue
            job, fo = sf
            for q in range(len(panos)):
                acc = np.zeros(P, F32)
                x = payloads[q][job][:, fo]                       # [P, hw]
                for k, w in zip(ks, ent_w[e0:e1]):
                    acc = acc + x[:, k] * w
                panos[q][:, f, row_dst[t]] = acc / ws_
            if mask_mode == 2:
                mask[:, f, row_dst[t]] = 1
            if mask_mode == 1 and f == 0:
                mask[row_dst[t]] = 1


def ring_merge_plan(windows, H, W, h, w):
    """dense plan_src [H*W] of ds_ring_merge_plan"""
    plan = np.full((H, W), -1, np.int32)
    for j, (l, t) in enumerate(windows):
        ys = (t + np.arange(h)) % H
        xs = (l + np.arange(w)) % W
        plan[np.ix_(ys, xs)] = (j * h + np.arange(h))[:, None] * w + np.arange(w)[None, :]
    return plan.reshape(-1)


# ------------------------------------------------ pixel-major kernels (csrc/pm.cu) ----
def pm_pack(pano):
    """[P, F, HW] -> pm [HW, F, P]"""
    return np.ascontiguousarray(pano.transpose(2, 1, 0))


def pm_unpack(pm):
    return np.ascontiguousarray(pm.transpose(2, 1, 0))


def _pm_view_range(view_ptr, vpf, s, job_begin, job_count):
    v_lo = min(max(job_begin - s * vpf, 0), vpf)
    v_hi = min(max(job_begin + job_count - s * vpf, 0), vpf)
    return int(view_ptr[v_lo]), int(view_ptr[v_hi])


def pm_merge_nearest(pms, payloads, plan, ftab, fbegin, vpf, f_len, hw, mask=None, mask_mode=0,
                     job_begin=0, job_count=1 << 30):
    """Scatter form like ds_pm_merge_nearest: pms[q] [HW, F, P] in place."""
    view_ptr, ent_src, ent_dst = plan
    F = pms[0].shape[1] if pms else mask.shape[1]
    n_fwin = 1 if fbegin is None else len(fbegin)
    for s in range(n_fwin):
        lo, hi = _pm_view_range(view_ptr, vpf, s, job_begin, job_count)
        fb = 0 if fbegin is None else fbegin[s] % F
        for e in range(lo, hi):
            v, k = divmod(int(ent_src[e]), hw)
            job, d = s * vpf + v - job_begin, ent_dst[e]
            assert 0 <= job < min(job_count, len(payloads[0]) if payloads else 1 << 30)
            if mask_mode == 1:
                mask[d] = 1
            for fo in range(f_len):
                f = (fb + fo) % F
                if ftab is not None and ftab[f] != s * f_len + fo:
                    continue
                for q in range(len(pms)):
                    pms[q][d, f, :] = payloads[q][job][:, fo, k]
                if mask_mode == 2:
                    mask[d, f, :] = 1


def pm_merge_bilinear(pms, payloads, plan, ftab, fbegin, vpf, f_len, hw, mask=None, mask_mode=0,
                      job_begin=0, job_count=1 << 30):
    view_ptr, row_dst, row_ptr, ent_src, ent_w = plan
    F = pms[0].shape[1]
    n_fwin = 1 if fbegin is None else len(fbegin)
    for s in range(n_fwin):
        lo, hi = _pm_view_range(view_ptr, vpf, s, job_begin, job_count)
        fb = 0 if fbegin is None else fbegin[s] % F
        for r in range(lo, hi):
            e0, e1 = row_ptr[r], row_ptr[r + 1]
            ws_ = F32(0)
            for w in ent_w[e0:e1]:
                ws_ = F32(ws_ + w)
            if not ws_ > 0:
                continue
            if mask_mode == 1:
                mask[row_dst[r]] = 1
            for fo in range(f_len):
                f = (fb + fo) % F
                if ftab is not None and ftab[f] != s * f_len + fo:
                    continue
                for q in range(len(pms)):
                    acc = np.zeros(pms[q].shape[2], F32)
                    for src, w in zip(ent_src[e0:e1], ent_w[e0:e1]):
                        v, k = divmod(int(src), hw)
                        acc = acc + payloads[q][s * vpf + v - job_begin][:, fo, k] * w
                    pms[q][row_dst[r], f, :] = acc / ws_
                if mask_mode == 2:
                    mask[row_dst[r], f, :] = 1


def pm_merge_bilinear_tiled(pms, payloads, plan, ftab, fbegin, vpf, f_len, hw, mask=None, mask_mode=0,
                            job_begin=0, job_count=1 << 30):
    """ds_pm_merge_bilinear_tiled, tile by tile (padding entries have ent_k < 0)."""
    view_ptr, tile, row_dst, row_start, ent_k, ent_w = plan
    F = pms[0].shape[1]
    n_fwin = 1 if fbegin is None else len(fbegin)
    for s in range(n_fwin):
        t_lo, t_hi = _pm_view_range(view_ptr, vpf, s, job_begin, job_count)
        fb = 0 if fbegin is None else fbegin[s] % F
        for t in range(t_lo, t_hi):
            r0, v, nr, ne = (int(x) for x in tile[t])
            base = t * 64
            assert nr <= 32 and ne <= 64 and np.all(ent_k[base:base + ne] >= 0) and np.all(ent_k[base + ne:base + 64] < 0)
            job = s * vpf + v - job_begin
            starts = [int(row_start[r0 + i]) for i in range(nr)] + [base + ne]
            for i in range(nr):
                e0, e1 = starts[i], starts[i + 1]
                assert base <= e0 <= e1 <= base + ne
                ws_ = F32(0)
                for w in ent_w[e0:e1]:
                    ws_ = F32(ws_ + w)
                if not ws_ > 0:
                    continue
                d = row_dst[r0 + i]
                if mask_mode == 1:
                    mask[d] = 1
                for fo in range(f_len):
                    f = (fb + fo) % F
                    if ftab is not None and ftab[f] != s * f_len + fo:
                        continue
                    for q in range(len(pms)):
                        acc = np.zeros(pms[q].shape[2], F32)
                        for k, w in zip(ent_k[e0:e1], ent_w[e0:e1]):
                            acc = acc + payloads[q][job][:, fo, int(k)] * w
                        pms[q][d, f, :] = acc / ws_
                    if mask_mode == 2:
                        mask[d, f, :] = 1
