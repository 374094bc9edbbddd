"""Host-side geometry of the product (dynamicscaler_b200/maps.py, step planning) against
the reference's golden vectors, by emulating the kernel contracts in numpy.  CPU only:
no kernels run here, the oracle is only the checker."""
import numpy as np
import pytest

from _emulate import gather_bilinear, gather_nearest, merge_bilinear, merge_nearest, ring_merge_plan
from conftest import randn
from dynamicscaler_b200 import maps
from dynamicscaler_b200.shift_window_utils import get_dimension_slices_and_sizes
from dynamicscaler_b200.step import shard_range
from oracle import panorama_oracle as po


def test_uv_maps_bit_exact_vs_reference(small):
    for i, (fov, th, ph, w, h, H, W) in enumerate(small["uv/cases"]):
        if f"uv/{i}/u" not in small:
            continue
        vm = maps.view_map(fov, th, ph, int(w), int(h), int(H), int(W))
        assert np.array_equal(vm.u, small[f"uv/{i}/u"]) and np.array_equal(vm.v, small[f"uv/{i}/v"])


def test_index_maps_bit_exact_config2(configs):
    """north_star: 'window index/wrap maps bit-exact'."""
    for n, (ph, th) in enumerate(configs["cfg2/views"]):
        vm = maps.view_map(120, th, ph, 64, 40, 128, 256)
        assert vm.idx_read.dtype == np.int32
        assert np.array_equal(vm.idx_read, configs["cfg2/idx_read"][n])
        dst, src = vm.write_plan
        tgt = configs["cfg2/idx_write"][n]
        win = po.last_writer_flags(tgt.astype(np.int64))
        assert np.array_equal(np.sort(tgt[win]), dst) and np.array_equal(tgt[src], dst)
        assert win[src].all()


def test_angles_not_reduced_mod_360():
    a = maps.view_map(120, 48, 45, 64, 40, 128, 256)
    b = maps.view_map(120, 408, 45, 64, 40, 128, 256)
    assert a is not b and (a.idx_read != b.idx_read).any()


def test_slices_helper(small):
    for i, (b, e, s) in enumerate(small["slices/cases"]):
        sl, sz = get_dimension_slices_and_sizes(int(b), int(e), int(s))
        assert [x.stop - x.start for x in sl] == sz
        assert np.array_equal(np.concatenate([np.arange(x.start, x.stop) for x in sl]), small[f"slices/{i}"])


def _planes(a):   # [B, C, F, H, W] -> [P, F, HW]
    B, C, F, H, W = a.shape
    return a.reshape(B * C, F, H * W).copy()


def test_per_view_plans_small(small):
    pano5 = randn(21, 1, 2, 3, 16, 32)
    for i, (fov, th, ph) in enumerate(small["sph/views"]):
        vm = maps.view_map(fov, th, ph, 10, 6, 16, 32)
        view_in = randn(300 + i, 1, 2, 3, 6, 10).reshape(1, 2, 3, 60)
        got = gather_nearest(_planes(pano5), vm.idx_read[None], [(0, 0)], 3)
        assert np.array_equal(got.reshape(1, 2, 3, 6, 10), small[f"sph/get_nearest/{i}"])
        idx4, wt4 = vm.read_taps()
        got = gather_bilinear(_planes(pano5), idx4[None], wt4[None], [(0, 0)], 3)
        assert np.abs(got.reshape(1, 2, 3, 6, 10) - small[f"sph/get_bilinear/{i}"]).max() <= 1e-6
        lat = _planes(pano5)
        merge_nearest([lat], [view_in], *vm.write_plan, None, 1, 3, 60)
        assert np.array_equal(lat.reshape(pano5.shape), small[f"sph/set_nearest/{i}"])
        lat = _planes(pano5)
        merge_bilinear([lat], [view_in], maps.bilinear_merge_plan([vm]), None, 1, 3, 60)
        assert np.array_equal(lat.reshape(pano5.shape), small[f"sph/set_bilinear/{i}"])
        # quirk Q3 of set_view_tensor: L = B*F = 3 images
        lat = pano5.copy()
        for b in range(3):
            dst, src = maps.round_merge_plan(vm, 3, b)
            img = lat[0, :, b].reshape(2, -1).copy()
            img[:, dst] = view_in[0, :, b][:, src]
            lat[0, :, b] = img.reshape(2, 16, 32)
        assert np.array_equal(lat, small[f"sph/set_round/{i}"])


def test_frame_ring_plans_small(small):
    rp5 = randn(31, 1, 2, 7, 16, 32)
    for i, (fov, th, ph, fb, fe) in enumerate(small["rsph/cases"]):
        fb, fe = int(fb), int(fe)
        nf = fe - fb
        vm = maps.view_map(fov, th, ph, 10, 6, 16, 32)
        got = gather_nearest(_planes(rp5), vm.idx_read[None], [(0, fb)], nf)
        assert np.array_equal(got.reshape(1, 2, nf, 6, 10), small[f"rsph/get_nearest/{i}"])
        view_in = randn(500 + i, 1, 2, nf, 6, 10).reshape(1, 2, nf, 60)
        ftab = maps.frame_table(7, [fb], nf)
        lat = _planes(rp5)
        merge_nearest([lat], [view_in], *vm.write_plan, ftab, 1, nf, 60)
        assert np.array_equal(lat.reshape(rp5.shape), small[f"rsph/set_nearest/{i}"])
        lat = _planes(rp5)
        merge_bilinear([lat], [view_in], maps.bilinear_merge_plan([vm]), ftab, 1, nf, 60)
        assert np.array_equal(lat.reshape(rp5.shape), small[f"rsph/set_bilinear/{i}"])


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
@pytest.mark.parametrize("world", [1, 2, 3])
def test_batched_plan_equals_reference_order(small, mode, world):
    """merge_all plan == the reference sets applied view after view, also when
    the job list is sharded over ranks and the disjoint partial results are summed."""
    lat0 = randn(51, 1, 2, 3, 16, 32)
    views = [(ph, th) for ph, th in small["step/views"]]
    vms = [maps.view_map(120, th, ph, 10, 6, 16, 32) for ph, th in views]
    wins = small[f"step/batched_{mode}/windows"]                     # [n, 1, 2, 3, 6, 10]
    x_prev = np.stack([wins[n] * np.float32(0.5) + randn(700 + n, 1, 2, 3, 6, 10) for n in range(len(views))])
    x0 = np.stack([wins[n] * np.float32(0.25) - randn(800 + n, 1, 2, 3, 6, 10) for n in range(len(views))])
    pay = [x_prev.reshape(len(views), 2, 3, 60), x0.reshape(len(views), 2, 3, 60)]
    plan = maps.nearest_merge_plan(vms) if mode == "nearest" else maps.bilinear_merge_plan(vms)
    sums = [np.zeros((2, 3, 512), np.float32) for _ in range(2)]
    weight = np.zeros(512, np.float32)
    for r in range(world):
        j0, nj = shard_range(len(views), r, world)
        part = [np.zeros((2, 3, 512), np.float32) for _ in range(2)]
        wpart = np.zeros(512, np.float32)
        local = [p[j0:j0 + nj] for p in pay]
        if mode == "nearest":
            merge_nearest(part, local, plan[0], plan[1], None, len(views), 3, 60, wpart, 1, j0, nj)
        else:
            merge_bilinear(part, local, plan, None, len(views), 3, 60, wpart, 1, j0, nj)
        for q in range(2):
            sums[q] += part[q]
        weight += wpart
    assert weight.max() == 1.0                                        # exactly one owner per pixel
    lat, den = _planes(lat0), np.zeros((2, 3, 512), np.float32)
    lat[:, :, weight > 0] = sums[0][:, :, weight > 0]
    den[:, :, weight > 0] = sums[1][:, :, weight > 0]
    tol = 0 if mode == "nearest" else 1e-5
    assert np.abs(lat.reshape(lat0.shape) - small[f"step/batched_{mode}/latent"]).max() <= tol
    assert np.abs(den.reshape(lat0.shape) - small[f"step/batched_{mode}/denoised"]).max() <= tol
    assert np.array_equal(weight.reshape(1, 16, 32), small[f"step/batched_{mode}/mask"])


@pytest.mark.parametrize("mode", ["nearest", "bilinear", "bilinear_tiled"])
@pytest.mark.parametrize("world", [1, 2])
def test_pixel_major_plan_equals_plane_major(mode, world):
    """The scatter-form plans of the pixel-major kernels (maps.pm_*_plan) give the same
    panorama as the plane-major plans (pinned to the golden above), with wrapping frame
    windows and a sharded job list; the lists are sorted by view with a view_ptr offset table."""
    from _emulate import pm_merge_bilinear, pm_merge_bilinear_tiled, pm_merge_nearest, pm_pack, pm_unpack
    H, W, F, h, w, f_len = 16, 32, 6, 6, 10, 4
    views = po.default_sphere_views(3)[:9]
    fbs = [0, 2, 4]                                                    # last window wraps: frames 4,5,0,1
    vms = [maps.view_map(120, th, ph, w, h, H, W) for ph, th in views]
    nj = len(views) * len(fbs)
    pay = [randn(900 + q, nj, 4, f_len, h * w) for q in range(2)]
    ftab = maps.frame_table(F, fbs, f_len)
    base = randn(5, 4, F, H * W)
    want = [base.copy(), base.copy()]
    wmask = np.zeros((4, F, H * W), np.float32)
    got = [pm_pack(base), pm_pack(base)]
    gmask = np.zeros((H * W, F, 4), np.float32)
    pm_plan = {"nearest": maps.pm_nearest_plan, "bilinear": maps.pm_bilinear_plan,
               "bilinear_tiled": maps.pm_bilinear_tiled_plan}[mode](vms)
    assert pm_plan is not None
    assert pm_plan[0][-1] == len(pm_plan[1]) and np.all(np.diff(pm_plan[0]) >= 0)
    for r in range(world):
        j0, n = shard_range(nj, r, world)
        local = [p[j0:j0 + n] for p in pay]
        if mode == "nearest":
            d, s_ = maps.nearest_merge_plan(vms)
            merge_nearest(want, local, d, s_, ftab, len(views), f_len, h * w, wmask, 2, j0, n)
            pm_merge_nearest(got, local, pm_plan, ftab, fbs, len(views), f_len, h * w, gmask, 2, j0, n)
        else:
            merge_bilinear(want, local, maps.bilinear_merge_plan(vms), ftab, len(views), f_len, h * w, wmask, 2, j0, n)
            emu = pm_merge_bilinear if mode == "bilinear" else pm_merge_bilinear_tiled
            emu(got, local, pm_plan, ftab, fbs, len(views), f_len, h * w, gmask, 2, j0, n)
    for q in range(2):
        assert np.array_equal(pm_unpack(got[q]), want[q])
    assert np.array_equal(pm_unpack(gmask), wmask)
    assert 0 < wmask.mean() < 1


def test_ring_merge_plan_emulation_matches_oracle():
    H, W, h, w = 16, 32, 6, 12
    windows = [(4 * a + 2, 3 * b + 1) for a in range(8) for b in range(5)]
    lat = randn(1, 1, 2, 3, H, W)
    pays = [randn(100 + n, 1, 2, 3, h, w) for n in range(len(windows))]
    ref = lat.copy()
    for n, (l, t) in enumerate(windows):
        po.ring_set(ref, pays[n], l, l + w, t, t + h)
    plan = ring_merge_plan(windows, H, W, h, w)
    got = _planes(lat)
    merge_nearest([got], [np.stack(pays).reshape(len(windows), 2, 3, h * w)], None, plan, None, len(windows), 3, h * w)
    assert np.array_equal(got.reshape(lat.shape), ref)


def test_frame_table_last_window_wins():
    ftab = maps.frame_table(8, [6, 1], 4)            # windows [6,7,0,1] then [1,2,3,4]
    assert ftab.tolist() == [2, 4, 5, 6, 7, -1, 0, 1]


def test_shard_range_partitions():
    for n in (1, 5, 44, 88, 572):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_loop_geometry_matches_oracle():
    """dynamicscaler_b200.geometry (what bench.py and callers use) vs the oracle's restatement
    of the reference loops, for every step phase."""
    from dynamicscaler_b200 import geometry as geo
    for i in range(0, 33):
        assert geo.sphere_views(i) == po.default_sphere_views(i)
        assert geo.ring_windows(i, 16, 16, 8192, 4096) == po.ring_window_grid(i, 16, 16, 8192, 4096)
        assert geo.ring_windows(i, 2, 2, 1024, 512) == po.ring_window_grid(i, 2, 2, 1024, 512)
        assert geo.ring_windows(i, 4, 4, 2048, 1024) == po.ring_window_grid(i, 4, 4, 2048, 1024)
        assert geo.frame_windows(i) == po.frame_window_begins(i)
        assert geo.frame_windows(i, overlap=0.5) == po.frame_window_begins(i, overlap=0.5)
    assert len(geo.sphere_views(0)) == 44 and len(geo.frame_windows(0)) == 13


@pytest.mark.parametrize("mode", ["nearest", "bilinear"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_round_robin_view_sharding_partitions_the_global_plan(mode, world):
    """Views dealt round-robin over ranks (step.plan_sphere): the ranks' filtered plans have
    disjoint support and, merged rank after rank in ANY order, reproduce the global merge."""
    from _emulate import merge_bilinear as emu_b, merge_nearest as emu_n
    H, W, F, h, w = 16, 32, 2, 6, 10
    views = po.default_sphere_views(4)[:11]
    nv, hw = len(views), h * w
    vms = [maps.view_map(120, th, ph, w, h, H, W) for ph, th in views]
    pay = [randn(40 + q, nv, 4, F, hw) for q in range(2)]
    base = randn(6, 4, F, H * W)
    want = [base.copy(), base.copy()]
    whole = maps.nearest_merge_plan(vms) if mode == "nearest" else maps.bilinear_merge_plan(vms)
    if mode == "nearest":
        emu_n(want, pay, whole[0], whole[1], None, nv, F, hw)
    else:
        emu_b(want, pay, whole, None, nv, F, hw)
    got = [base.copy(), base.copy()]
    seen = np.zeros(H * W, np.int32)
    for r in reversed(range(world)):                                   # order must not matter
        local = list(range(r, nv, world))
        if not local:
            continue
        loc_pay = [p[local] for p in pay]
        if mode == "nearest":
            d, s_ = maps.filter_nearest_plan(whole, hw, nv, local)
            emu_n(got, loc_pay, d, s_, None, len(local), F, hw)
            seen[d] += 1
            pm = maps.pm_nearest_plan(vms, (d, s_), len(local))
            assert pm[0][-1] == d.size and np.array_equal(np.sort(pm[2]), np.sort(d))
        else:
            part = maps.filter_bilinear_plan(whole, hw, nv, local)
            emu_b(got, loc_pay, part, None, len(local), F, hw)
            seen[part[0]] += 1
            tiled = maps.pm_bilinear_tiled_plan(vms, part, len(local))
            assert tiled is not None and tiled[1][:, 2].sum() == part[0].size
    assert seen.max() == 1 and seen.sum() == whole[0].size              # disjoint, complete
    for q in range(2):
        assert np.array_equal(got[q], want[q])


def test_ddim_coefficients_match_oracle_scalars():
    """step.ddim_coefficients (torch float32 ops, product side) == oracle.ddim_scalars (numpy float32),
    bit for bit, on the reference schedule entries stored with the golden vectors."""
    import os
    from dynamicscaler_b200.step import ddim_coefficients
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ddim.npz"))
    for eta, index, gs, a_t, a_prev, sig, s1m in g["cases"]:
        sa, s1, sp, sd, sg, gg = ddim_coefficients(a_t, a_prev, sig, s1m, gs)
        want = po.ddim_scalars(a_t, a_prev, sig, s1m)
        got = np.array([sa, s1, sp, sd, sg], np.float32)
        assert np.array_equal(got, np.array(want, np.float32)), (index, got, want)
        assert np.float32(gg) == np.float32(gs)


def test_wrap_slices_property():
    """get_dimension_slices_and_sizes (shift_window_utils.py:14-38) for random legal windows: the
    slices concatenate to the modular index vector (begin + j) mod size and agree with
    the oracle's segments."""
    from hypothesis import given, settings, strategies as stn

    @settings(max_examples=300, deadline=None)
    @given(stn.integers(1, 64).flatmap(lambda size: stn.tuples(
        stn.just(size), stn.integers(0, 2 * size - 1)).flatmap(lambda sb: stn.tuples(
            stn.just(sb[0]), stn.just(sb[1]), stn.integers(sb[1] + 1, 2 * sb[0])))))
    def check(args):
        size, begin, end = args
        slices, sizes = get_dimension_slices_and_sizes(begin, end, size)
        idx = np.concatenate([np.arange(s.start, s.stop) for s in slices])
        assert np.array_equal(idx, (begin + np.arange(end - begin)) % size)
        assert sizes == [s.stop - s.start for s in slices] and sum(sizes) == end - begin
        assert [(s.start, s.stop - s.start) for s in slices] == po.ring_segments(begin, end, size)

    check()


def test_random_view_lists_merge_plans_property():
    """Random ordered view lists -- angles outside [0, 360), views on the poles and across the
    seam, narrow and wide FOVs, odd view sizes -- merged through the product's host plans (plane-major,
    pixel-major and tiled, kernels emulated) equal the oracle's set_view_* calls applied one view
    after the other, and the gather maps equal the oracle's get_view_* calls."""
    from hypothesis import given, settings, strategies as stn
    from _emulate import pm_merge_bilinear, pm_merge_bilinear_tiled, pm_merge_nearest, pm_pack, pm_unpack

    view = stn.tuples(stn.sampled_from([-90.0, -60.0, -33.5, 0.0, 12.25, 45.0, 90.0]),           # phi
                      stn.one_of(stn.sampled_from([-370.0, -180.0, 0.0, 179.5, 355.0, 360.0, 725.0]),
                                 stn.floats(-400, 800, allow_nan=False, width=32)))               # theta
    case = stn.tuples(stn.lists(view, min_size=1, max_size=4), stn.sampled_from([50, 90, 120, 150]),
                      stn.sampled_from([(8, 16), (16, 32)]), stn.integers(2, 7), stn.integers(2, 9),
                      stn.integers(0, 10 ** 6))

    @settings(max_examples=80, deadline=None)
    @given(case)
    def check(args):
        views, fov, (H, W), h, w, seed = args
        P, nv, hw = 4, len(views), h * w
        vms = [maps.view_map(fov, th, ph, w, h, H, W) for ph, th in views]
        pay = randn(seed, nv, P, 1, hw)
        base = randn(seed + 1, P, 1, H * W)
        # gathers
        pano = base.reshape(P, H, W)
        for n, (ph, th) in enumerate(views):
            got = gather_nearest(base, np.stack([vm.idx_read for vm in vms]), [(n, 0)], 1)[0, :, 0]
            assert np.array_equal(got.reshape(P, h, w), po.sphere_get_nearest(pano, fov, th, ph, w, h)[0])
            idx4, wt4 = vms[n].read_taps()
            got = gather_bilinear(base, idx4[None], wt4[None], [(0, 0)], 1)[0, :, 0]
            assert np.array_equal(got.reshape(P, h, w), po.sphere_get_bilinear(pano, fov, th, ph, w, h))
        # merges
        want_n, want_b = pano.copy(), pano.copy()
        for n, (ph, th) in enumerate(views):
            po.sphere_set_nearest(want_n, pay[n, :, 0].reshape(P, h, w), fov, th, ph)
            po.sphere_set_bilinear(want_b, pay[n, :, 0].reshape(P, h, w), fov, th, ph)
        got = base.copy()
        d, s_ = maps.nearest_merge_plan(vms)
        merge_nearest([got], [pay], d, s_, None, nv, 1, hw)
        assert np.array_equal(got.reshape(P, H, W), want_n)
        got = pm_pack(base)
        pm_merge_nearest([got], [pay], maps.pm_nearest_plan(vms), None, None, nv, 1, hw)
        assert np.array_equal(pm_unpack(got).reshape(P, H, W), want_n)
        got = base.copy()
        merge_bilinear([got], [pay], maps.bilinear_merge_plan(vms), None, nv, 1, hw)
        assert np.array_equal(got.reshape(P, H, W), want_b)
        got = pm_pack(base)
        pm_merge_bilinear([got], [pay], maps.pm_bilinear_plan(vms), None, None, nv, 1, hw)
        assert np.array_equal(pm_unpack(got).reshape(P, H, W), want_b)
        tiled = maps.pm_bilinear_tiled_plan(vms)
        if tiled is not None:                      # None: a row with > 64 entries (generic kernel instead)
            _, tile, row_dst, row_start, ent_k, ent_w = tiled
            assert tile[:, 2].max(initial=0) <= maps.PM_TILE_ROWS and tile[:, 3].max(initial=0) <= maps.PM_TILE_ENTRIES
            assert tile[:, 2].sum() == row_dst.size and (ent_k >= 0).sum() == tile[:, 3].sum()
            got = pm_pack(base)
            pm_merge_bilinear_tiled([got], [pay], tiled, None, None, nv, 1, hw)
            assert np.array_equal(pm_unpack(got).reshape(P, H, W), want_b)

    check()


def test_chunk_sizes_match_torch_chunk():
    """vae_tiling._chunk_sizes mirrors torch.chunk (i2v_sphere_panorama_pipeline.py:977,990 rely on its
    uneven-tail behaviour): chunk size, number of chunks, size of the last one."""
    import torch
    from dynamicscaler_b200.vae_tiling import _chunk_sizes
    for n in (1, 5, 16, 17, 36, 40, 64, 100, 128, 255, 256, 1024):
        for chunks in (1, 2, 8, 16, 18, 33):
            parts = torch.chunk(torch.arange(n), chunks)
            size, count, last = _chunk_sizes(n, chunks)
            assert (size, count, last) == (parts[0].numel(), len(parts), parts[-1].numel()), (n, chunks)


def test_necessary_bytes_accounting():
    """necessary_bytes (the honest roofline numerator): a gather writes every window and reads each touched
    panorama row once; a merge never needs more payload than the SURVEY formula charges, and exactly the rows the
    plan owns are written."""
    from dynamicscaler_b200.geometry import frame_windows, sphere_views
    from dynamicscaler_b200.step import PanoramaStepper
    for (H, W, F, kw) in ((128, 256, 16, {}), (64, 128, 32, dict(frame_begins=frame_windows(1, total_f=32), frames=16))):
        st = PanoramaStepper(H, W, F, device="cpu")
        for mode in ("nearest", "bilinear"):
            plan = st.plan_sphere(sphere_views(2), 64, 40, mode=mode, **kw)
            ga, ma = st.algorithmic_bytes(plan, 4, 4, k=2)
            gn, mn, detail = st.necessary_bytes(plan, 4, 4, k=2)
            n_win_bytes = plan.n_jobs * 4 * plan.f_len * plan.hw * 4
            assert gn >= n_win_bytes and gn <= ga * 1.25
            assert detail["payload_bytes"] <= 2 * n_win_bytes            # never more payload than exists
            assert mn <= ma * 1.05                                        # pruned views only ever remove bytes (+ plan arrays)
            assert detail["rows_written"] == plan.extra["host_plan"][0].size
            assert detail["rows_written"] == int(round(plan.coverage * H * W))


def test_shift_window_schedule_matches_oracle():
    """loop.shift_window_schedule (windows x frame_begins, the product plan_ring takes) == the reference's triple
    loop restated in oracle.shift_window_jobs (pipeline/t2v_normal_pipeline.py:420-524) for every step, incl. the
    odd-step jumps (crossed w/h names kept), single-window axes and docking on any axis."""
    import itertools
    from dynamicscaler_b200.loop import shift_window_schedule
    from oracle import panorama_oracle as po
    cases = 0
    for nf, nh, nw in ((1, 1, 1), (2, 2, 3), (1, 3, 1), (3, 1, 2)):
        for jump, dock in itertools.product(((False, False, False), (True, False, True), (False, True, False)),
                                            ((False, False, False), (True, False, False), (False, True, True))):
            kw = dict(frames=8, latent_height=8, latent_width=12, num_windows_f=nf, num_windows_h=nh, num_windows_w=nw,
                      loop_step=4, shift_jump_odd_w=jump[0], shift_jump_odd_h=jump[1], shift_jump_odd_f=jump[2],
                      docking_w=dock[0], docking_h=dock[1], docking_f=dock[2], docking_step_range=[0, 3, 5])
            for i in range(9):
                windows, fbs = shift_window_schedule(i, **kw)
                got = [(l, t, fb) for fb in fbs for (l, t) in windows]
                assert got == po.shift_window_jobs(i, **kw), (nf, nh, nw, jump, dock, i)
                cases += 1
    assert cases == 4 * 9 * 9


def test_i2v_planar_schedule_matches_oracle():
    """loop.i2v_planar_schedule == the reference's f / w / h loops restated in oracle.i2v_planar_jobs
    (pipeline/i2v_sphere_panorama_pipeline.py:732-855): overlapping strides with round(), begin_index_offset,
    frame windows with and without docking (docked windows first / last, skipped when the shift is 0, windows past
    the end dropped)."""
    from dynamicscaler_b200.loop import i2v_planar_schedule
    from oracle import panorama_oracle as po
    n = 0
    spatial = ((128, 64, 384, 128, 3, 3, 4), (96, 64, 240, 128, 3, 3, 4), (512, 320, 8192, 4096, 16, 16, 16),
               (512, 320, 1024, 512, 4, 3, 8))                   # (width, height, total_w, total_h, nw, nh, loop_step)
    for (width, height, tw, th, nw, nh, loop_step) in spatial:
        for (frames, total_f, ov, lsf) in ((4, 8, 0.5, 2), (16, 64, 0.75, 8), (16, 16, 0.75, 8), (16, 32, 0.25, 4)):
            for dock in (False, True):
                kw = dict(width=width, height=height, frames=frames, total_w=tw, total_h=th, total_f=total_f,
                          num_windows_w=nw, num_windows_h=nh, loop_step=loop_step, overlap_ratio_f=ov, loop_step_frame=lsf,
                          dock_at_f=dock, begin_index_offset=n % 3)
                for i in range(10):
                    windows, fbs = i2v_planar_schedule(i, **kw)
                    assert [(l, t, fb) for fb in fbs for (l, t) in windows] == po.i2v_planar_jobs(i, **kw), (kw, i)
                    n += 1
    assert n == 4 * 4 * 2 * 10


def test_t2v_planar_schedule_matches_oracle():
    """loop.t2v_planar_schedule == the reference's f / w / h loops restated in oracle.t2v_planar_jobs
    (pipeline/t2v_sphere_panorama_pipeline.py:436-535): integer strides, per-step offsets, shifted frame windows,
    docking at the top / bottom edge (skipped when the shift is 0, windows past the bottom edge dropped)."""
    from dynamicscaler_b200.loop import t2v_planar_schedule
    from oracle import panorama_oracle as po
    n = 0
    for (width, height, tw, th, nw, nh, ls) in ((96, 64, 240, 128, 3, 3, 4), (128, 64, 384, 128, 3, 3, 4),
                                                (512, 320, 8192, 4096, 16, 16, 16), (512, 320, 1024, 512, 4, 3, 8)):
        for nf, frames in ((1, 16), (2, 4), (4, 16)):
            if nf > 1 and frames // ls == 0:
                continue
            for dock in (False, True):
                kw = dict(width=width, height=height, frames=frames, total_w=tw, total_h=th, num_windows_w=nw,
                          num_windows_h=nh, num_windows_f=nf, loop_step=ls, dock_at_h=dock)
                for i in range(10):
                    windows, fbs = t2v_planar_schedule(i, **kw)
                    assert [(l, t, fb) for fb in fbs for (l, t) in windows] == po.t2v_planar_jobs(i, **kw), (kw, i)
                    n += 1
    assert n == 200
