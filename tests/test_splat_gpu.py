"""GPU parity of the splat path (project -> bin/sort -> blend, fwd + bwd) against the CPU oracle, through the
C ABI (ctypes) and through the gsplat-compatible Python surface.

Bars: bit-exact for projection outputs that feed binning, for intersection keys, sorted ids and tile bins;
1e-4 relative for pixels and gradients (the blend uses the hardware ex2 like the reference's __expf)."""
import numpy as np
import pytest
import torch

from util import assert_close, small_scene, t2n

pytestmark = pytest.mark.gpu

CASES = {
    # name: (kwargs of small_scene, block_width, scale multiplier)
    "dense96": (dict(G=3000, img_h=96, img_w=80), 16, 12.0),
    "ragged": (dict(G=2500, img_h=70, img_w=93, seed=11, cam=3), 16, 15.0),
    "bw8": (dict(G=1500, img_h=64, img_w=48, seed=5), 8, 10.0),
    "ties": (dict(G=4000, img_h=96, img_w=80, depth_quant=True), 16, 8.0),
    "tiny_prims": (dict(G=5000, img_h=128, img_w=96, seed=9), 16, 1.0),
}


def _dev(d, dev):
    return {k: (torch.from_numpy(v).to(dev) if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def _project_gpu(s, dev, bw, mult):
    from goliath_b200.gsplat import project_gaussians

    t = _dev(s, dev)
    return project_gaussians(t["means3d"], t["scales"] * mult, 1.0, t["quats"], t["viewmat"], s["fx"], s["fy"],
                             s["cx"], s["cy"], s["img_h"], s["img_w"], bw, 0.1)


def _project_cpu(orc, s, bw, mult):
    return orc.project_fwd(s["means3d"], s["scales"] * np.float32(mult), 1.0, s["quats"], s["viewmat"], s["fx"],
                           s["fy"], s["cx"], s["cy"], s["img_h"], s["img_w"], bw, 0.1)


@pytest.mark.parametrize("case", list(CASES))
def test_project_forward_bit_exact(orc, cuda, case):
    kw, bw, mult = CASES[case]
    s = small_scene(**kw)
    xys, depths, radii, conics, comp, nth, cov3d = _project_gpu(s, cuda, bw, mult)
    ref = _project_cpu(orc, s, bw, mult)
    assert (ref["radii"] > 0).sum() > 100
    for name, got in (("xys", xys), ("depths", depths), ("radii", radii), ("conics", conics),
                      ("compensation", comp), ("num_tiles_hit", nth), ("cov3d", cov3d)):
        g = t2n(got)
        r = ref[name]
        assert g.dtype == r.dtype and g.shape == r.shape, name
        assert np.array_equal(g.view(np.uint32) if g.dtype == np.float32 else g,
                              r.view(np.uint32) if r.dtype == np.float32 else r), "%s not bit-exact" % name


@pytest.mark.parametrize("case", list(CASES))
def test_binning_bit_exact(orc, cuda, case):
    from goliath_b200.gsplat import utils as gu

    kw, bw, mult = CASES[case]
    s = small_scene(**kw)
    ref_p = _project_cpu(orc, s, bw, mult)
    ref = orc.bin_and_sort(ref_p["xys"], ref_p["depths"], ref_p["radii"], ref_p["num_tiles_hit"], s["img_h"],
                           s["img_w"], bw)
    d = lambda a: torch.from_numpy(a).to(cuda)
    n, cum = gu.compute_cumulative_intersects(d(ref_p["num_tiles_hit"]))
    assert n == ref["num_intersects"]
    assert np.array_equal(t2n(cum), ref["cum_tiles_hit"])
    tb = gu._tile_bounds(s["img_h"], s["img_w"], bw)
    isect, gids, isect_s, gids_s, bins = gu.bin_and_sort_gaussians(
        len(ref_p["radii"]), n, d(ref_p["xys"]), d(ref_p["depths"]), d(ref_p["radii"]), cum, tb, bw)
    assert np.array_equal(t2n(isect), ref["isect_ids"])
    assert np.array_equal(t2n(gids), ref["gaussian_ids"])
    assert np.array_equal(t2n(isect_s), ref["isect_ids_sorted"])
    assert np.array_equal(t2n(gids_s), ref["gaussian_ids_sorted"])
    assert np.array_equal(t2n(bins), ref["tile_bins"])


@pytest.mark.parametrize("n", [1, 31, 255, 256, 257, 4095, 4096, 4097, 100_003, 1_500_000])
def test_radix_sort_sizes(cuda, n):
    """Stable sort contract at block-boundary sizes, full 64-bit-wide tile ids, and uniform digits."""
    from goliath_b200.gsplat import utils as gu

    g = torch.Generator().manual_seed(n)
    tiles = torch.randint(0, 2688, (n,), generator=g, dtype=torch.int64)
    depth = (900.0 + 200.0 * torch.rand(n, generator=g)).to(torch.float32)
    depth[::7] = 1000.0  # many exact ties
    keys = (tiles << 32) | depth.view(torch.int32).to(torch.int64)
    vals = torch.arange(n, dtype=torch.int32)
    ks, vs = gu.sort_intersects(keys.to(cuda), vals.to(cuda), 2688)
    order = np.argsort(keys.numpy(), kind="stable")
    assert np.array_equal(t2n(ks), keys.numpy()[order])
    assert np.array_equal(t2n(vs), vals.numpy()[order])


@pytest.mark.parametrize("n", [1, 5, 2047, 2048, 2049, 300_000, 1_048_576])
def test_cumsum(cuda, n):
    from goliath_b200.gsplat import utils as gu

    x = torch.randint(0, 9, (n,), generator=torch.Generator().manual_seed(n), dtype=torch.int32)
    tot, cum = gu.compute_cumulative_intersects(x.to(cuda))
    ref = np.cumsum(x.numpy(), dtype=np.int64)
    assert tot == ref[-1] and np.array_equal(t2n(cum), ref.astype(np.int32))


def _blend_inputs(orc, s, bw, mult, C, rng):
    p = _project_cpu(orc, s, bw, mult)
    b = orc.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], s["img_h"], s["img_w"], bw)
    G = len(p["radii"])
    colors = rng.random((G, C)).astype(np.float32)
    if C == 4:
        colors[:, 3] = p["depths"]
    bg = rng.random(C).astype(np.float32)
    opac = (s["opacity"][:, 0] * p["compensation"]).astype(np.float32)
    return p, b, colors, bg, opac


@pytest.mark.parametrize("C", [3, 4])
@pytest.mark.parametrize("case", list(CASES))
def test_blend_forward_backward(orc, cuda, case, C):
    from goliath_b200 import _lib

    kw, bw, mult = CASES[case]
    s = small_scene(**kw)
    H, W = s["img_h"], s["img_w"]
    rng = np.random.default_rng(42)
    p, b, colors, bg, opac = _blend_inputs(orc, s, bw, mult, C, rng)
    out_r, Ts_r, fi_r = orc.rasterize_fwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"],
                                          colors, opac, bg)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    gids, bins, xys, conics, col, op, bgd = (d(b["gaussian_ids_sorted"]), d(b["tile_bins"]), d(p["xys"]),
                                             d(p["conics"]), d(colors), d(opac), d(bg))
    out = torch.empty(H, W, C, device=cuda)
    Ts = torch.empty(H, W, device=cuda)
    fi = torch.empty(H, W, device=cuda, dtype=torch.int32)
    L = _lib.lib()
    st = _lib.stream_ptr(cuda)
    _lib.check(L.gb_rasterize_fwd(H, W, bw, C, gids.data_ptr(), bins.data_ptr(), xys.data_ptr(), conics.data_ptr(),
                                  col.data_ptr(), op.data_ptr(), bgd.data_ptr(), out.data_ptr(), Ts.data_ptr(),
                                  fi.data_ptr(), st), "fwd")
    torch.cuda.synchronize()
    assert (Ts_r < 0.9).mean() > 0.02, "scene must cover pixels"
    # a borderline alpha<1/255 or T<=1e-4 decision can flip on a handful of pixels (ex2 vs libm expf)
    assert_close(t2n(out), out_r, rtol=1e-4, atol=2e-5, frac=0.9995, what="out_img")
    assert_close(t2n(Ts), Ts_r, rtol=1e-4, atol=2e-6, frac=0.9995, what="final_Ts")
    assert (t2n(fi) == fi_r).mean() >= 0.999

    # backward on the ORACLE's forward state so both sides differentiate the same blend
    v_out = rng.standard_normal((H, W, C)).astype(np.float32)
    v_alpha = rng.standard_normal((H, W)).astype(np.float32)
    ref = orc.rasterize_bwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], colors, opac,
                            bg, Ts_r, fi_r, v_out, v_alpha)
    G = len(opac)
    v_xy, v_conic = torch.zeros(G, 2, device=cuda), torch.zeros(G, 3, device=cuda)
    v_col, v_op = torch.zeros(G, C, device=cuda), torch.zeros(G, 1, device=cuda)
    Ts_d, fi_d, vo_d, va_d = d(Ts_r), d(fi_r), d(v_out), d(v_alpha)  # keep alive: raw pointers cross the C ABI
    _lib.check(L.gb_rasterize_bwd(H, W, bw, C, gids.data_ptr(), bins.data_ptr(), xys.data_ptr(), conics.data_ptr(),
                                  col.data_ptr(), op.data_ptr(), bgd.data_ptr(), Ts_d.data_ptr(),
                                  fi_d.data_ptr(), vo_d.data_ptr(), va_d.data_ptr(), v_xy.data_ptr(),
                                  v_conic.data_ptr(), v_col.data_ptr(), v_op.data_ptr(), st), "bwd")
    torch.cuda.synchronize()
    for name, got, want in (("v_xy", v_xy, ref[0]), ("v_conic", v_conic, ref[1]), ("v_colors", v_col, ref[2]),
                            ("v_opacity", v_op, ref[3])):
        assert_close(t2n(got), want, rtol=1e-4, atol=1e-4 * np.abs(want).max() * 0.1, frac=0.999, what=name)


@pytest.mark.parametrize("case", ["dense96", "ragged"])
def test_project_backward(orc, cuda, case):
    kw, bw, mult = CASES[case]
    s = small_scene(**kw)
    t = _dev(s, cuda)
    from goliath_b200.gsplat import project_gaussians

    means, scales, quats = (t["means3d"].requires_grad_(), (t["scales"] * mult).requires_grad_(),
                            t["quats"].requires_grad_())
    xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(
        means, scales, 1.0, quats, t["viewmat"], s["fx"], s["fy"], s["cx"], s["cy"], s["img_h"], s["img_w"], bw, 0.1)
    G = means.shape[0]
    rng = np.random.default_rng(3)
    v = [rng.standard_normal(sh).astype(np.float32) for sh in ((G, 2), (G,), (G, 3), (G,))]
    d = lambda a: torch.from_numpy(a).to(cuda)
    torch.autograd.backward([xys, depths, conics, comp], [d(v[0]), d(v[1]), d(v[2]), d(v[3])])
    p = _project_cpu(orc, s, bw, mult)
    ref = orc.project_bwd(s["means3d"], s["scales"] * np.float32(mult), 1.0, s["quats"], s["viewmat"], s["fx"],
                          s["fy"], p["cov3d"], p["radii"], p["conics"], p["compensation"], *v)
    assert_close(t2n(means.grad), ref["v_mean3d"], rtol=1e-5, what="v_mean3d")
    assert_close(t2n(scales.grad), ref["v_scale"], rtol=1e-5, what="v_scale")
    assert_close(t2n(quats.grad), ref["v_quat"], rtol=1e-5, what="v_quat")


def test_render_like_reference_wrapper(orc, cuda):
    """The exact call sequence of ca_code/utils/render_gsplat.py:41-106 (project, rasterise rgb, rasterise depth),
    forward + backward through autograd, against the oracle chain."""
    from goliath_b200.gsplat import project_gaussians, rasterize_gaussians

    kw, bw, mult = CASES["dense96"]
    s = small_scene(**kw)
    H, W = s["img_h"], s["img_w"]
    t = _dev(s, cuda)
    means, scales, quats = t["means3d"].requires_grad_(), (t["scales"] * mult).requires_grad_(), t["quats"].requires_grad_()
    opacity, colors = t["opacity"].requires_grad_(), t["colors"].requires_grad_()
    bg = torch.zeros(3, device=cuda)
    xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(
        means, scales, 1.0, quats, t["viewmat"], s["fx"], s["fy"], s["cx"], s["cy"], H, W, bw, 0.1)
    out_img, alpha = rasterize_gaussians(xys, depths, radii, conics, nth, colors, opacity * comp[:, None], H, W, bw, bg,
                                         return_alpha=True)
    out_depth = rasterize_gaussians(xys, depths, radii, conics, nth, depths[:, None].expand(-1, 3).contiguous(),
                                    opacity * comp[:, None], H, W, bw, bg, return_alpha=True)[0]
    rng = np.random.default_rng(5)
    w_img = torch.from_numpy(rng.standard_normal((H, W, 3)).astype(np.float32)).to(cuda)
    w_dep = torch.from_numpy((1e-3 * rng.standard_normal((H, W))).astype(np.float32)).to(cuda)
    ((out_img * w_img).sum() + (out_depth[..., 0] * w_dep).sum()).backward()

    # oracle chain
    p = _project_cpu(orc, s, bw, mult)
    b = orc.bin_and_sort(p["xys"], p["depths"], p["radii"], p["num_tiles_hit"], H, W, bw)
    opac = (s["opacity"][:, 0] * p["compensation"]).astype(np.float32)
    z3 = np.zeros(3, np.float32)
    o1, T1, f1 = orc.rasterize_fwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"],
                                   s["colors"], opac, z3)
    dcol = np.repeat(p["depths"][:, None], 3, 1)
    o2, T2, f2 = orc.rasterize_fwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], dcol,
                                   opac, z3)
    assert_close(t2n(out_img), o1, rtol=1e-4, atol=2e-5, frac=0.9995, what="rgb")
    assert_close(t2n(alpha), 1 - T1, rtol=1e-4, atol=2e-5, frac=0.9995, what="alpha")
    assert_close(t2n(out_depth), o2, rtol=1e-4, atol=2e-2, frac=0.9995, what="depth")
    za = np.zeros((H, W), np.float32)
    wd3 = np.zeros((H, W, 3), np.float32)
    wd3[..., 0] = t2n(w_dep)
    g1 = orc.rasterize_bwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], s["colors"],
                           opac, z3, T1, f1, t2n(w_img), za)
    g2 = orc.rasterize_bwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], dcol, opac, z3,
                           T2, f2, wd3, za)
    v_xy, v_conic = g1[0] + g2[0], g1[1] + g2[1]
    v_opac_eff = (g1[3] + g2[3])[:, 0]
    v_depth = g2[2].sum(1)
    v_comp = v_opac_eff * s["opacity"][:, 0]
    ref = orc.project_bwd(s["means3d"], s["scales"] * np.float32(mult), 1.0, s["quats"], s["viewmat"], s["fx"],
                          s["fy"], p["cov3d"], p["radii"], p["conics"], p["compensation"], v_xy, v_depth, v_conic,
                          v_comp)
    tol = dict(rtol=2e-4, frac=0.999)
    assert_close(t2n(colors.grad), g1[2], what="d colors", **tol)
    assert_close(t2n(opacity.grad)[:, 0], v_opac_eff * p["compensation"], what="d opacity", **tol)
    assert_close(t2n(means.grad), ref["v_mean3d"], what="d means", **tol)
    assert_close(t2n(scales.grad), ref["v_scale"], what="d scales", **tol)
    assert_close(t2n(quats.grad), ref["v_quat"], what="d quats", **tol)


def test_empty_and_culled(cuda):
    from goliath_b200.gsplat import project_gaussians, rasterize_gaussians

    H, W = 48, 40
    V = torch.eye(4, device=cuda)[:3].contiguous()
    means = torch.tensor([[0.0, 0.0, -5.0], [0.0, 0.0, 0.01]], device=cuda)  # behind / inside near plane
    scales = torch.full((2, 3), 0.1, device=cuda)
    quats = torch.tensor([[1.0, 0, 0, 0]] * 2, device=cuda)
    xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(means, scales, 1.0, quats, V, 50.0, 50.0, 20.0,
                                                                     24.0, H, W, 16, 0.1)
    assert int(nth.sum()) == 0 and int(radii.sum()) == 0
    bg = torch.tensor([0.25, 0.5, 0.75], device=cuda)
    img, alpha = rasterize_gaussians(xys, depths, radii, conics, nth, torch.rand(2, 3, device=cuda),
                                     torch.rand(2, 1, device=cuda), H, W, 16, bg, return_alpha=True)
    assert torch.allclose(img, bg.expand(H, W, 3))
    # reference quirk kept: with zero intersections final_Ts is zeros, so alpha == 1 (gsplat 0.1.11 rasterize.py)
    assert torch.all(alpha == 1)


def test_product_refuses_cpu_tensors(cuda):
    from goliath_b200.gsplat import project_gaussians

    with pytest.raises(RuntimeError):
        project_gaussians(torch.zeros(4, 3), torch.ones(4, 3), 1.0, torch.tensor([[1.0, 0, 0, 0]] * 4),
                          torch.eye(4)[:3].contiguous(), 10.0, 10.0, 8.0, 8.0, 16, 16, 16)


@pytest.fixture(params=["batch", "pipe", "affine", "mom", "mom-affine"])
def blend_mode(request):
    """The formulations of the packed blend: CTA-synchronous double buffer (csrc/splat_blend_packed.cu), warp-decoupled
    mbarrier pipeline (csrc/splat_blend_pipe.cu), the pipeline drawing its tiles from an SM-affine schedule, and the
    round-2 default: exact cull + 4-hit ILP forward + transposed (moment) backward (csrc/splat_blend_mom.cu)."""
    from goliath_b200 import _lib

    L = _lib.lib()
    before = L.gb_get_blend_mode()
    L.gb_set_blend_mode({"batch": 0, "pipe": 1, "affine": 2, "mom": 3, "mom-affine": 4}[request.param])
    yield request.param
    L.gb_set_blend_mode(before)


@pytest.mark.parametrize("C", [3, 4])
@pytest.mark.parametrize("case", ["dense96", "ragged", "ties", "tiny_prims", "big", "long_lists"])
def test_packed_blend_matches_generic_kernel(orc, cuda, case, C, blend_mode):
    """The bulk-copy/culling blend (both formulations) must give the generic kernel's pixels BIT-exactly
    (same per-pair arithmetic, culled pairs are exactly the alpha<1/255 ones) and its gradients to atomics order."""
    from goliath_b200 import _lib

    if case == "big":
        kw, bw, mult = dict(G=60000, img_h=300, img_w=260, seed=21), 16, 3.0
    elif case == "long_lists":  # few tiles, thousands of records each, mostly transparent: many pipeline laps per tile
        kw, bw, mult = dict(G=40000, img_h=48, img_w=40, seed=29), 16, 2.0
    else:
        kw, bw, mult = CASES[case]
    s = small_scene(**kw)
    H, W = s["img_h"], s["img_w"]
    rng = np.random.default_rng(7)
    p, b, colors, bg, opac = _blend_inputs(orc, s, bw, mult, C, rng)
    if case == "long_lists":
        opac[p["xys"][:, 0] >= 18.0] *= 0.05  # left part saturates early (warps leave the pipeline), right part never
    opac[::97] = 0.001   # below 1/255: culled everywhere
    opac[::89] = 1.0     # exercises the 0.99 / 0.999 clamps
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    gids, bins, xys, conics, col, op, bgd = (d(b["gaussian_ids_sorted"]), d(b["tile_bins"]), d(p["xys"]),
                                             d(p["conics"]), d(colors), d(opac), d(bg))
    n = len(b["gaussian_ids_sorted"])
    T = bins.shape[0]
    L = _lib.lib()
    st = _lib.stream_ptr(cuda)
    outs = []
    rec = torch.empty(n, 12, device=cuda)
    affine = blend_mode in ("affine", "mom-affine")
    order = torch.full((L.gb_tile_schedule_ints(T),), -1, dtype=torch.int32, device=cuda)
    _lib.check(L.gb_pack_records(n, C, gids.data_ptr(), xys.data_ptr(), conics.data_ptr(), col.data_ptr(),
                                 op.data_ptr(), rec.data_ptr(), st), "pack")
    _lib.check((L.gb_tile_schedule if affine else L.gb_tile_order)(T, bins.data_ptr(), order.data_ptr(), st), "order")
    ras_fwd = L.gb_rasterize_sched_fwd if affine else L.gb_rasterize_packed_fwd
    ras_bwd = L.gb_rasterize_sched_bwd if affine else L.gb_rasterize_packed_bwd
    torch.cuda.synchronize()
    o = t2n(order[:T])
    assert np.array_equal(np.sort(o), np.arange(T)), "tile order must be a permutation"
    if affine:
        assert bool((order[T:] == 0).all()), "draw counters of a fresh schedule"
    else:
        lens = (b["tile_bins"][:, 1] - b["tile_bins"][:, 0])[o]
        assert np.all(np.diff(np.minimum(lens >> 3, 1023)) <= 0), "longest lists first"
        assert bool((order[T:] == -1).all()), "gb_tile_order writes T entries"
    for packed in (False, True, True):  # the packed kernels twice: a schedule must be reusable launch after launch
        out = torch.empty(H, W, C, device=cuda)
        Ts = torch.empty(H, W, device=cuda)
        fi = torch.empty(H, W, device=cuda, dtype=torch.int32)
        if packed:
            _lib.check(ras_fwd(H, W, C, bins.data_ptr(), order.data_ptr(), rec.data_ptr(),
                               bgd.data_ptr(), out.data_ptr(), Ts.data_ptr(), fi.data_ptr(), st), "pf")
        else:
            _lib.check(L.gb_rasterize_fwd(H, W, bw, C, gids.data_ptr(), bins.data_ptr(), xys.data_ptr(),
                                          conics.data_ptr(), col.data_ptr(), op.data_ptr(), bgd.data_ptr(),
                                          out.data_ptr(), Ts.data_ptr(), fi.data_ptr(), st), "gf")
        torch.cuda.synchronize()
        outs.append((out, Ts, fi))
    for k in (1, 2):
        assert all(torch.equal(outs[0][i], outs[k][i]) for i in range(3)), "packed pixels != generic kernel (pass %d)" % k
    if affine:
        assert bool((order[T:] == 0).all()), "draw counters must be back to zero after a launch"
    v_out = d(rng.standard_normal((H, W, C)).astype(np.float32))
    v_alpha = d(rng.standard_normal((H, W)).astype(np.float32))
    G = len(opac)
    grads = []
    for packed in (False, True, True):
        g = [torch.zeros(G, 2, device=cuda), torch.zeros(G, 3, device=cuda), torch.zeros(G, C, device=cuda),
             torch.zeros(G, 1, device=cuda)]
        Ts, fi = outs[0][1], outs[0][2]
        if packed:
            _lib.check(ras_bwd(H, W, C, gids.data_ptr(), bins.data_ptr(), order.data_ptr(),
                               rec.data_ptr(), bgd.data_ptr(), Ts.data_ptr(), fi.data_ptr(),
                               v_out.data_ptr(), v_alpha.data_ptr(), g[0].data_ptr(), g[1].data_ptr(),
                               g[2].data_ptr(), g[3].data_ptr(), st), "pb")
        else:
            _lib.check(L.gb_rasterize_bwd(H, W, bw, C, gids.data_ptr(), bins.data_ptr(), xys.data_ptr(),
                                          conics.data_ptr(), col.data_ptr(), op.data_ptr(), bgd.data_ptr(),
                                          Ts.data_ptr(), fi.data_ptr(), v_out.data_ptr(), v_alpha.data_ptr(),
                                          g[0].data_ptr(), g[1].data_ptr(), g[2].data_ptr(), g[3].data_ptr(), st), "gb")
        torch.cuda.synchronize()
        grads.append([t2n(x) for x in g])
    for k in (1, 2):
        for name, x, y in zip(("v_xy", "v_conic", "v_colors", "v_opacity"), grads[k], grads[0]):
            # modes 0-2 differ from the generic kernel only in the order of the atomics; mode 3 re-associates the sums
            # over the pixels (image moments) and uses rcp.approx: same 1e-4 bar as against the oracle
            tol = dict(rtol=1e-4, atol=1e-5 * np.abs(y).max(), frac=0.999) if blend_mode.startswith("mom") else \
                dict(rtol=2e-5, atol=2e-5 * np.abs(y).max())
            assert_close(x, y, what=name + " packed vs generic (pass %d)" % k, **tol)
    # and against the oracle
    ref = orc.rasterize_bwd(H, W, bw, b["gaussian_ids_sorted"], b["tile_bins"], p["xys"], p["conics"], colors, opac,
                            bg, t2n(outs[0][1]), t2n(outs[0][2]), t2n(v_out), t2n(v_alpha))
    for name, x, y in zip(("v_xy", "v_conic", "v_colors", "v_opacity"), grads[1], ref):
        assert_close(x, y, rtol=1e-4, atol=1e-5 * np.abs(y).max(), frac=0.999, what=name + " packed vs oracle")


def test_fused_render_equals_two_pass_and_bin_cache(cuda):
    """goliath_b200.render.render: the fused 4-channel pass gives the two-pass pixels exactly; the second
    rasterisation of the two-pass form reuses the binning (no cumsum/sort launches) unless an input changed."""
    from goliath_b200 import _lib
    from goliath_b200.gsplat import rasterize as R
    from goliath_b200.render import render

    kw, bw, mult = CASES["dense96"]
    s = small_scene(**kw)
    t = _dev(s, cuda)
    args = (s["img_w"], s["img_h"], s["fx"], s["fy"], s["cx"], s["cy"], t["viewmat"], t["means3d"],
            t["quats"], t["scales"] * mult, t["opacity"], t["colors"])
    R.clear_bin_cache()
    L = _lib.lib()
    L.gb_launch_count_reset()
    a = render(*args, fused=False)
    two_pass_launches = L.gb_launch_count()
    R.clear_bin_cache()
    L.gb_launch_count_reset()
    b = render(*args, fused=True)
    fused_launches = L.gb_launch_count()
    for k in ("render", "alpha", "depth"):
        assert torch.equal(a[k], b[k]), k
    # one binning serves both passes of the two-pass form: only pack + blend are extra
    assert two_pass_launches - fused_launches == 2, (two_pass_launches, fused_launches)
    # gradients of the single-node fused render == gradients of the two-pass autograd chain (atomics order apart)
    grads = []
    for fused in (False, True):
        leaves = [t[k].clone().requires_grad_() for k in ("means3d", "quats", "scales", "opacity", "colors")]
        leaves[2] = (t["scales"] * mult).clone().requires_grad_()
        o = render(s["img_w"], s["img_h"], s["fx"], s["fy"], s["cx"], s["cy"], t["viewmat"], leaves[0], leaves[1],
                   leaves[2], leaves[3], leaves[4], fused=fused)
        w = torch.linspace(0.5, 1.5, s["img_h"] * s["img_w"], device=cuda).view(s["img_h"], s["img_w"])
        ((o["render"] * w).sum() + (o["depth"] * w * 1e-3).sum() + (o["alpha"] * w).sum()).backward()
        grads.append([t2n(x.grad) for x in leaves])
    for name, a_, b_ in zip(("means3d", "quats", "scales", "opacity", "colors"), grads[0], grads[1]):
        assert_close(b_, a_, rtol=1e-4, atol=2e-5 * np.abs(a_).max(), frac=0.999, what="fused vs two-pass grad " + name)
    # cache must miss after an in-place change of an input
    from goliath_b200.gsplat import project_gaussians, rasterize_gaussians
    xys, depths, radii, conics, comp, nth, cov3d = project_gaussians(
        t["means3d"], t["scales"] * mult, 1.0, t["quats"], t["viewmat"], s["fx"], s["fy"], s["cx"], s["cy"],
        s["img_h"], s["img_w"], bw, 0.1)
    bg = torch.zeros(3, device=cuda)
    i1 = rasterize_gaussians(xys, depths, radii, conics, nth, t["colors"], t["opacity"], s["img_h"], s["img_w"], bw, bg)
    depths.add_(5.0)  # bumps the version counter
    L.gb_launch_count_reset()
    i2 = rasterize_gaussians(xys, depths, radii, conics, nth, t["colors"], t["opacity"], s["img_h"], s["img_w"], bw, bg)
    assert L.gb_launch_count() > 6, "binning must be recomputed when depths changed"


def test_sync_free_render_and_cuda_graph(cuda, blend_mode):
    """capacity=N: no host sync, same pixels and gradients as the exact path; overflow is detected; the whole
    forward+backward can be captured in a CUDA graph and replayed on new inputs."""
    from goliath_b200.gsplat.fused import check_overflow
    from goliath_b200.render import render

    kw, bw, mult = CASES["dense96"]
    s = small_scene(**kw)
    t = _dev(s, cuda)
    H, W = s["img_h"], s["img_w"]

    def run(means, capacity):
        leaves = [means.clone().requires_grad_(), t["quats"].clone().requires_grad_(),
                  (t["scales"] * mult).clone().requires_grad_(), t["opacity"].clone().requires_grad_(),
                  t["colors"].clone().requires_grad_()]
        o = render(W, H, s["fx"], s["fy"], s["cx"], s["cy"], t["viewmat"], *leaves, capacity=capacity)
        (o["render"].sum() + 1e-3 * o["depth"].sum() + o["alpha"].sum()).backward()
        return o, [x.grad for x in leaves]

    o_ref, g_ref = run(t["means3d"], None)
    o_sf, g_sf = run(t["means3d"], 1 << 16)
    assert not check_overflow(cuda)
    for k in ("render", "depth", "alpha"):
        assert torch.equal(o_ref[k], o_sf[k]), k
    for a_, b_ in zip(g_ref, g_sf):
        assert_close(t2n(b_), t2n(a_), rtol=1e-4, atol=2e-5 * float(a_.abs().max()), frac=0.999, what="sync-free grads")
    # too small a capacity is reported, not silently ignored
    run(t["means3d"], 64)
    assert check_overflow(cuda) and not check_overflow(cuda)

    # CUDA graph: static input buffer, capture forward + backward, replay with other content
    static_means = t["means3d"].clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            run(static_means, 1 << 16)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o_g, g_g = run(static_means, 1 << 16)
    shifted = t["means3d"] + torch.tensor([3.0, -2.0, 1.0], device=cuda)
    static_means.copy_(shifted)
    graph.replay()
    torch.cuda.synchronize()
    o_e, g_e = run(shifted, None)
    assert torch.equal(o_g["render"], o_e["render"]) and torch.equal(o_g["alpha"], o_e["alpha"])
    assert_close(t2n(g_g[0]), t2n(g_e[0]), rtol=1e-4, atol=2e-5 * float(g_e[0].abs().max()), frac=0.999, what="graph grads")


def test_render_views_fused_matches_reference_sequence(cuda):
    """render_views(fused=True) — one rasteriser node + one post-processing kernel per view — against the reference's
    sequence (render_gsplat.render per view, stack, alpha from the detached final_T, depth / alpha.clamp(0.05, 1)),
    forward and gradients, two views."""
    from goliath_b200.render import render_views

    kw, bw, mult = CASES["dense96"]
    s = small_scene(**kw)
    t = _dev(s, cuda)
    Rt = torch.stack([t["viewmat"], t["viewmat"]])
    Rt[1, 0, 3] += 7.0  # second view shifted sideways
    intr = [(s["fx"], s["fy"], s["cx"], s["cy"])] * 2
    H, W = s["img_h"], s["img_w"]
    w = torch.linspace(0.5, 1.5, 2 * H * W, device=cuda).view(2, 1, H, W)
    res = []
    for fused in (False, True):
        leaves = dict(primpos=t["means3d"], primqvec=t["quats"], primscale=t["scales"] * mult, opacity=t["opacity"],
                      color=t["colors"])
        leaves = {k: v.clone()[None].expand(2, *v.shape).contiguous().requires_grad_() for k, v in leaves.items()}
        rgb, alpha, depth = render_views(W, H, None, Rt, leaves, intrinsics_host=intr, fused=fused)
        assert rgb.shape == (2, 3, H, W) and alpha.shape == (2, 1, H, W) and depth.shape == (2, 1, H, W)
        assert not alpha.requires_grad
        ((rgb * w).sum() + (depth * w * 1e-3).sum()).backward()
        res.append((t2n(rgb), t2n(alpha), t2n(depth), {k: t2n(v.grad) for k, v in leaves.items()}))
    (r0, a0, d0, g0), (r1, a1, d1, g1) = res
    assert np.array_equal(r0, r1) and np.array_equal(a0, a1)
    assert_close(d1, d0, rtol=1e-6, atol=1e-6, what="normalised depth")
    for k in g0:
        assert_close(g1[k], g0[k], rtol=1e-4, atol=2e-5 * np.abs(g0[k]).max(), frac=0.999, what="grad " + k)


@pytest.fixture(params=["buckets+split", "coop+split", "passes+fused", "buckets+fused"])
def rank_sort(request):
    """variants of the bucket binning: depth ranks by 2048 key buckets + in-bucket ranking (cooperative LSD fallback for
    degenerate depth distributions), as one cooperative LSD kernel over the varying key bits (default), or as the four
    radix passes of round 1; per-tile ordering as bitmap sort + grid-wide record gather (default) or as the
    single per-tile kernel of round 1"""
    from goliath_b200 import _lib

    L = _lib.lib()
    before = (L.gb_get_rank_sort_mode(), L.gb_get_tile_sort_mode())
    rs, ts = request.param.split("+")
    L.gb_set_rank_sort_mode({"coop": 0, "passes": 1, "buckets": 2}[rs])
    L.gb_set_tile_sort_mode({"split": 0, "fused": 1}[ts])
    yield request.param
    L.gb_set_rank_sort_mode(before[0])
    L.gb_set_tile_sort_mode(before[1])


@pytest.mark.parametrize("case", ["dense96", "ragged", "bw8", "ties", "tiny_prims", "one_cta", "two_ctas", "big",
                                  "many_tiles", "too_many_tiles", "sixteen_per_thread", "flat_depth", "wide_depth"])
def test_bin_tiles_matches_key_sort(cuda, case, rank_sort):
    """The bucket binning of the fused render (depth ranks + per-tile bitmap sort, csrc/splat_bin_tiles.cu) returns
    bit for bit the bins, sorted Gaussian ids and blend records of the key-sort path (csrc/splat_bin.cu + pack)."""
    from goliath_b200 import _lib
    from goliath_b200.gsplat import utils as gu

    extra = {
        "one_cta": (dict(G=4096, img_h=96, img_w=80, seed=3), 16, 10.0),
        "two_ctas": (dict(G=4097, img_h=96, img_w=80, seed=4, depth_quant=True), 16, 10.0),
        "big": (dict(G=150_000, img_h=512, img_w=384, seed=13), 16, 6.0),
        "many_tiles": (dict(G=30_000, img_h=768, img_w=1024, seed=17), 8, 2.0),        # 12288 tiles: opt-in smem
        "too_many_tiles": (dict(G=20_000, img_h=600, img_w=640, seed=19), 4, 1.5),     # 24000 tiles: global atomics
        "sixteen_per_thread": (dict(G=400_000, img_h=256, img_w=192, seed=23), 16, 1.5),
        "flat_depth": (dict(G=6000, img_h=96, img_w=80, seed=31, cam=0), 16, 10.0),   # every depth key identical: 0 passes
        "wide_depth": (dict(G=9000, img_h=96, img_w=80, seed=37, cam=0), 16, 10.0),   # depths over 12 octaves: 4 passes
    }
    kw, bw, mult = CASES[case] if case in CASES else extra[case]
    s = small_scene(**kw)
    if case == "flat_depth":    # camera 0 looks down the world z axis: depth = 1000 - z
        s["means3d"][:, 2] = 0.0
    if case == "wide_depth":
        rng = np.random.default_rng(37)
        s["means3d"][:, 2] = (1000.0 - np.exp2(rng.uniform(-2.0, 10.0, size=len(s["means3d"])))).astype(np.float32)
        s["means3d"][:, :2] *= 0.05
    H, W = s["img_h"], s["img_w"]
    xys, depths, radii, conics, comp, nth, cov3d = _project_gpu(s, cuda, bw, mult)
    t = _dev(s, cuda)
    colors, opacity = t["colors"].contiguous(), t["opacity"].contiguous()
    G = xys.shape[0]
    L = _lib.lib()
    st = _lib.stream_ptr(cuda)
    # key-sort path
    n, cum = gu.compute_cumulative_intersects(nth)
    assert n > 1000
    tb = gu._tile_bounds(H, W, bw)
    T = tb[0] * tb[1]
    _, _, _, gids_ref, bins_ref = gu.bin_and_sort_gaussians(G, n, xys, depths, radii, cum, tb, bw)
    rec_ref = torch.empty(n, 12, device=cuda)
    _lib.check(L.gb_pack_records_fused(n, gids_ref.data_ptr(), xys.data_ptr(), conics.data_ptr(), colors.data_ptr(),
                                       depths.data_ptr(), opacity.data_ptr(), comp.data_ptr(), rec_ref.data_ptr(), st),
               "pack")
    # bucket path, twice (the workspace is reused: stale counters must not leak), then with too small a capacity
    assert L.gb_bin_tiles_supported(G) == 1
    for cap in (n + 77, n + 77, n, n // 2):
        ws = torch.empty(L.gb_bin_tiles_workspace_bytes(G, T, cap), dtype=torch.uint8, device=cuda)
        bins = torch.full((T, 2), -7, dtype=torch.int32, device=cuda)
        order = torch.full((T,), -7, dtype=torch.int32, device=cuda)
        gids = torch.full((cap,), -7, dtype=torch.int32, device=cuda)
        rec = torch.full((cap, 12), float("nan"), device=cuda)
        n_out = torch.zeros(1, dtype=torch.int32, device=cuda)
        ovf = torch.zeros(1, dtype=torch.int32, device=cuda)
        _lib.check(L.gb_bin_tiles_pack(G, xys.data_ptr(), depths.data_ptr(), radii.data_ptr(), conics.data_ptr(),
                                       colors.data_ptr(), opacity.data_ptr(), comp.data_ptr(), H, W, bw, cap,
                                       bins.data_ptr(), order.data_ptr(), 0, gids.data_ptr(), rec.data_ptr(),
                                       n_out.data_ptr(), ovf.data_ptr(), ws.data_ptr(), st), "bin_tiles_pack")
        torch.cuda.synchronize()
        assert int(n_out) == n
        assert sorted(t2n(order).tolist()) == list(range(T))
        if cap >= n:
            assert int(ovf) == 0
            assert torch.equal(bins, bins_ref)
            assert torch.equal(gids[:n], gids_ref)
            assert torch.equal(rec[:n].view(torch.int32), rec_ref.view(torch.int32))
            assert bool((gids[n:] == -7).all())
        else:
            assert int(ovf) == 1
            b = t2n(bins).astype(np.int64)
            br = np.minimum(t2n(bins_ref).astype(np.int64), cap)
            br[br[:, 1] <= br[:, 0]] = 0
            assert np.array_equal(b, br)
            # every surviving bucket is a depth-ordered SUBSET of the reference list of its tile
            g, gr = t2n(gids), t2n(gids_ref)
            full = t2n(bins_ref)
            for tile in np.nonzero(b[:, 1] > b[:, 0])[0][:50]:
                mine = g[b[tile, 0]:b[tile, 1]]
                ref_list = gr[full[tile, 0]:full[tile, 1]]
                pos = {v: i for i, v in enumerate(ref_list.tolist())}
                idx = [pos[v] for v in mine.tolist()]
                assert idx == sorted(idx) and len(set(idx)) == len(idx)


@pytest.mark.parametrize("case", ["dense96", "ragged", "ties", "big", "flat_depth_like"])
@pytest.mark.parametrize("channels", [3, 4])
def test_ranked_binning_and_blend_match_packed(cuda, case, channels):
    """gb_bin_tiles_ranked + gb_rasterize_ranked_fwd/bwd (records staged by depth rank from the per-Gaussian table,
    no sorted-record gather) against gb_bin_tiles_pack + gb_rasterize_packed_fwd/bwd: rank_to_gid[ranks_sorted] and
    rec_by_rank[ranks_sorted] are the sorted ids / records bit for bit, images / final_T / final_idx identical, gradients
    equal up to the order of the atomic adds."""
    from goliath_b200 import _lib
    from goliath_b200.gsplat import utils as gu

    extra = {"big": (dict(G=150_000, img_h=512, img_w=384, seed=13), 16, 6.0),
             "flat_depth_like": (dict(G=6000, img_h=96, img_w=80, seed=31, cam=0), 16, 10.0)}
    kw, bw, mult = CASES[case] if case in CASES else extra[case]
    if bw != 16:
        pytest.skip("the blend kernels are block_width 16")
    s = small_scene(**kw)
    H, W = s["img_h"], s["img_w"]
    xys, depths, radii, conics, comp, nth, cov3d = _project_gpu(s, cuda, bw, mult)
    t = _dev(s, cuda)
    colors, opacity = t["colors"].contiguous(), t["opacity"].contiguous()
    G = xys.shape[0]
    L = _lib.lib()
    st = _lib.stream_ptr(cuda)
    n, _ = gu.compute_cumulative_intersects(nth)
    tb = gu._tile_bounds(H, W, bw)
    T = tb[0] * tb[1]
    cap = n + 33
    i32 = dict(dtype=torch.int32, device=cuda)
    ws = torch.empty(L.gb_bin_tiles_workspace_bytes(G, T, cap), dtype=torch.uint8, device=cuda)
    ovf = torch.zeros(1, **i32)
    # packed
    bins, order = torch.empty(T, 2, **i32), torch.empty(T, **i32)
    gids, rec = torch.empty(cap, **i32), torch.empty(cap, 12, device=cuda)
    _lib.check(L.gb_bin_tiles_pack(G, xys.data_ptr(), depths.data_ptr(), radii.data_ptr(), conics.data_ptr(),
                                   colors.data_ptr(), opacity.data_ptr(), comp.data_ptr(), H, W, bw, cap, bins.data_ptr(),
                                   order.data_ptr(), 0, gids.data_ptr(), rec.data_ptr(), None, ovf.data_ptr(),
                                   ws.data_ptr(), st), "bin_tiles_pack")
    # ranked (twice: the caller's arrays are rewritten, the shared workspace is reused)
    for _ in range(2):
        bins2, order2 = torch.empty(T, 2, **i32), torch.empty(T, **i32)
        ranks = torch.full((cap,), -7, **i32)
        rbr = torch.full((G, 12), float("nan"), device=cuda)
        r2g = torch.full((G,), -7, **i32)
        _lib.check(L.gb_bin_tiles_ranked(G, xys.data_ptr(), depths.data_ptr(), radii.data_ptr(), conics.data_ptr(),
                                         colors.data_ptr(), opacity.data_ptr(), comp.data_ptr(), H, W, bw, cap,
                                         bins2.data_ptr(), order2.data_ptr(), 0, ranks.data_ptr(), rbr.data_ptr(),
                                         r2g.data_ptr(), None, ovf.data_ptr(), ws.data_ptr(), None, st), "bin_tiles_ranked")
    torch.cuda.synchronize()
    assert int(ovf) == 0 and torch.equal(bins, bins2)
    assert sorted(t2n(order2).tolist()) == list(range(T))  # equal lengths may come in any order
    rk = ranks[:n].long()
    assert bool((rk >= 0).all()) and bool((ranks[n:] == -7).all())
    assert torch.equal(r2g[rk], gids[:n])
    assert torch.equal(rbr[rk].view(torch.int32), rec[:n].view(torch.int32))
    # blend
    C = channels
    if C == 3:  # 3-channel records: the colour quarter's 4th float is ignored
        pass
    bg = torch.rand(C, device=cuda)
    outs = []
    v_out = torch.randn(H, W, C, device=cuda)
    v_alpha = torch.randn(H, W, device=cuda)
    for ranked in (False, True):
        out = torch.empty(H, W, C, device=cuda)
        Ts = torch.empty(H, W, device=cuda)
        fi = torch.empty(H, W, **i32)
        gx, gc, gcol, go = (torch.zeros(G, 2, device=cuda), torch.zeros(G, 3, device=cuda),
                            torch.zeros(G, C, device=cuda), torch.zeros(G, 1, device=cuda))
        if ranked:
            _lib.check(L.gb_rasterize_ranked_fwd(H, W, C, bins.data_ptr(), order.data_ptr(), ranks.data_ptr(), rbr.data_ptr(),
                                                 bg.data_ptr(), out.data_ptr(), Ts.data_ptr(), fi.data_ptr(), st), "ranked fwd")
            _lib.check(L.gb_rasterize_ranked_bwd(H, W, C, r2g.data_ptr(), ranks.data_ptr(), bins.data_ptr(), order.data_ptr(),
                                                 rbr.data_ptr(), bg.data_ptr(), Ts.data_ptr(), fi.data_ptr(),
                                                 v_out.data_ptr(), v_alpha.data_ptr(), gx.data_ptr(), gc.data_ptr(),
                                                 gcol.data_ptr(), go.data_ptr(), st), "ranked bwd")
        else:
            _lib.check(L.gb_rasterize_packed_fwd(H, W, C, bins.data_ptr(), order.data_ptr(), rec.data_ptr(), bg.data_ptr(),
                                                 out.data_ptr(), Ts.data_ptr(), fi.data_ptr(), st), "packed fwd")
            _lib.check(L.gb_rasterize_packed_bwd(H, W, C, gids.data_ptr(), bins.data_ptr(), order.data_ptr(), rec.data_ptr(),
                                                 bg.data_ptr(), Ts.data_ptr(), fi.data_ptr(), v_out.data_ptr(),
                                                 v_alpha.data_ptr(), gx.data_ptr(), gc.data_ptr(), gcol.data_ptr(),
                                                 go.data_ptr(), st), "packed bwd")
        torch.cuda.synchronize()
        outs.append((t2n(out), t2n(Ts), t2n(fi), [t2n(g) for g in (gx, gc, gcol, go)]))
    (o0, T0, f0, g0), (o1, T1, f1, g1) = outs
    assert np.array_equal(o0, o1) and np.array_equal(T0, T1) and np.array_equal(f0, f1)
    for a, b, name in zip(g1, g0, ("v_xy", "v_conic", "v_colors", "v_opacity")):
        assert_close(a, b, rtol=1e-4, atol=1e-5 * float(np.abs(b).max()), frac=0.9999, what=name)
