"""Host-side mirror of the reference wrapper extensions/mvpraymarch/mvpraymarch.py: `build_accel` (:21-84),
`MVPRaymarch` (:87-311) and `mvpraymarch` (:314-418) with the same names, keyword arguments (ca_code/utils/
render_raymarcher.py:59 introspects `mvpraymarch.__code__.co_varnames`), asserts and gradient wiring, calling the
sm_100a kernels through `mvpraymarchlib`.

Differences that do not change results: the fixed-order heap topology (a function of N, K only) is cached per
device instead of being rebuilt with torch ops on every call; bounds scratch is a cached workspace."""
import torch
from torch.autograd import Function

from . import mvpraymarchlib

_TOPOLOGY = {}


def _fixedorder_topology(N, K, dev):
    """sortedobjid [N,K] = k; implicit heap over 2K-1 nodes: children (2i+1, 2i+2) for internal i < K-1 (the values
    stored for leaves are never read), parent floor((i-1)/2), root -1 (mvpraymarch.py:44-73)."""
    key = (N, K, dev.index)
    hit = _TOPOLOGY.get(key)
    if hit is None:
        sortedobjid = (torch.arange(N * K, dtype=torch.int32, device=dev) % K).view(N, K)
        nodechildren = (
            torch.cat([
                torch.arange(1, (K - 1) * 2 + 1, dtype=torch.int32, device=dev),
                torch.div(torch.arange(-2, -(K * 2 + 1) - 1, -1, dtype=torch.int32, device=dev), 2, rounding_mode="floor"),
            ], dim=0).view(1, K + K - 1, 2).repeat(N, 1, 1))
        nodeparent = (torch.div(torch.arange(-1, K * 2 - 2, dtype=torch.int32, device=dev), 2, rounding_mode="floor")
                      .view(1, -1).repeat(N, 1))
        hit = (sortedobjid.contiguous(), nodechildren.contiguous(), nodeparent.contiguous())
        if len(_TOPOLOGY) > 16:
            _TOPOLOGY.clear()
        _TOPOLOGY[key] = hit
    return hit


def build_accel(primtransfin, algo, fixedorder=False):
    """Bounds of the acceleration structure for the primitives (mvpraymarch.py:21-84).  Only the fixed-order heap is
    a functional path of the reference."""
    primpos, primrot, primscale = primtransfin
    N, K = primpos.size(0), primpos.size(1)
    dev = primpos.device
    if not fixedorder:
        raise NotImplementedError("usebvh=True builds a tree the reference traversal never reads; use 'fixedorder'")
    sortedobjid, nodechildren, nodeparent = _fixedorder_topology(N, K, dev)
    nodeaabb = torch.empty((N, K + K - 1, 2, 3), dtype=torch.float32, device=dev)
    mvpraymarchlib.compute_aabb(*primtransfin, sortedobjid, nodechildren, nodeparent, nodeaabb, algo)
    return sortedobjid, nodechildren, nodeaabb


class MVPRaymarch(Function):
    """Custom Function for raymarching Mixture of Volumetric Primitives."""

    @staticmethod
    def forward(self, raypos, raydir, stepsize, tminmax, primpos, primrot, primscale, template, warp, rayterm, gradmode,
                options):
        algo = options["algo"]
        usebvh = options["usebvh"]
        chlast = options["chlast"]
        with_shadow = options["with_shadow"]
        if isinstance(options["blocksize"], tuple):
            blocksizex, blocksizey = options["blocksize"]
        else:
            blocksizex, blocksizey = options["blocksize"], 1

        assert raypos.is_contiguous() and raypos.size(3) == 3
        assert raydir.is_contiguous() and raydir.size(3) == 3
        assert tminmax.is_contiguous() and tminmax.size(3) == 2
        assert primpos is None or primpos.is_contiguous() and primpos.size(2) == 3
        assert primrot is None or primrot.is_contiguous() and primrot.size(2) == 3
        assert primscale is None or primscale.is_contiguous() and primscale.size(2) == 3
        if chlast:
            assert template.is_contiguous()
            assert len(template.size()) == 6
            assert template.size(-1) == 4
            assert warp is None or (warp.is_contiguous() and warp.size(-1) == 3)
        else:
            assert template.is_contiguous() and len(template.size()) == 6 and template.size(2) == 4
            assert warp is None or (warp.is_contiguous() and warp.size(2) == 3)

        primtransfin = (primpos, primrot, primscale)
        if usebvh is not False:
            sortedobjid, nodechildren, nodeaabb = build_accel(primtransfin, algo, fixedorder=usebvh == "fixedorder")
        else:
            sortedobjid, nodechildren, nodeaabb = None, None, None

        N, H, W = raypos.size(0), raypos.size(1), raypos.size(2)
        rayrgba = torch.empty((N, H, W, 4), device=raypos.device)
        # the kernel writes raysat for every ray, so no fill(-1) pass is needed (reference: mvpraymarch.py:150)
        raysat = torch.empty((N, H, W, 3), dtype=torch.float32, device=raypos.device) if gradmode else None
        rayterm = None

        shadow = None
        if with_shadow:
            if chlast:
                Ns, P, D, Hs, Ws, C = template.shape
                shadow = torch.zeros((Ns, P, D, Hs, Ws, 2), device=template.device, dtype=torch.float32)
            else:
                Ns, P, C, D, Hs, Ws = template.shape
                shadow = torch.zeros((Ns, P, 2, D, Hs, Ws), device=template.device, dtype=torch.float32)

        mvpraymarchlib.raymarch_forward(
            raypos, raydir, stepsize, tminmax, sortedobjid, nodechildren, nodeaabb, *primtransfin, template, warp, rayrgba,
            raysat, rayterm, shadow, algo, options["sortprims"], options["maxhitboxes"], options["synchitboxes"], chlast,
            options["fadescale"], options["fadeexp"], options["accum"], options["termthresh"], options["griddim"],
            blocksizex, blocksizey)

        self.save_for_backward(raypos, raydir, tminmax, sortedobjid, nodechildren, nodeaabb, primpos, primrot, primscale,
                               template, warp, rayrgba, raysat, rayterm)
        self.options = options
        self.stepsize = stepsize
        return rayrgba, shadow

    @staticmethod
    def backward(self, grad_rayrgba, grad_shadow):
        (raypos, raydir, tminmax, sortedobjid, nodechildren, nodeaabb, primpos, primrot, primscale, template, warp, rayrgba,
         raysat, rayterm) = self.saved_tensors
        o = self.options
        if isinstance(o["bwdblocksize"], tuple):
            blocksizex, blocksizey = o["bwdblocksize"]
        else:
            blocksizex, blocksizey = o["bwdblocksize"], 1

        grad_primpos = torch.zeros_like(primpos)
        grad_primrot = torch.zeros_like(primrot)
        grad_primscale = torch.zeros_like(primscale)
        primtransfin = (primpos, grad_primpos, primrot, grad_primrot, primscale, grad_primscale)
        grad_template = torch.zeros_like(template)
        grad_warp = torch.zeros_like(warp) if warp is not None else None

        mvpraymarchlib.raymarch_backward(
            raypos, raydir, self.stepsize, tminmax, sortedobjid, nodechildren, nodeaabb, *primtransfin, template,
            grad_template, warp, grad_warp, rayrgba, grad_rayrgba.contiguous(), raysat, rayterm, o["algo"], o["sortprims"],
            o["maxhitboxes"], o["synchitboxes"], o["chlast"], o["fadescale"], o["fadeexp"], o["accum"], o["termthresh"],
            o["griddim"], blocksizex, blocksizey)

        return (None, None, None, None, grad_primpos, grad_primrot, grad_primscale, grad_template, grad_warp, None, None,
                None)


def mvpraymarch(
    raypos,
    raydir,
    stepsize,
    tminmax,
    primtransf,
    template,
    warp,
    rayterm=None,
    algo=0,
    usebvh="fixedorder",
    sortprims=False,
    randomorder=False,
    maxhitboxes=512,
    synchitboxes=True,
    chlast=True,
    fadescale=8.0,
    fadeexp=8.0,
    accum=2,
    termthresh=0.99,
    griddim=3,
    blocksize=(8, 16),
    bwdblocksize=(8, 16),
    with_shadow=False,
):
    """Main entry point for raymarching MVP (argument meaning: extensions/mvpraymarch/mvpraymarch.py:341-378).

    raypos, raydir [N,H,W,3]; tminmax [N,H,W,2]; primtransf = (primpos [N,K,3], primrot [N,K,3,3], primscale [N,K,3])
    or a packed [N,K,5,3] tensor; template [N,K,TD,TH,TW,4] (chlast); warp [N,K,WD,WH,WW,3] or None.
    algo 0 = no warp field, 1 = warp field.  Returns rayrgba [N,H,W,4] (and the shadow volume when with_shadow)."""
    if isinstance(primtransf, tuple):
        primpos, primrot, primscale = primtransf
    else:
        primpos, primrot, primscale = (primtransf[:, :, 0, :].contiguous(), primtransf[:, :, 1:4, :].contiguous(),
                                       primtransf[:, :, 4, :].contiguous())
    primtransfin = (primpos, primrot, primscale)
    out, shadow = MVPRaymarch.apply(
        raypos, raydir, stepsize, tminmax, *primtransfin, template, warp, rayterm, torch.is_grad_enabled(),
        {"algo": algo, "usebvh": usebvh, "sortprims": sortprims, "randomorder": randomorder, "maxhitboxes": maxhitboxes,
         "synchitboxes": synchitboxes, "chlast": chlast, "fadescale": fadescale, "fadeexp": fadeexp, "accum": accum,
         "termthresh": termthresh, "griddim": griddim, "blocksize": blocksize, "bwdblocksize": bwdblocksize,
         "with_shadow": with_shadow})
    if with_shadow:
        assert shadow is not None
        if chlast:
            shadow = shadow[..., 0:1] / (shadow[..., 1:] + 1e-5)
        else:
            shadow = shadow[:, :, 0:1] / (shadow[:, :, 1:] + 1e-5)
        return out, shadow
    return out
