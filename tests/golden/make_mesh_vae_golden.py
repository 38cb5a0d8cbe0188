"""Generate tests/golden/mesh_vae_ref.npz by RUNNING the reference's own body decoder on CPU:

  `ConvDecoder` is executed from /root/reference/ca_code/models/mesh_vae.py by slicing the class out of the file's AST
  at run time (the module itself cannot be imported: drtk / pytorch3d are absent), built from the reference's own
  `ca_code.nn.layers` / `ca_code.nn.blocks` — nothing is copied into this repository — at the mesh_vae_example.yml
  configuration (uv 1024, 58.7 M parameters), with the asset-dependent pieces (masks, seam sampler, geo_fn.from_uv)
  replaced by the same stand-ins oracle/mesh_vae_oracle.py offers.  The restatement gets the same seeded parameters and
  inputs; the script asserts that the two agree and freezes (a) the reference's state-dict keys and shapes, (b) summary
  statistics and sampled values of every output of the REFERENCE run.

Needs /root/reference; the .npz is committed.   Usage: python tests/golden/make_mesh_vae_golden.py
"""
import ast
import json
import logging
import os
import sys
import time
import types
import warnings
from typing import Dict
from unittest.mock import MagicMock

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "mesh_vae_ref.npz")
SRC = "/root/reference/ca_code/models/mesh_vae.py"
OUTPUTS = ("geom_delta_rec", "geom_uv_delta_rec", "tex_mean_rec", "embs_conv", "pose_conv")


def summarise(t):
    """what the fixture keeps of one output tensor: statistics + 256 evenly spaced samples"""
    f = t.detach().double().reshape(-1)
    idx = th.linspace(0, f.numel() - 1, 256).long()
    return np.concatenate([[f.mean().item(), f.std().item(), f.abs().max().item(), float(f.numel())], f[idx].numpy()])


def main():
    if not os.path.isfile(SRC):
        sys.exit("needs /root/reference (build container only)")
    sys.path.insert(0, "/root/reference")
    sys.path.insert(0, ROOT)
    warnings.filterwarnings("ignore")
    sys.modules.setdefault("turtle", MagicMock())  # blocks.py:8 imports it by accident; tkinter is absent here
    import ca_code.nn.layers as la
    from ca_code.nn.blocks import ConvBlock, UpConvBlockDeep, tile2d
    from oracle import mesh_vae_oracle as mo

    src = open(SRC).read()
    cls_src = next(ast.get_source_segment(src, n) for n in ast.parse(src).body
                   if isinstance(n, ast.ClassDef) and n.name == "ConvDecoder")
    ns = {"th": th, "nn": th.nn, "np": np, "la": la, "ConvBlock": ConvBlock, "UpConvBlockDeep": UpConvBlockDeep,
          "tile2d": tile2d, "Dict": Dict, "logger": logging.getLogger("mesh_vae")}
    exec(cls_src, ns)

    masks = mo.synthetic_masks()
    from_uv = mo.uv_vertex_gather()
    cfg = dict(uv_size=1024, init_uv_size=64, n_pose_dims=98, n_pose_enc_channels=16, n_embs=1024,
               n_embs_enc_channels=32, n_face_embs=256, n_init_channels=64, n_min_channels=4)  # mesh_vae_example.yml:25-34
    ref = ns["ConvDecoder"](geo_fn=types.SimpleNamespace(from_uv=from_uv), seam_sampler=types.SimpleNamespace(
        impaint=mo.identity_resample, resample=mo.identity_resample), assets=types.SimpleNamespace(**masks), **cfg)
    ours = mo.ConvDecoder(masks, mo.identity_resample, from_uv, **cfg)

    keys_ref = {k: list(v.shape) for k, v in ref.state_dict().items()}
    keys_ours = {k: list(v.shape) for k, v in ours.state_dict().items()}
    assert keys_ref == keys_ours, set(keys_ref.items()) ^ set(keys_ours.items())
    n_params = sum(p.numel() for p in ref.parameters())
    mo.seeded_fill(ref)
    mo.seeded_fill(ours)
    for (ka, a), (kb, b) in zip(sorted(ref.named_parameters()), sorted(ours.named_parameters())):
        assert ka == kb and th.equal(a, b), ka
    ours.load_state_dict(ref.state_dict())  # and the reference's own state_dict loads

    pose, embs, face = mo.seeded_inputs()
    th.set_num_threads(os.cpu_count() or 1)
    with th.no_grad():
        t0 = time.perf_counter()
        r = ref(pose.clone(), embs.clone(), face.clone())
        t_ref = time.perf_counter() - t0
        t0 = time.perf_counter()
        o = ours(pose.clone(), embs.clone(), face.clone())
        t_ours = time.perf_counter() - t0
    d = {"keys_json": np.frombuffer(json.dumps(keys_ref, sort_keys=True).encode(), dtype=np.uint8),
         "n_params": np.int64(n_params)}
    for k in OUTPUTS:
        err = (r[k] - o[k]).abs().max().item() / max(r[k].abs().max().item(), 1e-30)
        assert r[k].shape == o[k].shape and err < 2e-5, (k, err)
        d["out_" + k] = summarise(r[k])
        d["shape_" + k] = np.array(r[k].shape)
        print("%-18s shape %-22s max rel err restatement vs reference %.2e" % (k, tuple(r[k].shape), err))
    np.savez_compressed(OUT, **d)
    print("reference ConvDecoder: %.1f M parameters, forward %.2f s (restatement %.2f s) on %d threads -> %s"
          % (n_params / 1e6, t_ref, t_ours, th.get_num_threads(), OUT))


if __name__ == "__main__":
    main()
