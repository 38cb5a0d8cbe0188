"""CPU: the restatement of the reference's body decoder (oracle/mesh_vae_oracle.py, SURVEY.md §8 row R9 / BASELINE
configs[0]) against the fixture frozen from the reference's own `mesh_vae.ConvDecoder`
(tests/golden/make_mesh_vae_golden.py -> tests/golden/mesh_vae_ref.npz)."""
import json
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mesh_vae_ref.npz")
OUTPUTS = ("geom_delta_rec", "geom_uv_delta_rec", "tex_mean_rec", "embs_conv", "pose_conv")


def test_conv_decoder_matches_the_reference_fixture():
    from oracle import mesh_vae_oracle as mo

    g = np.load(GOLD)
    dec = mo.ConvDecoder(mo.synthetic_masks(), mo.identity_resample, mo.uv_vertex_gather())
    # the reference's checkpoint layout: same keys, same shapes, 58.7 M parameters
    keys = json.loads(bytes(g["keys_json"]).decode())
    assert {k: list(v.shape) for k, v in dec.state_dict().items()} == keys
    assert sum(p.numel() for p in dec.parameters()) == int(g["n_params"]) == 58_739_420
    mo.seeded_fill(dec)
    pose, embs, face = mo.seeded_inputs()
    with torch.no_grad():
        out = dec(pose, embs, face)
    for k in OUTPUTS:
        t = out[k].double().reshape(-1)
        ref = g["out_" + k]
        assert list(out[k].shape) == g["shape_" + k].tolist(), k
        assert int(ref[3]) == t.numel()
        got = np.concatenate([[t.mean().item(), t.std().item(), t.abs().max().item()],
                              t[torch.linspace(0, t.numel() - 1, 256).long()].numpy()])
        want = np.concatenate([ref[:3], ref[4:]])
        scale = ref[2]  # max |value| of the reference output
        assert np.all(np.abs(got - want) <= 1e-4 * scale + 1e-5 * np.abs(want)), (k, np.abs(got - want).max(), scale)


def test_conv_decoder_masks_and_scales():
    """Behaviour the reference's forward has beyond the conv stack (mesh_vae.py:577-630): the head region is removed
    from the pose branch, outputs carry verts_scale / tex_scale, the pose mask is integer."""
    from oracle import mesh_vae_oracle as mo

    masks = mo.synthetic_masks()
    dec = mo.seeded_fill(mo.ConvDecoder(masks, lambda x: x, lambda t: t[:, :, :4, 0], uv_size=128, init_uv_size=64,
                                        n_init_channels=16))
    assert dec.pose_cond_mask.dtype == torch.int32 and dec.n_blocks == 1 and dec.n_channels == [16, 8]
    pose, embs, face = mo.seeded_inputs()
    with torch.no_grad():
        out = dec(pose, embs, face)
    non_head = (dec.body_cond_mask * (1 - dec.face_cond_mask)).clip(0, 1)
    assert bool((out["pose_conv"][(non_head == 0).expand_as(out["pose_conv"])] == 0).all())
    assert out["geom_uv_delta_rec"].shape == (1, 3, 128, 128) and out["geom_delta_rec"].shape == (1, 3, 4)
    # global root pose (first 6 dims) does not enter the decoder
    pose2 = pose.clone()
    pose2[:, :6] += 1.0
    with torch.no_grad():
        out2 = dec(pose2, embs, face)
    assert torch.equal(out["tex_mean_rec"], out2["tex_mean_rec"])
