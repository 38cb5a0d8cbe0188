"""CPU-side checks of the hand-MVP host mirror: checkpoint layout against the reference's modules
(tests/golden/hand_mvp_ref.npz <- tests/golden/make_hand_mvp_golden.py) and the fail-loudly rule."""
import json
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hand_mvp_ref.npz")


def test_state_dict_layouts_match_reference():
    """checkpoint compatibility: same keys and shapes as the reference's modules at their real sizes"""
    from goliath_b200.hand_mvp import DeconvContentDecoder, PoseEncoder, TransDecoder

    layouts = json.loads(str(np.load(GOLD)["layouts"]))
    for name, mod in (("PoseEncoder", PoseEncoder(48, 64, 64)), ("TransDecoder", TransDecoder(64)),
                      ("DeconvContentDecoder", DeconvContentDecoder(8, 66, 3))):
        ours = sorted([k, list(v.shape)] for k, v in mod.state_dict().items())
        assert ours == sorted(layouts[name]), name



def test_hand_mvp_layers_refuse_cpu_tensors():
    from goliath_b200 import nn as gnn

    layer = gnn.Conv2dWNUB(3, 4, 8, 8, 3, 1, 1)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 3, 8, 8))
