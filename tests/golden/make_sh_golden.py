"""Golden vectors for goliath_b200.sh.dir2sh from the reference's own ca_code/utils/sh.py (imported from
/root/reference in the build container; the fixture travels, the reference does not).
    python tests/golden/make_sh_golden.py"""
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, "/root/reference")
from ca_code.utils import sh as ref_sh  # noqa: E402

g = th.Generator().manual_seed(81)
d = th.randn(257, 3, generator=g, dtype=th.float64)
d = d / d.norm(dim=-1, keepdim=True)
d[0] = th.tensor([0.0, 0.0, 1.0], dtype=th.float64)   # poles and axes
d[1] = th.tensor([0.0, 0.0, -1.0], dtype=th.float64)
d[2] = th.tensor([1.0, 0.0, 0.0], dtype=th.float64)
d[3] = th.tensor([0.0, -1.0, 0.0], dtype=th.float64)
vals = ref_sh.dir2sh_torch(8, d)
vals32 = ref_sh.dir2sh_torch(8, d.float())
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sh_ref.npz"), dirs=d.numpy(),
                    sh_deg8_f64=vals.numpy(), sh_deg8_f32=vals32.numpy())
print(vals.shape)
