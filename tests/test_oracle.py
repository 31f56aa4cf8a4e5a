"""CPU: the oracle restatement reproduces the reference's outputs stored in tests/golden/*.npz
(generated from the UNMODIFIED reference by tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import femasr_oracle as O
from tests.golden_util import GOLDEN, IDS, gt_indices_of, indices_of, load_case


def digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def sample(t):
    return t[:, ::17, ::3, ::3].contiguous().numpy()


def test_golden_present():
    assert len(GOLDEN) >= 14


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_oracle_matches_reference_golden(path):
    g, sd, cbs = load_case(path)
    assert digest(sd) == str(g["digest"]), "seeded weight generator drifted from the one used for the goldens"
    scale, entry = int(g["scale"]), str(g["entry"])
    x = torch.from_numpy(g["input"])
    with torch.no_grad():
        if entry == "forward":
            taps = {}
            out, loss, sem, idx = O.encode_and_decode(sd, x, scale, taps, cb_scales=[c[0] for c in cbs],
                                                      gt_indices=gt_indices_of(g))
            want_idx = indices_of(g)
            assert len(idx) == len(want_idx) == len(cbs)
            for a, b in zip(idx, want_idx):
                assert np.array_equal(a.numpy(), b), "codebook indices must be bit-exact"
            np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=1e-6)
            assert float(sem) == 0.0
            pairs = (("enc0", "swin"), ("enc1", "up1"), ("enc2", "up2"), ("z", "z"),
                     ("after_quant", "after_quant"), ("dec0", "dec0"), ("dec1", "dec1"), ("dec2", "dec2"))
            if scale == 1:      # HQ stage: enc_feats are the down blocks reversed; only the last one is hooked
                pairs = (("enc0", "down"),) + pairs[3:]
            for ours, theirs in pairs:
                np.testing.assert_allclose(sample(taps[ours]), g["tap_" + theirs], rtol=0, atol=1e-5)
        elif entry == "test":
            out = O.test(sd, x, scale)
        elif entry == "test_tile":
            out = O.test_tile(sd, x, scale, int(g["arg_tile_size"]), int(g["arg_tile_pad"]))
        else:
            out = O.decode_indices(sd, x)
    # same ATen CPU kernels, same op order: the restatement is expected to be bit-identical here;
    # 1e-5 leaves room for thread-count-dependent summation order on other hosts.
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-5)


import glob as _glob
import os as _os

CPU_ONLY = sorted(_glob.glob(_os.path.join(_os.path.dirname(__file__), "golden", "cpu_only", "*.npz")))


@pytest.mark.parametrize("path", CPU_ONLY, ids=[_os.path.basename(p)[:-4] for p in CPU_ONLY])
def test_oracle_constructor_flags_match_reference(path):
    """use_residual=False / use_quantize=False (femasr_arch.py:224,226,349-350,361-362), pinned for the oracle; the CUDA
    path is compared with the oracle for these flags in tests/test_net_gpu.py."""
    g, sd, cbs = load_case(path)
    assert digest(sd) == str(g["digest"])
    flags = {k[5:]: bool(g[k]) for k in g.files if k.startswith("ctor_")}
    assert flags, "cpu_only fixtures carry constructor flags"
    with torch.no_grad():
        out, loss, sem, idx = O.encode_and_decode(sd, torch.from_numpy(g["input"]), int(g["scale"]),
                                                  cb_scales=[c[0] for c in cbs], **flags)
    for a, b in zip(idx, indices_of(g)):
        assert np.array_equal(a.numpy(), b)
    np.testing.assert_allclose(loss.numpy(), g["loss"], rtol=1e-6)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=1e-5)


def test_cpu_only_goldens_present():
    assert len(CPU_ONLY) == 3


def test_flop_model_matches_survey():
    assert abs(O.flops_per_image(4, 128, 128, 256) / 1e9 - 754.53) < 0.01
    assert abs(O.flops_per_image(4, 128, 128, 512) / 1e9 - 762.05) < 0.01
    assert abs(O.flops_per_image(2, 256, 256, 256) / 1e9 - 841.91) < 0.01
    assert abs(O.flops_per_image(4, 144, 144, 256) / 1e9 - 954.96) < 0.5


def test_tile_plan_covers_image_once():
    from femasr_b200.net import tile_plan
    for (h, w, ts, tp) in ((72, 56, 32, 8), (1024, 1024, 256, 32), (100, 37, 240, 16), (33, 65, 32, 0)):
        cover = np.zeros((h, w), dtype=int)
        ref = O.tile_plan(h, w, ts, tp)
        mine = tile_plan(h, w, ts, tp)
        assert len(ref) == len(mine)
        for a, b in zip(ref, mine):
            assert a["in_win"] == b["in"] and a["out_win"] == b["out"] and a["crop"][:2] == b["crop"]
            y0, y1, x0, x1 = b["out"]
            cover[y0:y1, x0:x1] += 1
        assert (cover == 1).all()
