"""GPU: the REFERENCE's own entry script (tests/fixtures/inference_femasr.py, a verbatim copy of
/root/reference/inference_femasr.py, see tests/fixtures/README.md) executed unchanged with runpy against this repo's
`basicsr` surface, over synthetic PNGs at the reference testset's size mix (112x112 ... 800x592: images below
600x600 pixels go through FeMaSRNet.test, the two larger ones through test_tile() with its default (240, 16)).
Written PNGs are compared with the CPU oracle's (+-1 LSB) on a subset - the oracle needs about a minute for the big one."""
import os
import runpy
import sys

import cv2
import numpy as np
import pytest
import torch

from basicsr.utils import img2tensor, tensor2img
from femasr_b200.spec import random_state_dict
from oracle import femasr_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
SCRIPT = os.path.join(HERE, "fixtures", "inference_femasr.py")

# (h, w): one of every branch-relevant class of /root/reference/testset (38 images, 24 distinct sizes)
SIZES = [(112, 112), (160, 160), (224, 352), (240, 288), (256, 512), (336, 496), (464, 256), (592, 448), (640, 448),
         (720, 720), (800, 592)]
ORACLE_CHECK = [(112, 112), (224, 352), (800, 592)]       # test(), test() non-square, test_tile() default


def test_reference_entry_script_runs_unchanged_on_the_gpu(tmp_path, cuda):
    sd = random_state_dict(4, 512, seed=7, init="perturbed")         # the script hard-codes codebook 1024x512
    w = tmp_path / "rand.pth"
    torch.save({"params": sd}, w)
    rng = np.random.default_rng(8)
    in_dir, out_dir = tmp_path / "in", tmp_path / "out"
    os.makedirs(in_dir)
    for (h, wd) in SIZES:
        # smooth-ish content (upsampled noise): closer to photographs than white noise, still fully synthetic
        small = rng.integers(0, 256, (h // 8 + 1, wd // 8 + 1, 3), dtype=np.uint8)
        img = cv2.resize(small, (wd, h), interpolation=cv2.INTER_CUBIC)
        cv2.imwrite(str(in_dir / f"im_{h}x{wd}.png"), img)
    argv = sys.argv
    sys.argv = ["inference_femasr.py", "-s", "4", "-i", str(in_dir), "-o", str(out_dir), "-w", str(w)]
    try:
        runpy.run_path(SCRIPT, run_name="__main__")
    finally:
        sys.argv = argv
    torch.set_num_threads(min(16, torch.get_num_threads()))
    for (h, wd) in SIZES:
        got = cv2.imread(str(out_dir / f"im_{h}x{wd}.png"), cv2.IMREAD_UNCHANGED)
        assert got is not None and got.shape == (4 * h, 4 * wd, 3) and got.dtype == np.uint8
    for (h, wd) in ORACLE_CHECK:
        img = cv2.imread(str(in_dir / f"im_{h}x{wd}.png"), cv2.IMREAD_UNCHANGED)
        x = (img2tensor(img) / 255.0).unsqueeze(0)
        with torch.no_grad():
            want = O.test(sd, x, 4) if h * wd < 600 ** 2 else O.test_tile(sd, x, 4)
        want_img = tensor2img(want)
        got = cv2.imread(str(out_dir / f"im_{h}x{wd}.png"), cv2.IMREAD_UNCHANGED)
        diff = np.abs(got.astype(int) - want_img.astype(int))
        print(f"{h}x{wd}: max |diff| {diff.max()} LSB, fraction of differing bytes {(diff > 0).mean():.2e}")
        assert diff.max() <= 1 and (diff > 0).mean() < 0.01, "PNG outputs must match the oracle to +-1 LSB"
