"""GPU: parity at the BASELINE.json geometries (config 2: x4 128x128 LR batch 32; config 3: x2 256x256 LR; config 5:
test_tile(256, 32), reduced to 768x768 so the CPU oracle finishes in about a minute; plus the reference entry
script's default test_tile(240, 16) on a >600-pixel image), default init = the tie-heavy U(+-1/1024) codebook, where
the VQ stage has up to 131072 rows and 0.1-0.2 % of them are exact fp32 ties.

Bars (north_star): codebook indices BIT-EXACT against the CPU oracle (which is pinned to the unmodified reference),
output max-abs <= 1e-3.  Zero tolerance on both GEMM paths.  Every measured (flips, max-abs) pair is appended to
gpurun_out/parity_report.json (copied to profiles/parity_r2.json)."""
import json
import os

import pytest
import torch

from basicsr.archs.femasr_arch import FeMaSRNet
from femasr_b200.spec import random_state_dict
from oracle import femasr_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.json")


def record(name, **fields):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    data = {}
    if os.path.exists(REPORT):
        try:
            data = json.load(open(REPORT))
        except Exception:
            data = {}
    data[name] = fields
    with open(REPORT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def oracle_threads():
    torch.set_num_threads(min(16, torch.get_num_threads()))


def make(scale, gemm_path, cuda, sd):
    kw = {} if gemm_path is None else {"gemm_path": gemm_path}
    net = FeMaSRNet(codebook_params=[[32, 1024, 256]], LQ_stage=True, scale_factor=scale, **kw)
    net.load_state_dict(sd, strict=True)
    return net.to(cuda).eval()


@pytest.mark.parametrize("gemm_path", [0, 1])
def test_benchmark_geometry_parity(cuda, gemm_path):
    B = 4
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    oracle_threads()
    with torch.no_grad():
        want, wloss, _, widx = O.encode_and_decode(sd, x, 4)
    net = make(4, gemm_path, cuda, sd)
    out, loss, _, idx = net(x.to(cuda))
    flips = int((idx[0].cpu() != widx[0]).sum())
    diff = (out.cpu() - want).abs()
    rel_loss = abs(loss.item() - wloss.item()) / wloss.item()
    print(f"gemm_path={gemm_path}: {flips}/{widx[0].numel()} index flips, output max-abs {diff.max():.2e}, "
          f"median {diff.median():.2e}, loss rel diff {rel_loss:.2e}")
    record(f"config2_b4_gemm_path{gemm_path}", flips=flips, rows=widx[0].numel(), max_abs=diff.max().item(),
           loss_rel=rel_loss, entry="forward", lr=[128, 128], scale=4, batch=B)
    assert flips == 0, f"{flips} index flips at the benchmark geometry"
    assert diff.max().item() <= 1e-3
    assert rel_loss <= 2e-5


def test_config2_full_batch_against_oracle(cuda):
    """BASELINE config 2 at its FULL size on the default (tcgen05) path: batch 32, 131072 VQ rows, every index and every
    output pixel against the CPU oracle (about 45 s of host time)."""
    B = 32
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    oracle_threads()
    net = make(4, None, cuda, sd)
    out, loss, _, idx = net(x.to(cuda))
    out, idx0 = out.cpu(), idx[0].cpu()
    flips, worst = 0, 0.0
    losses = []
    with torch.no_grad():
        for i in range(0, B, 8):                       # the oracle is per-sample: chunks keep its memory small
            want, wloss, _, widx = O.encode_and_decode(sd, x[i:i + 8], 4)
            flips += int((idx0[i:i + 8] != widx[0]).sum())
            worst = max(worst, (out[i:i + 8] - want).abs().max().item())
            losses.append(wloss.item())
    rel_loss = abs(loss.item() - sum(losses) / len(losses)) / (sum(losses) / len(losses))
    print(f"config 2 full batch: {flips}/{idx0.numel()} flips, max-abs {worst:.2e}, loss rel {rel_loss:.2e}")
    record("config2_b32_default_path", flips=flips, rows=idx0.numel(), max_abs=worst, loss_rel=rel_loss,
           entry="forward", lr=[128, 128], scale=4, batch=B)
    assert flips == 0 and worst <= 1e-3 and rel_loss <= 2e-5


def test_config3_x2_against_oracle(cuda):
    """BASELINE config 3 (x2, 256x256 LR, batch 16): the oracle checks 2 images in full; batch-independence (bit-exact
    sub-batch reproduction) extends that to the other 14."""
    B = 16
    sd = random_state_dict(2, 256, seed=0, init="default")
    x = torch.rand(B, 3, 256, 256, generator=torch.Generator().manual_seed(3))
    oracle_threads()
    net = make(2, None, cuda, sd)
    out, loss, _, idx = net(x.to(cuda))
    with torch.no_grad():
        want, _, _, widx = O.encode_and_decode(sd, x[:2], 2)
    flips = int((idx[0][:2].cpu() != widx[0]).sum())
    err = (out[:2].cpu() - want).abs().max().item()
    print(f"config 3: {flips}/{widx[0].numel()} flips, max-abs {err:.2e}")
    record("config3_x2_256_b16", flips=flips, rows=widx[0].numel(), max_abs=err, entry="forward", lr=[256, 256],
           scale=2, batch=B, oracle_images=2)
    assert flips == 0 and err <= 1e-3
    eng = net._native(cuda)
    y_all, _, i_all = eng.forward(x.to(cuda))
    y_sub, _, i_sub = eng.forward(x[6:8].to(cuda).contiguous())
    assert torch.equal(y_sub, y_all[6:8]) and torch.equal(i_sub, i_all[6:8]), "a sub-batch must reproduce its rows bit-exactly"


def test_config5_tiled_reduced_against_oracle(cuda):
    """BASELINE config 5 (x4 test_tile(256, 32)) on a 768x768 LR image: 9 tiles in all three shape classes of the
    1024x1024 case (corner 288x288, edge 288x320 / 320x288, interior 320x320 -> padded 304..336, 9x9..10x10 windows).
    The oracle runs the reference's one-tile-at-a-time loop."""
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(1, 3, 768, 768, generator=torch.Generator().manual_seed(4))
    oracle_threads()
    net = make(4, None, cuda, sd)
    got = net.test_tile(x.to(cuda), 256, 32).cpu()
    with torch.no_grad():
        want = O.test_tile(sd, x, 4, 256, 32)
    diff = (got - want).abs()
    frac = float((diff > 1e-3).float().mean())
    print(f"config 5 (768x768): max-abs {diff.max():.2e}, fraction > 1e-3: {frac:.2e}")
    record("config5_tile256_pad32_768", max_abs=diff.max().item(), frac_gt_1e3=frac, entry="test_tile", lr=[768, 768],
           scale=4, tiles=9)
    assert tuple(got.shape) == (1, 3, 3072, 3072)
    assert diff.max().item() <= 1e-3


def test_default_tile_path_of_the_entry_script(cuda):
    """inference_femasr.py:58-61 sends images of >= 600x600 pixels through test_tile() with its DEFAULT (240, 16):
    a 616x488 image gives 3x3 tiles of five different shapes."""
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(1, 3, 488, 616, generator=torch.Generator().manual_seed(5))
    oracle_threads()
    net = make(4, None, cuda, sd)
    got = net.test_tile(x.to(cuda)).cpu()
    with torch.no_grad():
        want = O.test_tile(sd, x, 4)
    diff = (got - want).abs()
    print(f"default test_tile (488x616): max-abs {diff.max():.2e}")
    record("default_tile240_pad16_488x616", max_abs=diff.max().item(), entry="test_tile", lr=[488, 616], scale=4,
           tiles=len(O.tile_plan(488, 616, 240, 16)))
    assert diff.max().item() <= 1e-3


def test_full_batch_size_independent_properties(cuda):
    """BASELINE config 2 at its full size (batch 32, 128x128, 131072 VQ rows): properties that must hold regardless of
    size.  (a) determinism: two runs are bit-identical (fixed-order reductions everywhere); (b) images are independent
    (GroupNorm per sample, LN per token, attention per window, VQ per pixel): permuting the batch permutes the outputs
    bit-exactly, and a sub-batch reproduces its rows; (c) indices are valid codes and the reported codebook loss equals
    1.25 * mean((e_idx - z)^2) recomputed from the returned stage tensors, and the chosen code is a true nearest
    neighbour up to the fp32 grid.  (The oracle comparison of all 32 images is test_config2_full_batch_against_oracle.)"""
    B = 32
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(1)).to(cuda)
    net = make(4, None, cuda, sd)
    eng = net._native(cuda)
    y0, l0, i0 = eng.forward(x)
    y1, l1, i1 = eng.forward(x)
    assert torch.equal(y0, y1) and torch.equal(i0, i1) and torch.equal(l0, l1), "forward must be deterministic"
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(2)).to(cuda)
    yp, _, ip = eng.forward(x[perm].contiguous())
    assert torch.equal(yp, y0[perm]) and torch.equal(ip, i0[perm]), "batch permutation must commute bit-exactly"
    ys, _, isub = eng.forward(x[5:9].contiguous())
    assert torch.equal(ys, y0[5:9]) and torch.equal(isub, i0[5:9]), "a sub-batch must reproduce its rows bit-exactly"
    assert int(i0.min()) >= 0 and int(i0.max()) < 1024 and i0.dtype == torch.int64 and tuple(i0.shape) == (B, 1, 64, 64)
    _, l2, i2, taps = eng.forward(x, taps=["z", "zq"])
    z = taps["z"].reshape(-1, 256)
    e = sd["quantize_group.0.embedding.weight"].to(cuda)[i2.reshape(-1)]
    want_loss = 1.25 * torch.mean((e - z).double() ** 2)
    assert abs(l2.item() - want_loss.item()) <= 1e-6 * want_loss.item()
    assert torch.equal(taps["zq"].reshape(-1, 256), z + (e - z)), "straight-through output must be z + (e_idx - z)"
    # distances: the chosen code is a true nearest neighbour in exact arithmetic up to the fp32 grid
    d_sel = ((e - z).double() ** 2).sum(1)
    zd = z[:4096].double()
    d_all = (zd * zd).sum(1, keepdim=True) + (sd["quantize_group.0.embedding.weight"].double().to(cuda) ** 2).sum(1) \
        - 2 * zd @ sd["quantize_group.0.embedding.weight"].double().to(cuda).t()
    gap = d_sel[:4096] - d_all.min(1).values
    assert float(gap.max()) <= 2.5e-4, "selected code must be nearest up to fp32 rounding of a ~300-magnitude distance"
