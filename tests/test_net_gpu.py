"""GPU: the whole hot path through the reference-facing surface (basicsr.archs.femasr_arch.FeMaSRNet ->
femasr_b200.net -> C ABI) against (a) the committed golden vectors produced by the unmodified reference
and (b) the CPU oracle on fresh seeded inputs, stage by stage.

Bars (BASELINE.json north_star): output max-abs <= 1e-3 fp32; codebook indices bit-exact.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from femasr_b200 import lib as L

from basicsr.archs.femasr_arch import FeMaSRNet
from femasr_b200.spec import random_state_dict
from oracle import femasr_oracle as O
from tests.golden_util import GOLDEN, IDS, gt_indices_of, indices_of, load_case

pytestmark = pytest.mark.gpu
OUT_ATOL = 1e-3


def make_net(scale, e_dim, sd, cuda, codebooks=None, **kw):
    # scale 1 = the HQ autoencoder (LQ_stage=False).  gemm_path 0: the fp32-FFMA path (ATen-level rounding, tight tolerances below); the default tensor-core
    # path is exercised against the same goldens in tests/test_tc_gpu.py.
    kw.setdefault("gemm_path", 0)
    net = FeMaSRNet(codebook_params=codebooks or [[32, 1024, e_dim]], LQ_stage=scale != 1, scale_factor=scale, **kw)
    net.load_state_dict(sd, strict=True)
    return net.to(cuda).eval()


def check_indices(idx, g):
    want = indices_of(g)
    assert len(idx) == len(want)
    for k, (a, b) in enumerate(zip(idx, want)):
        assert a.dtype == torch.int64 and tuple(a.shape) == b.shape
        mism = int((a.cpu().numpy() != b).sum())
        assert mism == 0, f"codebook {k}: {mism}/{b.size} index mismatches"


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_golden_through_public_surface(cuda, path):
    g, sd, cbs = load_case(path)
    scale, e_dim, entry = int(g["scale"]), int(g["e_dim"]), str(g["entry"])
    net = make_net(scale, e_dim, sd, cuda, codebooks=cbs)
    x = torch.from_numpy(g["input"]).to(cuda)
    with torch.no_grad():
        if entry == "forward":
            gt = gt_indices_of(g)
            out, loss, sem, idx = net(x, gt) if gt is not None else net(x)
            check_indices(idx, g)
            np.testing.assert_allclose(loss.item(), float(g["loss"]), rtol=2e-5)
            assert sem.item() == 0.0
        elif entry == "test":
            out = net.test(x)
        elif entry == "test_tile":
            out = net.test_tile(x, int(g["arg_tile_size"]), int(g["arg_tile_pad"]))
        else:
            out = net.decode_indices(x)
    assert tuple(out.shape) == g["out"].shape
    err = np.abs(out.cpu().numpy() - g["out"]).max()
    assert err <= OUT_ATOL, f"output max-abs {err:.3e}"


@pytest.mark.parametrize("scale,e_dim,shape", [(4, 256, (2, 3, 48, 32)), (2, 512, (1, 3, 64, 96))])
def test_stage_taps_against_oracle(cuda, scale, e_dim, shape):
    sd = random_state_dict(scale, e_dim, seed=31, init="perturbed")
    net = make_net(scale, e_dim, sd, cuda)
    x = torch.rand(shape, generator=torch.Generator().manual_seed(32))
    taps = {}
    with torch.no_grad():
        want, wloss, _, widx = O.encode_and_decode(sd, x, scale, taps)
    eng = net._native(cuda)
    names = ["swin", "up1", "up2", "z", "zq", "after_quant", "dec0", "dec1", "dec2"]
    out, loss, idx, got = eng.forward(x.to(cuda), taps=names)
    expect = {"swin": taps["enc0"], "up1": taps["enc1"], "up2": taps["enc2"], "z": taps["z"], "zq": taps["zq"],
              "after_quant": taps["after_quant"], "dec0": taps["dec0"] + taps["enc1"],
              "dec1": taps["dec1"] + taps["enc2"], "dec2": taps["dec2"]}
    report = {}
    for n in names:
        g = got[n].permute(0, 3, 1, 2).cpu()
        report[n] = (g - expect[n]).abs().max().item() / max(1e-6, expect[n].abs().max().item())
    print("relative stage errors:", {k: f"{v:.2e}" for k, v in report.items()})
    assert torch.equal(idx.cpu(), widx[0]), "indices not bit-exact"
    for n, v in report.items():
        assert v <= 2e-4, f"stage {n}: relative max error {v:.3e}"
    assert (out.cpu() - want).abs().max().item() <= OUT_ATOL
    assert abs(loss.item() - wloss.item()) <= 2e-5 * abs(wloss.item())


def test_weight_update_is_picked_up(cuda):
    sd = random_state_dict(4, 256, seed=33, init="perturbed")
    net = make_net(4, 256, sd, cuda)
    x = torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(34)).to(cuda)
    y0 = net.test(x)
    sd2 = random_state_dict(4, 256, seed=35, init="perturbed")
    net.load_state_dict(sd2, strict=True)
    y1 = net.test(x)
    with torch.no_grad():
        want = O.test(sd2, x.cpu(), 4)
    assert (y1.cpu() - want).abs().max().item() <= OUT_ATOL
    assert (y1 - y0).abs().max().item() > 1e-2


def test_use_residual_and_use_quantize_flags(cuda):
    sd = random_state_dict(4, 256, seed=36, init="perturbed")
    x = torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(37))
    net = make_net(4, 256, sd, cuda, use_quantize=False)
    out = net(x.to(cuda))[0]
    # use_quantize=False: z_quant = feat_to_quant (femasr_arch.py:349-350); restate with the oracle pieces
    with torch.no_grad():
        feats = O.multiscale_encoder(sd, x, 4)[-3:]
        z = O._conv(sd, "before_quant_group.0", feats[0], 1, 0)
        t = O._conv(sd, "after_quant_group.0.conv", z)
        for i in range(3):
            if i > 0:
                t = t + feats[i]
            t = O.decoder_block(sd, f"decoder_group.{i}", t)
        want = O._conv(sd, "out_conv", t)
    assert (out.cpu() - want).abs().max().item() <= OUT_ATOL


def test_geometry_errors(cuda):
    from femasr_b200.lib import FemasrError
    sd = random_state_dict(4, 256, seed=38)
    net = make_net(4, 256, sd, cuda)
    with pytest.raises(FemasrError):
        net(torch.rand(1, 3, 40, 40, device=cuda))      # Swin stage 20x20: reference raises too
    with pytest.raises(AssertionError):
        net.decode_indices(torch.zeros(4, 4, dtype=torch.int64, device=cuda))


def test_cuda_graph_replay_matches_eager(cuda):
    sd = random_state_dict(4, 256, seed=39, init="perturbed")
    net = make_net(4, 256, sd, cuda, gemm_path=1)
    eng = net._native(cuda)
    g = torch.Generator().manual_seed(40)
    for shape in ((2, 3, 32, 48), (1, 3, 48, 32), (2, 3, 32, 48), (2, 3, 32, 48), (1, 3, 48, 32)):
        x = torch.rand(shape, generator=g).to(cuda)
        y0, l0, i0 = eng.forward(x)
        y1, l1, i1 = (t.clone() for t in eng.forward_graph(x))
        assert torch.equal(y0, y1) and torch.equal(i0, i1) and torch.equal(l0, l1)
    # first sighting runs eagerly, the second captures: both shapes came back, so both are graphs now
    assert set(eng._graphs) == {(2, 3, 32, 48), (1, 3, 48, 32)}
    # the public surface uses the graph path and returns tensors that survive the next call
    xa, xb = torch.rand(1, 3, 32, 32, generator=g).to(cuda), torch.rand(1, 3, 32, 32, generator=g).to(cuda)
    ya = net(xa)[0]
    keep = ya.clone()
    net(xb)
    net(xa)
    assert torch.equal(ya, keep)


def test_many_shapes_keep_memory_bounded(cuda):
    """VERDICT r1 / ADVICE: a folder with many image sizes (the reference testset has 38) must not accumulate one
    graph + workspace per shape.  12 distinct shapes, each twice, through sr_uint8: at most graph_cache_size graphs
    stay alive and the peak allocation stays near (cache size + 1) x the largest per-shape footprint."""
    sd = random_state_dict(4, 256, seed=41, init="perturbed")
    net = make_net(4, 256, sd, cuda, gemm_path=1)
    eng = net._native(cuda)
    rng = np.random.default_rng(42)
    shapes = [(40 + 8 * i, 56 + 8 * (i % 5)) for i in range(12)]
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    need = []
    for _ in range(2):
        for (h, w) in shapes:
            img = torch.from_numpy(rng.integers(0, 256, (1, h, w, 3), dtype=np.uint8)).to(cuda)
            out = net.sr_uint8(img)
            assert tuple(out.shape) == (1, 4 * h, 4 * w, 3)
            hp, wp = (h // 16 + 1) * 16, (w // 16 + 1) * 16
            nb = C.c_size_t()
            L.check(eng.lib.femasr_net_workspace_bytes(eng._h, 1, hp, wp, C.byref(nb)))
            need.append(nb.value + 16 * hp * wp * 3 * 4 * 2)
    torch.cuda.synchronize()
    assert len(eng._graphs) <= eng.graph_cache_size
    peak = torch.cuda.max_memory_allocated() - base
    bound = (eng.graph_cache_size + 2) * max(need)
    print(f"12 shapes x 2: {len(eng._graphs)} graphs alive, peak {peak / 2**20:.0f} MiB, bound {bound / 2**20:.0f} MiB, "
          f"sum over shapes {sum(need[:12]) / 2**20:.0f} MiB")
    assert peak <= bound
    eng.release_graphs()


def test_out_of_range_indices_raise(cuda):
    """The reference's scatter_ raises on an index outside the codebook; here the host checks the range (kernels clamp)."""
    from femasr_b200.lib import FemasrError
    sd = random_state_dict(4, 256, seed=43)
    net = make_net(4, 256, sd, cuda)
    bad = torch.full((1, 1, 4, 4), 1024, dtype=torch.int64, device=cuda)
    with pytest.raises(FemasrError, match="out of range"):
        net.decode_indices(bad)
    with pytest.raises(FemasrError, match="out of range"):
        net(torch.rand(1, 3, 32, 32, device=cuda), [torch.full((1, 1, 16, 16), -1, dtype=torch.int64, device=cuda)])


def test_two_devices_in_one_process(cuda):
    """The reference surface is `.to(any device)`: two engines on two GPUs of one process must both work (kernel
    attributes / SM counts are per device).  Needs a 2-GPU box; skipped otherwise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sd = random_state_dict(4, 256, seed=44, init="perturbed")
    x = torch.rand(1, 3, 32, 128, generator=torch.Generator().manual_seed(45))
    outs = []
    for dev in (torch.device("cuda", 1), torch.device("cuda", 0), torch.device("cuda", 1)):
        net = make_net(4, 256, sd, dev, gemm_path=1)
        out, _, _, idx = net(x.to(dev))
        outs.append((out.cpu(), idx[0].cpu()))
        assert out.device == dev
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[2][0])


def test_compact_codes_on_the_device(cuda):
    """encode_codes / decode_codes: the device pack / unpack kernels produce the same stream as the CPU tensor arithmetic
    of femasr_b200/wire.py, and decoding the packed codes equals decode_indices on the index map."""
    from femasr_b200.wire import pack_codes, packed_nbytes, unpack_codes
    sd = random_state_dict(4, 256, seed=46, init="perturbed")
    net = make_net(4, 256, sd, cuda, gemm_path=1)
    x = torch.rand(2, 3, 32, 48, generator=torch.Generator().manual_seed(47)).to(cuda)
    packed, shape = net.encode_codes(x)
    idx = net(x)[3][0]
    assert packed.is_cuda and packed.dtype == torch.uint8 and packed.numel() == packed_nbytes(idx.numel(), 1024) and shape == tuple(idx.shape)
    assert torch.equal(packed.cpu(), pack_codes(idx.cpu(), 1024))
    assert torch.equal(unpack_codes(packed, shape, 1024), idx)
    assert torch.equal(net.decode_codes(packed, shape), net.decode_indices(idx))
    for n_e, n in ((1000, 77), (2, 9), (65536, 64), (512, 4096)):       # odd widths, ragged tails
        codes = torch.randint(0, n_e, (n,), generator=torch.Generator().manual_seed(n)).to(cuda)
        p = pack_codes(codes, n_e)
        assert torch.equal(p.cpu(), pack_codes(codes.cpu(), n_e)) and torch.equal(unpack_codes(p, (n,), n_e), codes)
    with pytest.raises(ValueError):
        pack_codes(torch.tensor([0, 1024], device=cuda), 1024)
