"""GPU: parity at the benchmark geometry (x4, 128x128 LR, default init = the tie-heavy U(+-1/1024) codebook),
where the VQ stage has tens of thousands of rows and 0.1-0.2% of them are exact fp32 ties.

What can be promised: given the SAME z the indices are bit-exact (tests/test_ops_gpu.py).  End to end, z carries
the encoder's accumulated rounding (different summation order than ATen-CPU; on the tensor-core path also the
accumulator truncation), and a row whose two best codes are closer than that noise may resolve differently - the
reference itself does so between fp32 and fp64 (0.1-0.2% of rows).  This test measures it and bounds it."""
import pytest
import torch

from basicsr.archs.femasr_arch import FeMaSRNet
from femasr_b200.spec import random_state_dict
from oracle import femasr_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gemm_path,max_flips", [(0, 2), (1, 12)])
def test_benchmark_geometry_parity(cuda, gemm_path, max_flips):
    B = 4
    sd = random_state_dict(4, 256, seed=0, init="default")
    x = torch.rand(B, 3, 128, 128, generator=torch.Generator().manual_seed(1))
    torch.set_num_threads(min(16, torch.get_num_threads()))
    with torch.no_grad():
        want, wloss, _, widx = O.encode_and_decode(sd, x, 4)
    net = FeMaSRNet(codebook_params=[[32, 1024, 256]], LQ_stage=True, scale_factor=4, gemm_path=gemm_path)
    net.load_state_dict(sd, strict=True)
    net = net.to(cuda).eval()
    out, loss, _, idx = net(x.to(cuda))
    flips = int((idx[0].cpu() != widx[0]).sum())
    diff = (out.cpu() - want).abs()
    frac_bad = float((diff > 1e-3).float().mean())
    print(f"gemm_path={gemm_path}: {flips}/{widx[0].numel()} index flips, output max-abs {diff.max():.2e}, "
          f"median {diff.median():.2e}, fraction of pixels > 1e-3: {frac_bad:.2e}, loss rel diff "
          f"{abs(loss.item() - wloss.item()) / wloss.item():.2e}")
    assert flips <= max_flips
    if flips == 0:
        assert diff.max().item() <= 1e-3
    else:
        assert frac_bad <= 0.01      # a flipped code only disturbs its own neighbourhood
