"""The program compute-sanitizer runs (scripts/run_sanitizer.sh): one small forward + test() + decode_indices through
the public surface, compared with nothing - the sanitizer's own report is the result."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from basicsr.archs.femasr_arch import FeMaSRNet  # noqa: E402
from femasr_b200.spec import random_state_dict  # noqa: E402

dev = torch.device("cuda", 0)
sd = random_state_dict(4, 256, seed=5, init="perturbed")
net = FeMaSRNet(codebook_params=[[32, 1024, 256]], LQ_stage=True, scale_factor=4)
net.load_state_dict(sd, strict=True)
net = net.to(dev).eval()
x = torch.rand(1, 3, 32, 128, generator=torch.Generator().manual_seed(6)).to(dev)
out, loss, _, idx = net(x)
y = net.test(x[:, :, :24, :40])
z = net.decode_indices(idx[0])
torch.cuda.synchronize()
print(f"target: forward {tuple(out.shape)} test {tuple(y.shape)} decode {tuple(z.shape)} "
      f"launches {net._engine.last_launch_count()} gemm_path {os.environ.get('FEMASR_GEMM_PATH')}")
