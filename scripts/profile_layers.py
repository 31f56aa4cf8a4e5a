"""Per-layer-shape time table of one batch-32 forward (CUDA events around every launch, FEMASR_PROFILE_DETAIL=1).
Usage (GPU box): python scripts/profile_layers.py [B] > gpurun_out/layers.txt"""
import os
import sys

os.environ["FEMASR_PROFILE_DETAIL"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from femasr_b200.net import NativeNet  # noqa: E402
from femasr_b200.spec import random_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
net = NativeNet(4, 1024, 256, gemm_path=1)
net.load_state_dict(random_state_dict(4, 256, seed=1), dev)
x = torch.rand(B, 3, 128, 128, device=dev)
for _ in range(3):
    net.forward(x)
torch.cuda.synchronize()
net.set_profile(True)
reps = 3
for _ in range(reps):
    net.forward(x)
prof = net.profile()
net.set_profile(False)
rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
tot = sum(v["ms"] for _k, v in rows)
print(f"# batch {B}, {reps} forwards, total {tot / reps:.2f} ms per forward (event-timed launches)")
print(f"{'kernel:shape':52s} {'launches':>8s} {'ms/fwd':>8s} {'share':>6s} {'alg TF/s':>9s}")
for k, v in rows:
    tf = v["flops"] / v["ms"] / 1e9 if v["ms"] > 0 and v["flops"] > 0 else 0.0
    print(f"{k:52s} {v['launches'] // reps:8d} {v['ms'] / reps:8.3f} {v['ms'] / tot * 100:5.1f}% {tf:9.1f}")
