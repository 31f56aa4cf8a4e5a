import sys, torch
sys.path.insert(0, '/root/repo')
from basicsr.archs.femasr_arch import FeMaSRNet
from femasr_b200.spec import random_state_dict
dev = torch.device('cuda', 0)
sd = random_state_dict(4, 512, seed=2, init='perturbed')
net = FeMaSRNet(codebook_params=[[32, 1024, 512]], LQ_stage=True, scale_factor=4)
net.load_state_dict(sd); net = net.to(dev).eval()
for shape in ((1, 3, 32, 48), (1, 3, 48, 32)):
    y = net.test(torch.rand(shape, device=dev)); torch.cuda.synchronize(); print('test ok', shape, y.shape, flush=True)
y = net.test_tile(torch.rand(1, 3, 80, 64, device=dev)); torch.cuda.synchronize(); print('tile ok', y.shape, flush=True)
y = net(torch.rand(2, 3, 32, 32, device=dev))[0]; torch.cuda.synchronize(); print('graph fwd ok', y.shape, flush=True)
y = net(torch.rand(2, 3, 32, 32, device=dev))[0]; torch.cuda.synchronize(); print('graph replay ok', y.shape, flush=True)
