import sys; sys.path.insert(0, '.')
import torch, leaf_pytorch_b200 as L
fe = L.Leaf().cuda()
x = (torch.randn(256, 1, 16000, generator=torch.Generator().manual_seed(1)).clamp_(-4, 4) / 4).cuda()
with torch.no_grad():
    for _ in range(3):
        fe(x)
torch.cuda.synchronize()
