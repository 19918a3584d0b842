import sys; sys.path.insert(0, '.')
import torch, leaf_pytorch_b200 as L
for F in (40, 80):
    for algo in ("tc", "tc_full", "fp32"):
        fe = L.Leaf(n_filters=F, algo=algo).cuda()
        x = torch.randn(3, 1, 3000, generator=torch.Generator().manual_seed(1)).cuda()
        with torch.no_grad():
            o = fe(x)
        torch.cuda.synchronize()
        print(F, algo, float(o.sum()))
fe = L.Leaf().cuda()
x = torch.randn(3, 1, 3000).cuda()
(fe(x) ** 2).sum().backward()
torch.cuda.synchronize()
print("bwd ok")
