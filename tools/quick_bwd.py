"""Development timing helper: forward+backward (param grads) at BASELINE configs[2]-like shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import leaf_pytorch_b200 as L

F = int(os.environ.get("F", 80)); B = int(os.environ.get("B", 1024)); T = int(os.environ.get("T", 16000))
ALGO = os.environ.get("ALGO", "auto")
g = torch.Generator().manual_seed(1234)
x = (torch.randn(B, 1, T, generator=g).clamp_(-4, 4) / 4).cuda()
fe = L.Leaf(n_filters=F, algo=ALGO).cuda()
out = fe(x)
G = torch.randn(out.shape, generator=torch.Generator().manual_seed(1235)).cuda()


def step():
    fe.zero_grad(set_to_none=True)
    o = fe(x)
    o.backward(G)


for _ in range(3):
    step()
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
iters = int(os.environ.get("ITERS", 10))
with torch.no_grad():
    e[0].record()
    for _ in range(iters):
        fe(x)
    e[1].record()
torch.cuda.synchronize()
t_f = e[0].elapsed_time(e[1]) / iters
e[1].record()
for _ in range(iters):
    o = fe(x)
e[2].record()
for _ in range(iters):
    step()
e[3].record()
torch.cuda.synchronize()
t_ft = e[1].elapsed_time(e[2]) / iters
t_fb = e[2].elapsed_time(e[3]) / iters
print(f"F={F} B={B} T={T} algo={ALGO}: inference forward {t_f:.3f} ms, training forward {t_ft:.3f} ms, forward+backward {t_fb:.3f} ms "
      f"({B*T/16000/(t_fb*1e-3):.0f} audio-s/s fwd+bwd), grads finite: "
      f"{all(torch.isfinite(p.grad).all().item() for p in fe.parameters())}")
