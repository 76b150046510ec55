"""Optimiser step on the two position tables at the cfg2 / her2st shape (65536 x G, B tokens):
lazy row replay (catch-up + step) against torch.optim.Adam (fused) on dense gradients."""
import sys, torch
from torch import nn
sys.path.insert(0, ".")
from mclstexp_b200 import optim as mo
dev = torch.device("cuda", 0)
R, G, B = 65536, int(sys.argv[1]) if len(sys.argv) > 1 else 785, 1024
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tabs = [nn.Parameter(torch.randn(R, G, device=dev) * 0.1) for _ in range(2)]
dense_tabs = [nn.Parameter(t.detach().clone()) for t in tabs]
lazy = mo.LazyEmbeddingAdam(tabs, lr=1e-4, weight_decay=1e-3)
fused = torch.optim.Adam(dense_tabs, lr=1e-4, weight_decay=1e-3, fused=True)
foreach = torch.optim.Adam([nn.Parameter(t.detach().clone()) for t in tabs], lr=1e-4, weight_decay=1e-3)
pos = torch.randint(0, 64, (B, 2), device=dev).float()
d_out = torch.randn(B, G, device=dev)
def lazy_step():
    lazy.catch_up(pos); lazy.record(pos, d_out); lazy.step()
def dense_step(opt):
    def f():
        for p in opt.param_groups[0]["params"]:
            g = torch.zeros_like(p)                      # dense gradient: memset + scatter
            g.index_add_(0, pos[:, 0].long(), d_out)
            p.grad = g
        opt.step()
    return f
t_lazy = timed(lazy_step)
t_fused = timed(dense_step(fused))
t_foreach = timed(dense_step(foreach))
bytes_dense = 2 * R * G * 4 * 7
print(f"G={G} B={B}: lazy catch-up+step {t_lazy:.3f} ms | torch Adam fused + dense grads {t_fused:.3f} ms "
      f"| torch Adam foreach {t_foreach:.3f} ms | dense-pass HBM floor {bytes_dense / 6.55e12 * 1e3:.3f} ms")
t0 = timed(lambda: None, 1)
lazy.flush(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(50): lazy_step()
e0.record(); lazy.flush(); e1.record(); torch.cuda.synchronize()
print(f"flush after 50 deferred steps: {e0.elapsed_time(e1):.3f} ms")
