import sys, torch
sys.path.insert(0, ".")
from mclstexp_b200 import _lib, loss as mloss
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
for B in [int(a) for a in sys.argv[1:]] or (8192, 32768):
    x = torch.randn(B, 256, generator=g, device=dev); S = ((x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True)).requires_grad_(True)
    y = torch.randn(B, 256, generator=g, device=dev); I = ((y - y.mean(1, keepdim=True)) / y.std(1, keepdim=True)).requires_grad_(True)
    for _ in range(2):
        mloss.contrastive_loss(S, I, 1.0, "soft").backward()
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    mloss.contrastive_loss(S, I, 1.0, "soft").backward()
    prof = _lib.profile_collect(); _lib.profile_enable(False)
    agg = {}
    for n, t in prof: agg[n] = agg.get(n, 0) + t
    print(B, {k: round(v, 3) for k, v in agg.items()}, "total", round(sum(agg.values()), 3))
