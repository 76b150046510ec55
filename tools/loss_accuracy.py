"""Accuracy of the loss gradients against a float64 closed form, next to the reference's own fp32
formulation in stock PyTorch on the same GPU (gpurun; prints one line per case)."""
import sys, os
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from checkers import loss_closed_form_f64
from mclstexp_b200 import loss as mloss
from bench import torch_soft_loss


def big_inputs(B, seed):
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    c = 4.0 * torch.randn(64, 256, generator=g, device="cuda")
    def emb():
        x = c[torch.randint(0, 64, (B,), generator=g, device="cuda")] + torch.randn(B, 256, generator=g, device="cuda")
        x = x - x.mean(1, keepdim=True)
        return x / x.std(1, keepdim=True, unbiased=False)
    return emb(), emb()


def rel(got, want):
    got = got.double(); err = (got - want).abs(); m = want.abs().max()
    out = []
    for floor in (1e-3, 1e-2):
        big = want.abs() > floor * m
        out.append(float((err[big] / want.abs()[big]).max()))
    return out + [float(err.max() / m), float((got - want).norm() / want.norm())]


for B in (1024, 4096, 16384):
    S, I = big_inputs(B, 100 + B)
    for targets in ("eye", "soft"):
        l64, dS64, dI64 = loss_closed_form_f64(S, I, 1.0, targets)
        a, b = S.clone().requires_grad_(True), I.clone().requires_grad_(True)
        l = mloss.contrastive_loss(a, b, 1.0, targets); l.backward()
        print(f"B={B} {targets}: ours  loss_rel={abs(l.item()-l64)/abs(l64):.2e} dS [elem>1e-3max, elem>1e-2max, maxnorm, norm] = "
              + " ".join(f"{x:.2e}" for x in rel(a.grad, dS64)) + "  dI " + " ".join(f"{x:.2e}" for x in rel(b.grad, dI64)), flush=True)
        a, b = S.clone().requires_grad_(True), I.clone().requires_grad_(True)
        if targets == "soft":
            lt = torch_soft_loss(a, b, 1.0)
        else:
            lg = a @ b.T
            lab = torch.eye(B, device="cuda")
            lt = (torch.nn.functional.cross_entropy(lg, lab) + torch.nn.functional.cross_entropy(lg.T, lab.T)) / 2
        lt.backward()
        print(f"B={B} {targets}: torch loss_rel={abs(lt.item()-l64)/abs(l64):.2e} dS [elem>1e-3max, elem>1e-2max, maxnorm, norm] = "
              + " ".join(f"{x:.2e}" for x in rel(a.grad, dS64)) + "  dI " + " ".join(f"{x:.2e}" for x in rel(b.grad, dI64)), flush=True)
        del a, b, dS64, dI64
        torch.cuda.empty_cache()
