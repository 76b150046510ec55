"""Per-call time of ops.matmul (3xTF32 in-kernel split) on the shapes of the cfg2 training step, next
to torch.matmul fp32 (cuBLAS SGEMM) on the same tensors.  CUDA events, L2 left warm (as inside a step)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mclstexp_b200 import ops
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda")


def t(fn, iters=20):
    """us per call, replayed from a CUDA graph of `iters` calls (no host launch cost in the figure)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * iters) * 1e3


shapes = [("fwd  x[1024,1000] W[1000,1000]", (1024, 1000), (1000, 1000), False, False),
          ("dx   dy[1024,1000] W[1000,1000]^", (1024, 1000), (1000, 1000), False, True),
          ("dW   dy^[1024,1000] x^[1024,1000]", (1024, 1000), (1024, 1000), True, True),
          ("proj x[1024,1024] W[256,1024]", (1024, 1024), (256, 1024), False, False),
          ("fc   x[1024,256] W[256,256]", (1024, 256), (256, 256), False, False),
          ("qkv  x[1024,1000] W[1536,1000]", (1024, 1000), (1536, 1000), False, False),
          ("QK^T [8,1024,64]x[8,1024,64]", (8, 1024, 64), (8, 1024, 64), False, False),
          ("PV   [8,1024,1024]x[8,1024,64]^", (8, 1024, 1024), (8, 1024, 64), False, True),
          ("dV   P^[8,1024,1024] dO^[8,1024,64]", (8, 1024, 1024), (8, 1024, 64), True, True)]
for name, sa, sb, at, bt in shapes:
    a = torch.randn(*sa, device=dev)
    b = torch.randn(*sb, device=dev)
    ours = t(lambda: ops.matmul(a, b, at, bt))
    opa = a.transpose(-1, -2) if at else a
    opb = b if bt else b.transpose(-1, -2)
    ref = t(lambda: torch.matmul(opa, opb))
    err = (ops.matmul(a, b, at, bt).double() - opa.double() @ opb.double()).abs().max().item()
    print(f"{name:38s} ours {ours:7.1f} us   cuBLAS fp32 {ref:7.1f} us   max err {err:.2e}")
