"""Debug build only (-DTF_TIMING): per-K-block clock64 stamps of CTA 0 of the in-kernel-split GEMM."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mclstexp_b200 import ops, _lib
lib = _lib.load()
M, N, K = (int(x) for x in (sys.argv[1:4] or (1024, 256, 256)))
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 1
lead = (nb,) if nb > 1 else ()
a = torch.randn(lead + (M, K), device="cuda"); b = torch.randn(lead + (N, K), device="cuda")
for _ in range(3):
    ops.matmul(a, b)
torch.cuda.synchronize()
buf = (C.c_longlong * (64 * 16))()
lib.mclst_debug_tf32_timing.argtypes = [C.c_void_p, C.c_int]
lib.mclst_debug_tf32_timing(buf, 64 * 16)
nkb = min(64, (K + 31) // 32)
t0 = buf[0]
names = ["top", "loads issued", "empty ok", "stored", "fenced", "arrived", "", "", "mma: wait full", "mma: full ok", "mma: issued"]
print("kb  " + "  ".join(f"{n[:12]:>12s}" for n in names if n))
for kb in range(nkb):
    row = [buf[kb * 16 + s] - t0 for s in (0, 1, 2, 3, 4, 5, 8, 9, 10)]
    print(f"{kb:2d}  " + "  ".join(f"{v:12d}" for v in row))
