"""Debug build only (-DGM_TIMING): globaltimer stamps (ns) of CTA 0 of the packed GEMM, next to the
kernel duration CUDA events see."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mclstexp_b200 import ops, _lib
lib = _lib.load()
M, N, K = (int(x) for x in (sys.argv[1:4] or (1024, 1000, 1000)))
a = torch.randn(M, K, device="cuda"); b = torch.randn(N, K, device="cuda")
for _ in range(5):
    ops.matmul(a, b)
torch.cuda.synchronize()
buf = (C.c_ulonglong * 16)()
lib.mclst_debug_gemm_timing.argtypes = [C.c_void_p, C.c_int]
lib.mclst_debug_gemm_timing(buf, 16)
names = ["entry", "prologue done", "first stage full", "all MMAs issued", "accumulator ready", "epilogue done",
         "final sync", "TMEM freed", "chunk 0: TMEM loaded", "chunk 0: transposed", "chunk 0: stored"]
t0 = buf[0]
for i, n in enumerate(names):
    print(f"{n:20s} {buf[i] - t0:8d} ns")
