"""Run one dbg GEMM variant a few times (for ncu)."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ance_b200 import _lib  # noqa: E402

variant = int(sys.argv[1])
M, N, K = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (65536, 2304, 768)
lib = _lib.load()
A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
B = (torch.randn(N, K, device="cuda") * 0.5).bfloat16()
Cc = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    _lib.check(lib.ance_dbg_gemm(A.data_ptr(), B.data_ptr(), M, N, K, 1, variant, None, None, 0, Cc.data_ptr(), None, st))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    lib.ance_dbg_gemm(A.data_ptr(), B.data_ptr(), M, N, K, 1, variant, None, None, 0, Cc.data_ptr(), None, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"variant {variant} {M}x{N}x{K}: {ms:.4f} ms  {2.0*M*N*K/ms/1e9:.1f} TFLOP/s")
