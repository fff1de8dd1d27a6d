"""Memory-side ceiling of the line-tile access pattern (TMA boxes 32 x 128 B)
and of a plain contiguous copy, for the roofline discussion in DESIGN.md."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import cuda  # noqa: E402

L = cuda.lib().L
L.sre_cuda_tma_ceiling.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_int, C.c_void_p]
n, pitch = 1 << 20, 1024
buf = torch.randint(0, 255, (n, pitch), dtype=torch.uint8, device="cuda")
rc = torch.empty(n, dtype=torch.int32, device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for promo in (0, 1, 2, 3):
  L.sre_cuda_set_l2_promotion(promo)
  for v in range(5):
    for _ in range(3):
        L.sre_cuda_tma_ceiling(buf.data_ptr(), n, pitch, pitch, rc.data_ptr(), v, None)
    a.record()
    for _ in range(20):
        L.sre_cuda_tma_ceiling(buf.data_ptr(), n, pitch, pitch, rc.data_ptr(), v, None)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    print(f"tma ceiling promo {promo} variant {v}: {ms:.4f} ms  {n * pitch / ms / 1e6:.0f} GB/s read")
from sregex_b200 import corpus  # noqa: E402
prog = cuda.CudaProgram(corpus.C2_REGEX)
for promo in (0, 1, 2, 3):
    L.sre_cuda_set_l2_promotion(promo)
    for eng, var in ((cuda.ENGINE_DFA_TILED, 0), (cuda.ENGINE_DFA_TILED, 20), (cuda.ENGINE_DFA_TILED, 24), (cuda.ENGINE_DFA_SKIP, 30), (cuda.ENGINE_DFA_SKIP, 31)):
        cuda.set_variant(var)
        for _ in range(3):
            prog.thompson_lines(buf.view(-1), n, pitch, pitch, engine=eng, out=rc)
        a.record()
        for _ in range(20):
            prog.thompson_lines(buf.view(-1), n, pitch, pitch, engine=eng, out=rc)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        print(f"dfa engine {eng} variant {var} promo {promo}: {ms:.4f} ms  {n * pitch / ms / 1e6:.0f} GB/s")
dst = torch.empty_like(buf)
for _ in range(3):
    dst.copy_(buf)
a.record()
for _ in range(20):
    dst.copy_(buf)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"torch copy 1 GiB: {ms:.4f} ms  {2 * n * pitch / ms / 1e6:.0f} GB/s (read+write)")
s = buf.view(torch.int64)
for _ in range(3):
    s.sum()
a.record()
for _ in range(20):
    s.sum()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print(f"torch int64 sum 1 GiB: {ms:.4f} ms  {n * pitch / ms / 1e6:.0f} GB/s read")
