"""C2 regex with captures over 1M x 1 KB lines: gate + hint (word-skipping) + Pike on the 10 % that match."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

n = 1 << 20
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)])
p2 = cuda.CudaProgram(corpus.C2_REGEX)
rc = torch.empty(n, dtype=torch.int32, device="cuda")
ov = torch.empty((n, p2.nslots), dtype=torch.int64, device="cuda")
p2.pike_lines(dev, n, 1024, 1024, out_rc=rc, out_ovec=ov)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    p2.pike_lines(dev, n, 1024, 1024, out_rc=rc, out_ovec=ov)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"C2 regex, rc + ovector, 1M lines: {ms:.3f} ms {n * 1024 / ms / 1e6:.1f} GB/s, hits {int((rc == 0).sum())}")
