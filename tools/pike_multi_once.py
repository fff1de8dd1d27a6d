"""One 64-pattern Pike pass over N lines (for ncu).  Usage: pike_multi_once.py [lines] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.cat([corpus.log_lines(min(1 << 17, N), 1024, device="cuda", first_line=i) for i in range(0, N, 1 << 17)])
pm = cuda.CudaProgram(corpus.multi_pattern_set(64))
for _ in range(reps):
    rc, ov = pm.pike_lines(dev, N, 1024, 1024)
torch.cuda.synchronize()
print("matched", int((rc >= 0).sum()))
