import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sregex_b200 import corpus, cuda
n = 1 << 20
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)]).view(-1)
p1 = cuda.CudaProgram(corpus.BENCH_REGEX)
for _ in range(2):
    print(p1.thompson_stream(dev, dev.numel(), 65536, True))
