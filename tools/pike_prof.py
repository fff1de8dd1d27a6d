import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sregex_b200 import corpus, cuda
n = 1 << 18
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)])
p3 = cuda.CudaProgram(corpus.C3_REGEX)
for _ in range(2):
    rc, ov = p3.pike_lines(dev, n, 1024, 1024)
torch.cuda.synchronize()
print("retry lines:", int((rc == -100).sum()), "matched:", int((rc == 0).sum()))
