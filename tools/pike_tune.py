import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sregex_b200 import corpus, cuda
n = 1 << 18
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)])
p3 = cuda.CudaProgram(corpus.C3_REGEX)
def timed(name, fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print(f"{name:40s} {dt*1e3:9.3f} ms {n*1024/dt/1e9:8.2f} GB/s", flush=True)
print("nctx", os.environ.get("SRE_PIKE_NCTX"))
timed("thompson gate only", lambda: p3.thompson_lines(dev, n, 1024, 1024))
timed("pike (internal gate+hint)", lambda: p3.pike_lines(dev, n, 1024, 1024))
sel = p3.thompson_lines(dev, n, 1024, 1024)
timed("pike gated, no hint (from 0)", lambda: p3.pike_lines(dev, n, 1024, 1024, select=sel), reps=1)
