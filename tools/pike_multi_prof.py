"""How the 64-pattern Pike pass (C4: matched id + ovector) scales with the
number of lines: a flat curve means one long-running line bounds the launch.
Usage: python tools/pike_multi_prof.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

N = 1 << 18
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, N, 1 << 17)])
pm = cuda.CudaProgram(corpus.multi_pattern_set(64))
print("prog_len", pm.info.prog_len, "ctx bytes", pm.info.pike_ctx_bytes, "slots", pm.info.pike_slots)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for n in (32, 1024, 4096, 16384, 65536, 262144):
    dt = timed(lambda: pm.pike_lines(dev, n, 1024, 1024))
    print(f"lines {n:7d}: {dt * 1e3:8.3f} ms  {n * 1024 / dt / 1e9:8.2f} GB/s", flush=True)

# where does the match sit relative to the line, and which ids win?
rc, ov = pm.pike_lines(dev, 65536, 1024, 1024)
m = rc >= 0
print("matched", int(m.sum()), "ids", torch.bincount(rc[m]).tolist())
span = (ov[m, 1] - ov[m, 0]).float()
print("span mean/max", float(span.mean()), float(span.max()), "start mean", float(ov[m, 0].float().mean()))
# single slow lines: time the lines one block of 32 at a time for the first 2048
worst = []
for i in range(0, 2048, 32):
    dt = timed(lambda: pm.pike_lines(dev[i:i + 32], 32, 1024, 1024), reps=1)
    worst.append((dt, i))
worst.sort(reverse=True)
print("slowest 32-line blocks (ms, first line):", [(round(a * 1e3, 3), b) for a, b in worst[:5]],
      "median", round(sorted(worst)[len(worst) // 2][0] * 1e3, 3))
