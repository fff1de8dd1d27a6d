"""'\n'-delimited text end to end on the device: line index, ragged Thompson
verdicts, ragged Pike captures.  Usage: python tools/grep_prof.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

n = 1 << 20
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)])
# variable-length lines: keep the last 120..376 bytes before the padding, newline-terminated
g = torch.Generator(device="cuda").manual_seed(5)
keep = torch.randint(120, 377, (n,), device="cuda", generator=g)
end = 1024 - 60
cols = torch.arange(1024, device="cuda")[None, :]
mask = (cols >= (end - keep)[:, None]) & (cols < end)
rows = dev.clone()
rows[:, end] = 10
mask[:, end] = True
flat = rows[mask].contiguous()
print("bytes", flat.numel(), "lines", n, "avg line", flat.numel() / n)


def timed(name, fn, nbytes, reps=5):
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{name:34s} {ms:8.3f} ms {nbytes / ms / 1e6:9.1f} GB/s", flush=True)


off = cuda.index_lines(flat)
assert off.numel() == n + 1
timed("index_lines", lambda: cuda.index_lines(flat, max_lines=n), flat.numel())
p2 = cuda.CudaProgram(corpus.C2_REGEX)
p3 = cuda.CudaProgram(corpus.C3_REGEX)
timed("thompson_ragged C2", lambda: p2.thompson_ragged(flat, off), flat.numel())
timed("pike ragged C3 (4 groups)", lambda: p3.pike_lines(flat, n, 0, 0, offsets=off), flat.numel())
rc, ov = p3.pike_lines(flat, n, 0, 0, offsets=off)
print("matched", int((rc == 0).sum()), "tier", prog.last_pike_tier())
