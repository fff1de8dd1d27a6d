"""Pike + captures for a program WITH look-ahead assertions: the determinised Pike VM (k_pike_lineage,
default tier when SRE_PDFA_LOOKAHEAD allows it) against the closure-table kernel (tier 3), and the
all-matches scan (row f2), on the C2/C3 log lines.  Run on the GPU box; prints one JSON line each.
Results are checked against the CPU oracle on the first lines."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle import cpu_baseline as baseline
from sregex_b200 import corpus, cuda

N = int(os.environ.get("N", 1 << 18))
PITCH = 1024
dev = torch.cat([corpus.log_lines(1 << 17, PITCH, device="cuda", first_line=i) for i in range(0, N, 1 << 17)])
host = dev[:4096].cpu().numpy()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


REGEXES = (rb'\b(GET|HEAD|POST|PUT) (\S+) HTTP/(\d)\.(\d)\b', rb'"\s(\d+)\b.*?(\.*)$', rb'(\w+) (\S+) HTTP/(\d)\.(\d)',
           rb'^(\d+)\.(\d+)\.(\d+)\.(\d+)\b', rb'\b(\d\d\d)\b \.+$', rb'(\w+)\B(\d)" (5\d\d)\b',
           rb'\[(\d+)/(\w+)/(\d+)\b', rb'(?:^|\s)(/x/\d+)\s')
if os.environ.get("ONLY") == "long":        # matches longer than the lineage kernel's ring, and C3
    REGEXES = (REGEXES[1], REGEXES[4], REGEXES[2], rb'(\S+) (\S+) HTTP.*?(\.+)$')
for rx in REGEXES:
    prog = cuda.CudaProgram(rx)
    ns = prog.nslots
    rc = torch.empty(N, dtype=torch.int32, device="cuda")
    ov = torch.empty((N, ns), dtype=torch.int64, device="cuda")
    _, wrc, wov = baseline.run_lines("oracle", rx, None, host, 4096, PITCH, PITCH, baseline.ENGINE_PIKE,
                                     nthreads=8, ovec_slots=ns)
    row = {"regex": rx.decode(), "lines": N, "slots": ns}
    for tier in (0, 3):
        prog.set_pike_tier(tier)
        ms = timed(lambda: prog.pike_lines(dev, N, PITCH, PITCH, out_rc=rc, out_ovec=ov))
        ok = bool((rc[:4096].cpu().numpy() == wrc).all() and (ov[:4096].cpu().numpy() == wov).all())
        row[f"tier{tier}"] = {"kernel_tier": prog.last_pike_tier(), "ms": round(ms, 4),
                              "gbs": round(N * PITCH / ms / 1e6, 1), "matched": int((rc >= 0).sum()),
                              "retry_left": int((rc == -100).sum()), "oracle_ok": ok}
    print(json.dumps(row), flush=True)

if os.environ.get("ONLY"):
    sys.exit(0)
# f2: all non-overlapping matches per line (general kernel)
M = min(N, 1 << 16)
for rx, mm in ((rb'\d+', 16), (rb'(\w+)/', 8)):
    prog = cuda.CudaProgram(rx)
    ms = timed(lambda: prog.pike_lines_all(dev, M, PITCH, PITCH, mm), reps=3)
    cnt, _, _ = prog.pike_lines_all(dev, M, PITCH, PITCH, mm)
    print(json.dumps({"all_matches": rx.decode(), "lines": M, "max_matches": mm, "ms": round(ms, 3),
                      "gbs": round(M * PITCH / ms / 1e6, 1), "matches": int(cnt.sum())}), flush=True)
