"""Run the non-headline kernels once each (stream scan, Pike, NFA) so that an
ncu launch list (`ncu --metrics gpu__time_duration.sum -k regex:k_ ...`) shows
their durations.  Usage: python tools/prof_extras.py [lines]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
dev = torch.cat([corpus.log_lines(min(1 << 17, n - i), 1024, device="cuda", first_line=i)
                 for i in range(0, n, 1 << 17)])
flat = dev.view(-1)


def timed(name, fn, nbytes, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:40s} {dt * 1e3:9.3f} ms  {nbytes / dt / 1e9:9.2f} GB/s", flush=True)


p1 = cuda.CudaProgram(corpus.BENCH_REGEX)
p2 = cuda.CudaProgram(corpus.C2_REGEX)
p3 = cuda.CudaProgram(corpus.C3_REGEX)
print("dfa states", p1.info.dfa_states, p2.info.dfa_states, p3.info.dfa_states)
timed("stream scan BENCH_REGEX", lambda: p1.thompson_stream(flat, flat.numel(), 65536, True), flat.numel())
timed("stream scan C2_REGEX", lambda: p2.thompson_stream(flat, flat.numel(), 65536, True), flat.numel())
timed("stream scan C3_REGEX", lambda: p3.thompson_stream(flat, flat.numel(), 65536, True), flat.numel())
timed("thompson lines C2 tiled", lambda: p2.thompson_lines(dev, n, 1024, 1024), flat.numel())
timed("thompson lines C3 tiled", lambda: p3.thompson_lines(dev, n, 1024, 1024), flat.numel())
timed("thompson lines C2 generic", lambda: p2.thompson_lines(dev, n, 1024, 1024, engine=cuda.ENGINE_DFA_GENERIC), flat.numel())
m = min(n, 1 << 16)
timed("thompson lines C2 nfa", lambda: p2.thompson_lines(dev, m, 1024, 1024, engine=cuda.ENGINE_NFA), m * 1024)
timed("pike lines C3", lambda: p3.pike_lines(dev, m, 1024, 1024), m * 1024, reps=1)
sel = p2.thompson_lines(dev, m, 1024, 1024)
timed("pike lines C2 gated (10% hits)", lambda: p2.pike_lines(dev, m, 1024, 1024, select=sel), m * 1024, reps=1)
pm = cuda.CudaProgram(corpus.multi_pattern_set(64))
print("multi64: nfa", pm.info.nfa_states, "dfa", pm.info.dfa_states, "classes", pm.info.dfa_classes)
timed("multi64 thompson gate (auto)", lambda: pm.thompson_lines(dev, n, 1024, 1024), flat.numel())
timed("multi64 thompson nfa", lambda: pm.thompson_lines(dev, m, 1024, 1024, engine=cuda.ENGINE_NFA), m * 1024, reps=1)
timed("multi64 pike (gate+hint inside)", lambda: pm.pike_lines(dev, m, 1024, 1024), m * 1024, reps=1)
rcm, _ = pm.pike_lines(dev, m, 1024, 1024)
print("multi64 matched fraction", float((rcm >= 0).float().mean()))
