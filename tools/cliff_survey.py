"""Throughput survey over ordinary log regexes: Thompson verdicts (AUTO engine), Pike + captures
(default tiers), text grep, on 131,072 x 1 KB log lines -- to find regex shapes that fall off the
fast paths.  Run on the GPU box; one JSON line per regex."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sregex_b200 import corpus, cuda

N, PITCH = 1 << 17, 1024
dev = corpus.log_lines(N, PITCH, device="cuda")
REGEXES = [rb'error', rb'[Ee]rror|WARN', rb'\d{3}', rb'HTTP/1\.[01]" 404 ', rb'^\d+\.\d+\.\d+\.\d+',
           rb'(\d+)\.(\d+)\.(\d+)\.(\d+)', rb'"(GET|POST|HEAD) ([^ ]+)', rb'\[([^\]]+)\]', rb'\s(\d+)$',
           rb'\w+@\w+\.com', rb'(a|b|c|d|e)+z', rb'.*foo', rb'[0-9a-f]{8}', rb'x{2,5}y', rb'(\S+)\s+(\S+)\s+(\S+)',
           rb'\b\d{1,3}\b', rb'/x/(\d+)', rb'\d+\.?', rb'a+b?', rb'"\s(\d{3})\s', rb'(?:GET|HEAD) /x/\d+ HTTP',
           rb'(\w+)=(\w+)', rb'^.*$', rb'\S+$', rb'[A-Z]{3,} ', rb'(\d\d):(\d\d):(\d\d)']


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for rx in REGEXES:
    try:
        prog = cuda.CudaProgram(rx)
    except Exception as e:                      # noqa: BLE001
        print(json.dumps({"regex": rx.decode(), "error": str(e)[:80]}), flush=True)
        continue
    rc = torch.empty(N, dtype=torch.int32, device="cuda")
    prc = torch.empty(N, dtype=torch.int32, device="cuda")
    pov = torch.empty((N, prog.nslots), dtype=torch.int64, device="cuda")
    row = {"regex": rx.decode(), "dfa_states": prog.info.dfa_states, "nfa_states": prog.info.nfa_states}
    ms = timed(lambda: prog.thompson_lines(dev, N, PITCH, PITCH, out=rc))
    row["thompson_gbs"] = round(N * PITCH / ms / 1e6, 1)
    row["matched"] = int((rc == 0).sum())
    ms = timed(lambda: prog.pike_lines(dev, N, PITCH, PITCH, out_rc=prc, out_ovec=pov))
    row["pike_gbs"] = round(N * PITCH / ms / 1e6, 1)
    row["pike_tier"] = prog.last_pike_tier()
    row["pike_left"] = int((prc == -100).sum())
    print(json.dumps(row), flush=True)
