"""Does the row pitch matter to the TMA-staged DFA kernel?  Same 1 GiB, same
regex, rows of 1 KB / 2 KB / 4 KB / 8 KB (tiled engine, no word skip)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

n = 1 << 20
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)]).view(-1)
for rx in (corpus.BENCH_REGEX, corpus.C2_REGEX):
    p = cuda.CudaProgram(rx)
    for pitch in (1024, 2048, 4096, 8192):
        rows = dev.numel() // pitch
        for eng, name in ((cuda.ENGINE_DFA_TILED, "tiled"), (cuda.ENGINE_AUTO, "auto")):
            p.thompson_lines(dev, rows, pitch, pitch, engine=eng)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                p.thompson_lines(dev, rows, pitch, pitch, engine=eng)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            print(f"{rx[:14]!r:20} pitch {pitch:5d} {name:5s}: {ms:7.4f} ms {dev.numel() / ms / 1e6:8.1f} GB/s", flush=True)
