"""Stream scan (C5) on the two kinds of text: the log corpus and the reference's
bench/gen-data.pl periodic text ("abccc" ..., worst case for table bank conflicts).
Usage: python tools/stream_c1.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

n = 1 << 20
logs = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)]).view(-1)
per = corpus.gen_data_buffer((1 << 30) // 5, device="cuda")
p1 = cuda.CudaProgram(corpus.BENCH_REGEX)
for name, buf in (("log corpus", logs), ("abccc text", per)):
    r = p1.thompson_stream(buf, buf.numel(), 65536, True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        p1.thompson_stream(buf, buf.numel(), 65536, True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {dt * 1e3:8.3f} ms {buf.numel() / dt / 1e9:8.1f} GB/s  result {r}", flush=True)
