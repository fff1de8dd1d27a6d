import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sregex_b200 import corpus, cuda
n = 1 << 21          # 2 GiB stream
dev = torch.cat([corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i) for i in range(0, n, 1 << 17)]).view(-1)
L = cuda.lib().L
for rx in (corpus.BENCH_REGEX, corpus.C3_REGEX):
    p = cuda.CudaProgram(rx)
    ref = None
    for piece in (1024, 2048, 4096, 8192):
        for promo in (2, 3):
            L.sre_cuda_set_stream_piece(piece)
            L.sre_cuda_set_l2_promotion(promo)
            r = p.thompson_stream(dev, dev.numel(), 65536, True)
            ref = ref or r
            assert r == ref, (r, ref)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                p.thompson_stream(dev, dev.numel(), 65536, True)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            print(f"{rx[:12]} dfa {p.info.dfa_states:3d} piece {piece:5d} promo {promo}: {dt*1e3:8.3f} ms {dev.numel()/dt/1e9:8.1f} GB/s  {r}", flush=True)
