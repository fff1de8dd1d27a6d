"""One-pass text grep (verdicts only) at several input sizes and piece sizes.
Usage: python tools/text_bench.py [MiB ...]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import capi, corpus, cuda  # noqa: E402

sizes = [int(a) for a in sys.argv[1:]] or [4, 16, 64, 256, 1024]
prog = cuda.CudaProgram(corpus.C2_REGEX)
for mib in sizes:
    n = mib << 20
    flat = torch.cat([corpus.log_lines(min(1 << 17, n >> 10), 1024, device="cuda", first_line=i).view(-1)
                      for i in range(0, n >> 10, 1 << 17)])[:n].clone()
    pos = torch.arange(n, dtype=torch.int64, device="cuda")
    hsh = (pos * -7046029254386353131) ^ (pos >> 13)
    flat[((hsh >> 17) & 0xFFFF) % 180 == 0] = 10
    del pos, hsh
    nl = int((flat == 10).sum()) + 1
    rc = torch.empty(nl, dtype=torch.int32, device="cuda")
    cnt = ctypes.c_size_t(0)

    def step():
        r = prog.lib.L.sre_cuda_thompson_exec_text(prog.cp, flat.data_ptr(), n, None, rc.data_ptr(), nl,
                                                   ctypes.byref(cnt), torch.cuda.current_stream().cuda_stream)
        assert r == capi.SRE_OK

    for piece in (os.environ.get("PIECES", "auto,512,1024,4096").split(",")):
        if piece == "auto":
            os.environ.pop("SRE_CUDA_TEXT_PIECE", None)
        else:
            os.environ["SRE_CUDA_TEXT_PIECE"] = piece
        for _ in range(3):
            step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        a.record()
        for _ in range(reps):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        print(f"{mib:5d} MiB piece {piece:>5s}: {ms:8.3f} ms {n / ms / 1e6:9.1f} GB/s  lines {cnt.value}", flush=True)
