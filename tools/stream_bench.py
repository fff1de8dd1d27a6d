"""Times the chunk-parallel stream scan (sre_cuda_thompson_exec_stream) on one GPU:
C5's text ("abccc" x N + "aaabbccb", bench/gen-data.pl:9) at several sizes with the
bench regex, and 1 GiB of log text with larger automata.
  python tools/stream_bench.py [GiB ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import capi, corpus, cuda  # noqa: E402


def gen_stream(nbytes, device="cuda"):
    """first nbytes-8 bytes of the periodic text, then the 8-byte tail that matches"""
    unit = torch.tensor(list(b"abccc"), dtype=torch.uint8, device=device)
    out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    step = 5 << 24
    base = unit.repeat(step // 5)
    for i in range(0, nbytes, step):
        m = min(step, nbytes - i)
        out[i:i + m] = base[:m]             # step is a multiple of 5: the phase is kept
    out[nbytes - 8:] = torch.tensor(list(b"aaabbccb"), dtype=torch.uint8, device=device)
    return out


def timed(fn, reps=3):
    fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        r = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, r


sizes = [float(x) for x in sys.argv[1:]] or [1, 8]
p1 = cuda.CudaProgram(corpus.BENCH_REGEX)
print("bench regex: dfa states", p1.info.dfa_states, "image states", p1.info.image_states, flush=True)
for g in sizes:
    n = int(g * (1 << 30))
    buf = gen_stream(n)
    ms, (rc, st, mc) = timed(lambda: p1.thompson_stream(buf, n, 65536, True))
    print(f"C5 {g:g} GiB: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  rc {rc} match_chunk {mc} (want {(n - 1) // 65536})",
          flush=True)
    assert rc == capi.SRE_OK and mc == (n - 1) // 65536
    ms, (rc, st, mc) = timed(lambda: p1.thompson_stream(buf, n - 8, 65536, True))
    print(f"C5 {g:g} GiB no-match variant: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s  rc {rc}", flush=True)
    assert rc == capi.SRE_DECLINED
    del buf

n = 1 << 20
dev = torch.empty((n, 1024), dtype=torch.uint8, device="cuda")
for i in range(0, n, 1 << 17):
    dev[i:i + (1 << 17)] = corpus.log_lines(1 << 17, 1024, device="cuda", first_line=i, hit_rate=0.0)
flat = dev.view(-1)
for name, rx in (("bench regex", corpus.BENCH_REGEX), ("C2 regex", corpus.C2_REGEX), ("C3 regex", corpus.C3_REGEX),
                 ("64-pattern set", corpus.multi_pattern_set(64))):
    p = cuda.CudaProgram(rx)
    ms, (rc, st, mc) = timed(lambda: p.thompson_stream(flat, flat.numel(), 65536, True))
    print(f"log text 1 GiB, {name} (dfa {p.info.dfa_states}, image {p.info.image_states}): {ms:.3f} ms  "
          f"{flat.numel() / ms / 1e6:.1f} GB/s  rc {rc} match_chunk {mc}", flush=True)
print("launches", cuda.launch_count())
