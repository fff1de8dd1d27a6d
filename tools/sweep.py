"""Time the Thompson line kernels over (engine, variant, L2 promotion) combos on
the C2 corpus (1M x 1 KB lines), verifying the verdicts each time.
Usage: python tools/sweep.py engine:variant:promo [...]   e.g. skip:0:3 tiled:0:2"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from sregex_b200 import corpus, cuda  # noqa: E402

n, pitch = 1 << 20, 1024
dev = torch.cat([corpus.log_lines(1 << 17, pitch, device="cuda", first_line=i) for i in range(0, n, 1 << 17)])
prog = cuda.CudaProgram(corpus.C2_REGEX)
L = cuda.lib().L
ref = prog.thompson_lines(dev, n, pitch, pitch, engine=cuda.ENGINE_DFA_GENERIC)
rc = torch.empty(n, dtype=torch.int32, device="cuda")
eng = {"tiled": cuda.ENGINE_DFA_TILED, "skip": cuda.ENGINE_DFA_SKIP, "auto": cuda.ENGINE_AUTO,
       "generic": cuda.ENGINE_DFA_GENERIC}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for spec in sys.argv[1:]:
    e, v, p = spec.split(":")
    cuda.set_variant(int(v))
    L.sre_cuda_set_l2_promotion(int(p))
    try:
        for _ in range(5):
            prog.thompson_lines(dev, n, pitch, pitch, engine=eng[e], out=rc)
        a.record()
        for _ in range(50):
            prog.thompson_lines(dev, n, pitch, pitch, engine=eng[e], out=rc)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 50
        ok = bool(torch.equal(rc, ref))
        print(f"{spec:16s} {ms:.4f} ms  {(n * pitch + 4 * n) / ms / 1e6:6.0f} GB/s  frac {(n * pitch + 4 * n) / ms / 1e6 / 6547.2:.3f}  parity {ok}",
              flush=True)
    except Exception as ex:
        print(spec, "failed:", ex, flush=True)
