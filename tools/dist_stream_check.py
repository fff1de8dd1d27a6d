"""Multi-GPU check of the sharded single-stream scan (SURVEY 8e / BASELINE config
5): every rank owns a contiguous part of one stream, reduces it to a transfer
function on its GPU, the functions are all-gathered over NCCL and composed in
rank order, each rank resolves its first match from its true entry state.
Run:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_stream_check.py
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from sregex_b200 import capi, corpus, cuda  # noqa: E402
from sregex_b200 import dist as sdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

repeat = int(os.environ.get("SRE_REPEAT", str(40_000_000)))       # x5 bytes: 200 MB default
total = repeat * 5 + 8
first, count = sdist.shard_range(total, rank, world)
first = first // 4096 * 4096 if rank else 0                         # aligned shard starts
nxt = (sdist.shard_range(total, rank + 1, world)[0] // 4096 * 4096) if rank + 1 < world else total
count = nxt - first
# the shard of "abccc" x repeat + "aaabbccb" (bench/gen-data.pl:9), built on the GPU
idx = torch.arange(first, first + count, device="cuda", dtype=torch.int64)
unit = torch.tensor(list(b"abccc"), dtype=torch.uint8, device="cuda")
tail = torch.tensor(list(b"aaabbccb"), dtype=torch.uint8, device="cuda")
shard = torch.where(idx < repeat * 5, unit[idx % 5], tail[(idx - repeat * 5).clamp(0, 7)])
prog = cuda.CudaProgram(corpus.BENCH_REGEX)

for variant, extra in (("match at the very end", b""), ("+ 3 more bytes", b"xyz")):
    if extra and rank == world - 1:
        shard_v = torch.cat([shard, torch.tensor(list(extra), dtype=torch.uint8, device="cuda")])
    else:
        shard_v = shard
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc, off = sdist.stream_match_sharded(prog, shard_v, shard_v.numel(), first, eof=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        want_off = -1 if not extra else total        # the step on 'x' sees the MATCH thread
        ok = rc == capi.SRE_OK and off == want_off
        print(f"[{variant}] world {world}: rc {rc} first-match step offset {off} (want {want_off}) "
              f"{'OK' if ok else 'MISMATCH'}  {total / dt / 1e9:.1f} GB/s incl. exchange", flush=True)
        assert ok
dist.barrier()
dist.destroy_process_group()
