"""world_size-2 gloo tests (CPU) of the multi-rank host logic: line sharding and
the stream exchange step (all-gather + ordered composition of transfer functions).
The per-shard transfer functions come from the CPU lowering checker here; on the
GPU box they come from sre_cuda_thompson_stream_reduce."""
import ctypes as C
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sregex_b200 import capi, corpus
from sregex_b200 import dist as sdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lc():
    lc = C.CDLL(os.path.join(ROOT, "oracle", "liblowercheck.so"))
    lc.lc_create.restype = C.c_void_p
    lc.lc_create.argtypes = [C.c_void_p, C.c_uint]
    lc.lc_dfa_fn.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p]
    lc.lc_dfa_fin.argtypes = [C.c_void_p, C.c_uint]
    lc.lc_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
    return lc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = capi.load("oracle")
        lc = _lc()
        p = o.compile(corpus.BENCH_REGEX)
        h = lc.lc_create(p.prog, 4096)
        info = (C.c_uint * 6)()
        lc.lc_info(h, info)
        first, count = sdist.shard_range(len(data), rank, world)
        shard = data[first:first + count]
        fn = C.create_string_buffer(256)
        assert lc.lc_dfa_fn(h, shard, len(shard), fn) == 0
        fns = sdist.exchange_functions(fn.raw[: info[3]], "cpu")
        entry = sdist.stream_entry_state(fns, rank)
        final = sdist.compose(fns, 0)
        # local first-match offset from the true entry state (CPU stand-in for stream_resolve)
        whole = C.create_string_buffer(256)
        off, s = -1, entry
        if entry != sdist.ACC:
            for i in range(len(shard)):
                lc.lc_dfa_fn(h, shard[i:i + 1], 1, whole)
                s = whole.raw[s]
                if s == sdist.ACC:
                    off = i
                    break
        g = sdist.first_match_global(off, first, "cpu")
        hits = sdist.allreduce_sum(rank + 1, "cpu")
        q.put((rank, entry, final, bool(lc.lc_dfa_fin(h, final)), g, hits))
    finally:
        dist.destroy_process_group()


def _run(data, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    return res


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 1 << 20):
        for w in (1, 2, 3, 8):
            parts = [sdist.shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))


def test_stream_exchange_two_ranks_gloo():
    o = capi.load("oracle")
    p = o.compile(corpus.BENCH_REGEX)
    base = bytes(corpus.gen_data_buffer(3000).numpy())          # match ends at the last byte
    for data in (base, base + b"zz", base[:-8], b"aaabbccb" + base[:-8]):
        want = o.thompson(p, data)
        res = _run(data)
        finals = {(r[2], r[3]) for r in res}
        assert len(finals) == 1                                 # every rank composes the same answer
        final, fin = finals.pop()
        got = capi.SRE_OK if (final == sdist.ACC or fin) else capi.SRE_DECLINED
        assert got == want
        assert len({r[4] for r in res}) == 1 and all(r[5] == 3 for r in res)
        first = res[0][4]
        if final == sdist.ACC:
            # the step that sees the match is one past the match end (sre_vm_thompson.c:233)
            rc, ov = o.pike(p, data)
            assert first == ov[1]
        else:
            assert first == -1
