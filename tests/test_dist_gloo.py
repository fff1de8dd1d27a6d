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

import oracle  # noqa: F401  (registers the checker libraries with capi)
from sregex_b200 import capi, corpus
from sregex_b200 import dist as sdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lc():
    lc = C.CDLL(os.path.join(ROOT, "oracle", "liblowercheck.so"))
    lc.lc_create.restype = C.c_void_p
    lc.lc_create.argtypes = [C.c_void_p, C.c_uint]
    lc.lc_dfa_fin.argtypes = [C.c_void_p, C.c_uint]
    lc.lc_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
    lc.lc_stream_part_record.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint,
                                         C.c_char_p]
    lc.lc_stream_part_run.restype = C.c_longlong
    lc.lc_stream_part_run.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint, C.POINTER(C.c_uint)]
    return lc


class _Info:
    dfa_start, dfa_acc = 0, 1


class _CpuScan:
    """CPU stand-in for cuda.StreamScan (records from the lowering checker)"""

    def __init__(self, prog, data, fn):
        self.prog, self.data, self.fn = prog, data, fn

    def resolve(self, entry):
        ex = C.c_uint(0)
        off = self.prog.lc.lc_stream_part_run(self.prog.h, self.data, len(self.data), entry, C.byref(ex))
        rec = C.create_string_buffer(32)
        self.prog.lc.lc_stream_part_record(self.prog.h, self.data, len(self.data), None, 0, entry, rec)
        self.fn = rec.raw
        return ex.value, (0 if entry == _Info.dfa_acc else off)

    def close(self):
        pass


class _CpuProg:
    """CPU stand-in for cuda.CudaProgram as far as dist.stream_match_sharded uses it"""
    info = _Info

    def __init__(self, lc, h, hide_halo=False):
        self.lc, self.h, self.hide_halo = lc, h, hide_halo

    def stream_reduce(self, shard, shard_len, halo=None, entry_state=sdist.UNKNOWN):
        data = bytes(shard[:shard_len].tolist())
        rec = C.create_string_buffer(32)
        hb = bytes(halo.tolist()) if halo is not None and not self.hide_halo else None
        entry = 0 if entry_state == 0xFFFFFFFF else entry_state
        assert self.lc.lc_stream_part_record(self.h, data, len(data), hb, len(hb) if hb else 0, entry, rec) == 0
        return _CpuScan(self, data, rec.raw)

    def dfa_fin(self, state):
        return bool(self.lc.lc_dfa_fin(self.h, state))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, data, hide_halo, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sregex_b200 import cuda        # host-only use: the record format's apply()
        o = capi.load("oracle")
        lc = _lc()
        p = o.compile(corpus.BENCH_REGEX)
        h = lc.lc_create(p.prog, 4096)
        first, count = sdist.shard_range(len(data), rank, world)
        shard = torch.tensor(list(data[first:first + count]), dtype=torch.uint8)
        # hide_halo: rank 1's record comes out unresolved, which forces the second exchange round
        prog = _CpuProg(lc, h, hide_halo=hide_halo)
        rc, g = sdist.stream_match_sharded(prog, shard, count, first, eof=True, apply=cuda.fn_apply)
        hits = sdist.allreduce_sum(rank + 1, "cpu")
        q.put((rank, rc, g, hits))
    finally:
        dist.destroy_process_group()


def _run(data, world=2, hide_halo=False):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, data, hide_halo, q)) for r in range(world)]
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
    """halo exchange + record exchange + ordered chaining over gloo; the records
    come from the CPU lowering checker (on the GPU box: sre_cuda_thompson_stream_reduce)"""
    o = capi.load("oracle")
    p = o.compile(corpus.BENCH_REGEX)
    base = bytes(corpus.gen_data_buffer(3000).numpy())          # match ends at the last byte
    cases = [(base, False), (base + b"zz", False), (base[:-8], False), (b"aaabbccb" + base[:-8], False),
             (base + b"zz", True), (b"aaabbccbx" + base[:-8], True)]
    for data, hide_halo in cases:
        want = o.thompson(p, data)
        res = _run(data, hide_halo=hide_halo)
        assert len({(r[1], r[2]) for r in res}) == 1            # every rank composes the same answer
        _, rc, first, hits = res[0]
        assert rc == want and hits == 3
        prc, ov = o.pike(p, data)
        if prc >= 0 and ov[1] < len(data):
            # the step that sees the match is one past the match end (sre_vm_thompson.c:233)
            assert first == ov[1]
        else:
            assert first == -1


def test_chain_entries_with_unresolved_record():
    from sregex_b200 import cuda
    import struct
    def rec(pairs, marker=None):
        c = [0xFFFF] * 8
        e = [0xFFFF] * 8
        for i, (a, b) in enumerate(pairs):
            c[i], e[i] = a, b
        if marker is not None:
            c[0] = marker
        return struct.pack("<16H", *(c + e))
    f0, f1, unres = rec([(0, 5)]), rec([(5, 7), (3, 1)]), rec([], marker=0xFFFE)
    assert sdist.chain_entries([f0, f1], 0, cuda.fn_apply) == ([0, 5], 7)
    assert sdist.chain_entries([f0, unres, f1], 0, cuda.fn_apply) == ([0, 5, sdist.UNKNOWN], sdist.UNKNOWN)
    # ACC (1) is absorbing whatever the record holds
    assert sdist.chain_entries([rec([(0, 1)]), f1, unres], 0, cuda.fn_apply)[1] == 1
