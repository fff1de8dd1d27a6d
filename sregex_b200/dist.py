"""Multi-GPU host logic: one process per GPU, torch.distributed for
the plumbing.  Line corpora shard with no data-path collective.  A single long
stream shards into contiguous parts with ONE exchange step: the ranks all-gather
the last bytes of their parts (halos) and, after the local reduce, the parts'
32-byte records; chaining the records in rank order gives every rank the state
its part is entered in."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import cuda

STREAM_HALO = cuda.STREAM_HALO
FN_BYTES = cuda.STREAM_FN_BYTES
UNKNOWN = cuda.STATE_UNKNOWN


def shard_range(n: int, rank: int, world: int):
    """contiguous [first, first+count) of n units for this rank"""
    per, rem = divmod(n, world)
    first = rank * per + min(rank, rem)
    return first, per + (1 if rank < rem else 0)


def _world():
    return dist.get_world_size() if dist.is_initialized() else 1


def _rank():
    return dist.get_rank() if dist.is_initialized() else 0


def allreduce_sum(value: int, device) -> int:
    t = torch.tensor([value], dtype=torch.int64, device=device)
    if _world() > 1:
        dist.all_reduce(t)
    return int(t.item())


_gather_bufs = {}


def allgather_bytes(mine: torch.Tensor):
    """all-gather a small uint8 tensor (same size on every rank) -> [world, n] tensor, rows in rank
    order (one collective into one cached buffer: the exchange sits inside the timed region)"""
    world = _world()
    if world == 1:
        return mine.unsqueeze(0)
    key = (mine.numel(), str(mine.device), world)
    out = _gather_bufs.get(key)
    if out is None:
        out = _gather_bufs[key] = torch.empty((world, mine.numel()), dtype=torch.uint8, device=mine.device)
    try:
        dist.all_gather_into_tensor(out.view(-1), mine.contiguous().view(-1))
    except (RuntimeError, NotImplementedError):
        dist.all_gather(list(out.unbind(0)), mine.contiguous().view(-1))
    return out


def chain_entries(fns, start_state: int, apply=None):
    """state in which every part begins (UNKNOWN behind a part whose record is still
    unresolved) and the state after the last part; fns in rank order"""
    apply = apply or cuda.fn_apply
    entries, s = [], start_state
    for f in fns:
        entries.append(s)
        s = apply(f, s) if s != UNKNOWN else UNKNOWN
    return entries, s


def first_match_global(local_offset: int, shard_base: int, device) -> int:
    """global offset of the first step that sees a match (-1: none)"""
    big = 1 << 62
    t = torch.tensor([shard_base + local_offset if local_offset >= 0 else big], dtype=torch.int64,
                     device=device)
    if _world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    v = int(t.item())
    return -1 if v == big else v


def stream_match_sharded(prog, shard: torch.Tensor, shard_len: int, shard_base: int, eof: bool = True,
                         apply=None):
    """Chunk-parallel Thompson match of one stream sharded across ranks (every part
    but the last at least STREAM_HALO bytes long).
    -> (rc, global offset of the first matching step or -1)"""
    from . import capi
    world, rank = _world(), _rank()
    start, acc = prog.info.dfa_start, prog.info.dfa_acc
    # 1. halos: what the next rank needs to know about the bytes in front of its part
    halo = torch.zeros(STREAM_HALO, dtype=torch.uint8, device=shard.device)
    if rank + 1 < world:
        assert shard_len >= STREAM_HALO, "a part that is not the last must hold at least one halo"
        halo.copy_(shard[shard_len - STREAM_HALO:shard_len])
    halos = allgather_bytes(halo)
    # 2. local reduce: the part as a record {possible entry state -> exit state}
    scan = prog.stream_reduce(shard, shard_len, halo=halos[rank - 1] if rank else None,
                              entry_state=cuda.STATE_INIT if rank == 0 else UNKNOWN)
    # 3. exchange the records; a rank whose entry state is known resolves its part.  One round
    # unless some part's record was unresolved (then its rank publishes {entry -> exit} and
    # the ranks behind it learn their entry states in the next round).
    entry, off, final = None, -1, UNKNOWN
    for _ in range(world):
        mine = torch.frombuffer(bytearray(scan.fn), dtype=torch.uint8).to(shard.device, non_blocking=True)
        gathered = allgather_bytes(mine).cpu().numpy()              # one read-back for all ranks' records
        fns = [gathered[r].tobytes() for r in range(world)]
        entries, final = chain_entries(fns, start, apply)
        if entry is None and entries[rank] != UNKNOWN:
            entry = entries[rank]
            _, off = scan.resolve(entry)
            if entry == acc:
                off = -1            # matched in an earlier part
        if final != UNKNOWN:
            break
    assert final != UNKNOWN and entry is not None
    scan.close()
    first = first_match_global(off, shard_base, shard.device)
    if final == acc:
        return capi.SRE_OK, first
    if eof:
        return (capi.SRE_OK if prog.dfa_fin(final) else capi.SRE_DECLINED), first
    return capi.SRE_AGAIN, first
