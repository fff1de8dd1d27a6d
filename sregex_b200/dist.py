"""Multi-GPU host logic: one process per GPU, torch.distributed for
the plumbing.  Line corpora shard with no data-path collective; a single long
stream shards into contiguous parts whose DFA transfer functions (<= 32 bytes
each) are all-gathered and composed in rank order."""
from __future__ import annotations

import torch
import torch.distributed as dist

ACC = 1     # absorbing "a step saw a live MATCH thread" state of the lowered DFA
FN_BYTES = 32


def shard_range(n: int, rank: int, world: int):
    """contiguous [first, first+count) of n units for this rank"""
    per, rem = divmod(n, world)
    first = rank * per + min(rank, rem)
    return first, per + (1 if rank < rem else 0)


def allreduce_sum(value: int, device) -> int:
    t = torch.tensor([value], dtype=torch.int64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return int(t.item())


def compose(fns, state: int) -> int:
    for f in fns:
        state = f[state]
    return state


def exchange_functions(fn: bytes, device):
    """all-gather this rank's transfer function; returns the list in rank order"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    mine = torch.zeros(FN_BYTES, dtype=torch.uint8, device=device)
    mine[: len(fn)] = torch.tensor(list(fn), dtype=torch.uint8)
    if world == 1:
        return [bytes(mine.cpu().tolist())]
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(t.cpu().tolist()) for t in out]


def stream_entry_state(fns, rank: int, start_state: int = 0) -> int:
    """state in which rank's shard begins = composition of the earlier shards"""
    return compose(fns[:rank], start_state)


def first_match_global(local_offset: int, shard_base: int, device) -> int:
    """global offset of the first step that sees a match (-1: none)"""
    big = 1 << 62
    t = torch.tensor([shard_base + local_offset if local_offset >= 0 else big], dtype=torch.int64,
                     device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    v = int(t.item())
    return -1 if v == big else v


def stream_match_sharded(prog, shard: torch.Tensor, shard_len: int, shard_base: int, eof: bool = True):
    """Chunk-parallel Thompson match of one stream sharded across ranks.
    -> (rc, global offset of the first matching step or -1)"""
    from . import capi
    fn = prog.stream_reduce(shard, shard_len)
    fns = exchange_functions(fn, shard.device)
    rank = dist.get_rank() if dist.is_initialized() else 0
    entry = stream_entry_state(fns, rank)
    exit_state, off = prog.stream_resolve(entry)
    if entry == ACC:
        off = -1            # matched in an earlier shard
    first = first_match_global(off, shard_base, shard.device)
    final = compose(fns, 0)
    if final == ACC:
        return capi.SRE_OK, first
    if eof:
        fin = bool(prog.lib.L.sre_cuda_dfa_fin(prog.cp, final))
        return (capi.SRE_OK if fin else capi.SRE_DECLINED), first
    return capi.SRE_AGAIN, first
