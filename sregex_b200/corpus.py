"""Deterministic synthetic corpora of SURVEY.md section 8(d), written with torch
integer ops so the same bytes come out on the CPU (tests) and on the GPU (bench).

* gen_data_buffer: the reference's bench/gen-data.pl:9 verbatim --
  "abccc" x repeat + "aaabbccb".
* log_lines: n fixed-pitch log-like lines,
    <ip> - - [18/Oct/2026:08:47:SS +0000] "<filler> <METHOD> /x/<n> HTTP/1.<v>" <status> ....
  filler from the 40-byte alphabet a-z0-9/_-. ; status 5xx with p = 0.1.
"""
from __future__ import annotations

import torch

BENCH_REGEX = rb'(?:a|b)aa(?:aa|bb)cc(?:a|b)'           # bench/Makefile:62
C2_REGEX = rb'HTTP/1\.[01]" 5\d\d '
C3_REGEX = rb'(\w+) (\S+) HTTP/(\d)\.(\d)'

_ALPHABET = b"abcdefghijklmnopqrstuvwxyz0123456789/_-."
_METHODS = [b"GET", b"POST", b"PUT", b"HEAD"]
_MASK = (1 << 63) - 1


def gen_data_buffer(repeat: int = 1048576, device="cpu") -> torch.Tensor:
    unit = torch.tensor(list(b"abccc"), dtype=torch.uint8, device=device)
    tail = torch.tensor(list(b"aaabbccb"), dtype=torch.uint8, device=device)
    return torch.cat([unit.repeat(repeat), tail])


def _mix(x: torch.Tensor) -> torch.Tensor:
    """splitmix64-style finaliser on int64 tensors (wrap-around arithmetic)."""
    x = (x ^ (x >> 30)) * -4658895280553007687      # 0xbf58476d1ce4e5b9
    x = (x ^ (x >> 27) & 0x1FFFFFFFFF) * -7723592293110705685  # 0x94d049bb133111eb
    return x ^ ((x >> 31) & 0x1FFFFFFFF)


def log_lines(n: int, pitch: int = 1024, device="cpu", seed: int = 0x5EED, first_line: int = 0,
              hit_rate: float = 0.1) -> torch.Tensor:
    """-> uint8 tensor [n, pitch]; line i depends only on (seed, first_line + i)."""
    dev = torch.device(device)
    idx = torch.arange(first_line, first_line + n, dtype=torch.int64, device=dev)
    h = _mix(idx * 2685821657736338717 + seed)      # xorshift64* multiplier as stride
    alphabet = torch.tensor(list(_ALPHABET), dtype=torch.uint8, device=dev)

    # filler everywhere first
    cols = torch.arange(pitch, dtype=torch.int64, device=dev)
    r = _mix(h[:, None] + cols[None, :] * 0x9E3779B97F4A7C15 % (1 << 63))
    lines = alphabet[((r >> 8) & 0xFFFF) % 40]

    def put(col, text_rows):
        """write text_rows [n, L] at per-line column col [n]"""
        L = text_rows.shape[1]
        cc = col[:, None] + torch.arange(L, device=dev)[None, :]
        lines.scatter_(1, cc, text_rows)

    def digits(v, width):
        p = torch.tensor([10 ** (width - 1 - k) for k in range(width)], dtype=torch.int64, device=dev)
        return ((v[:, None] // p[None, :]) % 10 + 48).to(torch.uint8)

    def const(text):
        return torch.tensor(list(text), dtype=torch.uint8, device=dev)[None, :].expand(n, -1)

    zero = torch.zeros(n, dtype=torch.int64, device=dev)
    # "ddd.ddd.ddd.ddd - - [18/Oct/2026:08:47:SS +0000] \""
    col = 0
    for k in range(4):
        octet = 100 + ((h >> (8 * k)) & 0xFF) % 156
        put(zero + col, digits(octet, 3))
        col += 3
        put(zero + col, const(b"." if k < 3 else b" "))
        col += 1
    head = b"- - [18/Oct/2026:08:47:"
    put(zero + col, const(head)); col += len(head)
    put(zero + col, digits(((h >> 33) & 0xFF) % 60, 2)); col += 2
    tail = b' +0000] "'
    put(zero + col, const(tail)); col += len(tail)

    # request tail: " METHOD /x/NNNNNN HTTP/1.V\" SSS " ending at a per-line column
    method = ((h >> 41) & 3)
    version = ((h >> 43) & 1)
    is5 = (((h >> 44) & 0xFFFF).to(torch.float64) / 65536.0) < hit_rate
    status = torch.where(is5, 500 + ((h >> 20) & 0xFF) % 100, zero + 200)
    num = ((h >> 12) & 0xFFFFF) % 1000000
    end = pitch - 40 - ((h >> 50) & 0xFF)           # where the padding starts
    for m, name in enumerate(_METHODS):
        sel = method == m
        if not bool(sel.any()):
            continue
        rows = torch.nonzero(sel)[:, 0]
        k = rows.numel()
        text = torch.cat([
            const(b" " + name + b" /x/")[:k], digits(num[rows], 6), const(b" HTTP/1.")[:k],
            (version[rows] + 48).to(torch.uint8)[:, None], const(b'" ')[:k], digits(status[rows], 3),
            const(b" ")[:k]], dim=1)
        L = text.shape[1]
        start = end[rows] - L
        cc = start[:, None] + torch.arange(L, device=dev)[None, :]
        lines[rows[:, None], cc] = text
    # '.' padding after the request
    pad = cols[None, :] >= end[:, None]
    lines[pad] = ord(".")
    return lines


def multi_pattern_set(n: int = 64):
    """BASELINE config 4: a set of n patterns for sre_regex_parse_multi --
    4 families x 16 keywords: literal keywords followed by
    digits, request lines, status codes, and keyword pairs with a gap."""
    words = ["alpha", "bravo", "charlie", "delta", "echo", "foxtrot", "golf", "hotel", "india", "juliet",
             "kilo", "lima", "mike", "november", "oscar", "papa"]
    methods = ["GET", "POST", "PUT", "HEAD"]
    pats = []
    for i, w in enumerate(words):
        pats.append(rb"%s\d+" % w.encode())                                   # keyword + digits
        pats.append(rb"%s /x/%d\d* HTTP/1\.[01]" % (methods[i % 4].encode(), i % 10))   # request lines
        pats.append(rb'" 5%02d ' % (i * 6))                                    # specific 5xx status
        pats.append(rb"%s[a-z]{0,3}%s" % (w[:3].encode(), words[(i + 5) % 16][:2].encode()))  # pair with a gap
    return pats[:n]
