"""GPU parity tests proper: libsregex_cuda (through its C ABI) against the
reference's golden vectors and against the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

from conftest import runnable
from oracle import cpu_baseline as baseline
from sregex_b200 import capi, corpus

pytestmark = pytest.mark.gpu





@pytest.fixture(scope="module")
def cu():
    from sregex_b200 import cuda
    assert cuda.lib().L.sre_cuda_device_available() == 1, "GPU tests need a CUDA device"
    return cuda


# ---- the reference's own API, driven like src/sre_cli.c ----------------------

def test_classic_thompson_golden(golden, cuda):
    for b in runnable(golden):
        p = cuda.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        assert cuda.thompson(p, b["subject_b"]) == b["thompson"], (b["file"], b["name"])
        p.close()


def test_classic_thompson_jit_api_golden(golden, cuda):
    for b in runnable(golden):
        p = cuda.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        assert cuda.thompson(p, b["subject_b"], jit=True) == b["thompson"], (b["file"], b["name"])
        p.close()


def test_classic_thompson_streaming_golden(golden, cuda):
    """1-byte chunks with SRE_AGAIN carry: whole rc sequence."""
    for b in runnable(golden):
        p = cuda.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        got = cuda.thompson(p, b["subject_b"], capi.split_chunks(b["subject_b"]))
        assert got == b["thompson_split"], (b["file"], b["name"])
        p.close()


def test_classic_pike_golden(golden, cuda):
    for b in runnable(golden):
        p = cuda.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        rc, ov = cuda.pike(p, b["subject_b"])
        assert (rc, ov) == (b["pike"]["rc"], b["pike"]["ov"]), (b["file"], b["name"])
        p.close()


def test_classic_pike_streaming_golden(golden, cuda):
    """1-byte chunks: final rc/ovector and the temp-capture / pending trace."""
    for b in runnable(golden):
        p = cuda.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        trace, rc, ov = cuda.pike(p, b["subject_b"], capi.split_chunks(b["subject_b"]))
        want = b["pike_split"]
        assert (rc, ov) == (want["rc"], want["ov"]), (b["file"], b["name"])
        assert [list(t) for t in trace] == want["trace"], (b["file"], b["name"])
        p.close()


# ---- batch entry points --------------------------------------------------------

def _golden_batch(golden, cu, engine):
    """every golden block as a 1-line batch through a given Thompson tier"""
    bad = []
    for b in runnable(golden):
        prog = cu.CudaProgram(b["regexes_b"], b["flags"], multi=b["multi"])
        if engine in (cu.ENGINE_DFA_TILED, cu.ENGINE_DFA_GENERIC) and prog.info.dfa_states == 0:
            continue
        if engine == cu.ENGINE_DFA_SKIP and prog.info.dfa_leave_bytes == 0:
            continue
        s = b["subject_b"]
        pitch = max(16, (len(s) + 15) // 16 * 16)
        buf = torch.zeros(pitch, dtype=torch.uint8)
        buf[: len(s)] = torch.frombuffer(bytearray(s), dtype=torch.uint8) if s else buf[:0]
        rc = prog.thompson_lines(buf.cuda(), 1, pitch, len(s), engine=engine)
        if int(rc[0]) != b["thompson"]:
            bad.append((b["file"], b["name"]))
        prog.program.close()
    assert not bad, bad[:10]


def test_batch_golden_dfa_tiled(golden, cu):
    _golden_batch(golden, cu, cu.ENGINE_DFA_TILED)


def test_batch_golden_dfa_generic(golden, cu):
    _golden_batch(golden, cu, cu.ENGINE_DFA_GENERIC)


def test_batch_golden_nfa(golden, cu):
    """(1-line batches shorter than 16 bytes take the warp kernel either way)"""
    _golden_batch(golden, cu, cu.ENGINE_NFA)


def test_batch_golden_nfa_warp(golden, cu):
    _golden_batch(golden, cu, cu.ENGINE_NFA_WARP)


def test_batch_golden_dfa_skip(golden, cu):
    """skip-scan tier on every block whose start state is left by <= 4 byte values"""
    _golden_batch(golden, cu, cu.ENGINE_DFA_SKIP)


@pytest.mark.parametrize("general_only", [0, 1, 2, 3])
def test_batch_golden_pike_with_start_hint(golden, cu, general_only):
    """every golden block as a 1-line batch through sre_cuda_pike_exec_lines with
    its internal gate + start-hint pass: rc and the whole ovector.  Once through
    the default tiers (0: the determinised Pike VM k_pike_lineage where the program
    has one, else the closure-table kernel), once through the general kernel alone
    (1), once through the walking shared-memory tier (k_pike_small, 2) and once
    through the closure-table tier for every program (k_pike_table, 3)."""
    _golden_pike_batch(golden, cu, general_only)


def _golden_pike_batch(golden, cu, tier_mode):
    bad = []
    for b in runnable(golden):
        prog = cu.CudaProgram(b["regexes_b"], b["flags"], multi=b["multi"])
        prog.set_pike_tier(tier_mode)
        s = b["subject_b"]
        pitch = max(16, (len(s) + 15) // 16 * 16)
        buf = torch.zeros(pitch, dtype=torch.uint8)
        if s:
            buf[: len(s)] = torch.frombuffer(bytearray(s), dtype=torch.uint8)
        rc, ov = prog.pike_lines(buf.cuda(), 1, pitch, len(s))
        want_rc, want_ov = b["pike"]["rc"], b["pike"]["ov"]
        got_ov = ov[0].tolist() if want_rc >= 0 else None
        if int(rc[0]) != want_rc or got_ov != want_ov:
            bad.append((b["file"], b["name"], int(rc[0]), got_ov, want_rc, want_ov))
        prog.program.close()
    assert not bad, bad[:5]


@pytest.mark.parametrize("nlines,linelen,pitch", [(4096, 1024, 1024), (1000, 1000, 1008), (33, 17, 32),
                                                  (5, 0, 16), (257, 63, 64), (64, 129, 144)])
def test_thompson_tiers_vs_oracle(cu, nlines, linelen, pitch):
    lines = corpus.log_lines(nlines, 1024)[:, 1024 - pitch:] if pitch <= 1024 else None
    lines = lines.contiguous()
    host = lines.numpy()
    _, want, _ = baseline.run_lines("oracle", corpus.C2_REGEX, None, host, nlines, pitch, linelen,
                                    baseline.ENGINE_THOMPSON)
    prog = cu.CudaProgram(corpus.C2_REGEX)
    dev = lines.cuda()
    for engine in (cu.ENGINE_DFA_TILED, cu.ENGINE_DFA_GENERIC, cu.ENGINE_NFA, cu.ENGINE_NFA_WARP, cu.ENGINE_DFA_SKIP,
                   cu.ENGINE_AUTO):
        variants = {cu.ENGINE_DFA_TILED: (0, 1, 2), cu.ENGINE_DFA_SKIP: (0, 1, 2)}.get(engine, (0,))
        for variant in variants:
            got = prog.thompson_lines(dev, nlines, pitch, linelen,
                                      engine=cu.engine_variant(engine, variant)).cpu().numpy()
            assert (got == want).all(), (engine, variant, int((got != want).sum()))
    if linelen == 1024:
        assert 0 < (want == 0).sum() < nlines


def test_thompson_ragged_vs_oracle(cu):
    rng = np.random.default_rng(0x5EED)
    lines = corpus.log_lines(512, 1024).numpy()
    lens = rng.integers(0, 1024, size=512)
    lens[:4] = [0, 1, 15, 16]
    chunks = [lines[i, 1024 - lens[i]:] for i in range(512)]
    flat = np.concatenate(chunks + [np.zeros(16, np.uint8)])
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    prog = cu.CudaProgram(corpus.C2_REGEX)
    want = []
    o = capi.load("oracle")
    po = o.compile(corpus.C2_REGEX)
    for c in chunks:
        want.append(o.thompson(po, c.tobytes()))
    for engine in (cu.ENGINE_AUTO, cu.ENGINE_NFA):
        got = prog.thompson_ragged(torch.from_numpy(flat).cuda(), torch.from_numpy(offsets).cuda(),
                                   engine=engine).cpu().numpy()
        assert got.tolist() == want, engine


def test_pike_lines_vs_oracle(cu):
    n = 2048
    lines = corpus.log_lines(n, 1024)
    host = lines.numpy()
    prog = cu.CudaProgram(corpus.C3_REGEX)
    _, want_rc, want_ov = baseline.run_lines("oracle", corpus.C3_REGEX, None, host, n, 1024, 1024,
                                             baseline.ENGINE_PIKE, ovec_slots=prog.nslots)
    rc, ov = prog.pike_lines(lines.cuda(), n, 1024, 1024)
    assert (rc.cpu().numpy() == want_rc).all()
    assert (ov.cpu().numpy() == want_ov).all()
    assert (want_rc == 0).all()             # every line holds a request line
    # gated by a Thompson pass: same answer
    sel = prog.thompson_lines(lines.cuda(), n, 1024, 1024)
    rc2, ov2 = prog.pike_lines(lines.cuda(), n, 1024, 1024, select=sel)
    assert torch.equal(rc, rc2) and torch.equal(ov, ov2)


def test_pike_ragged_lines_vs_oracle(cu):
    """ragged offsets: gate + start hint come from the class-table kernel"""
    rng = np.random.default_rng(7)
    lines = corpus.log_lines(300, 1024).numpy()
    lens = rng.integers(0, 1024, size=300)
    lens[:3] = [0, 1, 1023]
    chunks = [lines[i, 1024 - lens[i]:] for i in range(300)]
    flat = np.concatenate(chunks + [np.zeros(16, np.uint8)])
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    prog = cu.CudaProgram(corpus.C3_REGEX)
    o = capi.load("oracle")
    po = o.compile(corpus.C3_REGEX)
    rc, ov = prog.pike_lines(torch.from_numpy(flat).cuda(), 300, 0, 0, offsets=torch.from_numpy(offsets).cuda())
    rc, ov = rc.cpu().tolist(), ov.cpu().tolist()
    for i, c in enumerate(chunks):
        wrc, wov = o.pike(po, c.tobytes())
        assert rc[i] == wrc and (wov is None or ov[i] == wov), i
    assert sum(1 for r in rc if r == 0) > 50


MULTI = [rb"HTTP/1\.[01]\" 5\d\d ", rb"(GET|HEAD) /x/(\d+)", rb"POST /x/0", rb"^1[0-4]\d\.", rb"08:47:0(\d)",
         rb"zzz+q", rb"\bPUT\b", rb"[a-f]{6}\d"]


def test_multi_regex_id_vs_oracle(cu):
    n = 1024
    lines = corpus.log_lines(n, 1024)
    prog = cu.CudaProgram(MULTI)
    _, want_rc, want_ov = baseline.run_lines("oracle", MULTI, None, lines.numpy(), n, 1024, 1024,
                                             baseline.ENGINE_PIKE, ovec_slots=prog.nslots)
    _, want_bool, _ = baseline.run_lines("oracle", MULTI, None, lines.numpy(), n, 1024, 1024,
                                         baseline.ENGINE_THOMPSON)
    dev = lines.cuda()
    sel = prog.thompson_lines(dev, n, 1024, 1024)
    assert (sel.cpu().numpy() == want_bool).all()
    rc, ov = prog.pike_lines(dev, n, 1024, 1024, select=sel)
    assert (rc.cpu().numpy() == want_rc).all()
    assert (ov.cpu().numpy() == want_ov).all()
    assert len(set(want_rc.tolist())) > 3   # several different patterns win


def test_multi_regex_64_patterns_vs_oracle(cu):
    """BASELINE config 4 at test size: a 64-pattern sre_regex_parse_multi set;
    Thompson verdict, matched regex id and ovector against the oracle."""
    pats = corpus.multi_pattern_set(64)
    n = 384
    lines = corpus.log_lines(n, 1024)
    prog = cu.CudaProgram(pats)
    assert prog.info.nregexes == 64 and prog.info.nfa_states > 500
    which = "ref" if baseline.available("ref") else "oracle"
    _, want_rc, want_ov = baseline.run_lines(which, pats, None, lines.numpy(), n, 1024, 1024,
                                             baseline.ENGINE_PIKE, nthreads=8, ovec_slots=prog.nslots)
    dev = lines.cuda()
    for engine in (cu.ENGINE_AUTO, cu.ENGINE_NFA):
        sel = prog.thompson_lines(dev, n, 1024, 1024, engine=engine)
        assert ((sel.cpu().numpy() == 0) == (want_rc >= 0)).all(), engine
    rc, ov = prog.pike_lines(dev, n, 1024, 1024, select=sel)
    assert (rc.cpu().numpy() == want_rc).all()
    assert (ov.cpu().numpy() == want_ov).all()
    rc2, ov2 = prog.pike_lines(dev, n, 1024, 1024)          # internal gate
    assert torch.equal(rc, rc2) and torch.equal(ov, ov2)
    assert len(set(want_rc.tolist())) > 8
    # the set runs on the determinised Pike VM, not on a fallback
    assert prog.last_pike_tier() == 3
    # ... and the closure-table kernel and the general kernel alone give the same rows
    for mode, tier in ((3, 0), (1, 1)):
        prog.set_pike_tier(mode)
        rc3, ov3 = prog.pike_lines(dev, n, 1024, 1024)
        assert prog.last_pike_tier() == tier
        assert torch.equal(rc, rc3) and torch.equal(ov, ov3)
    prog.set_pike_tier(0)
    assert torch.equal(rc, rc3) and torch.equal(ov, ov3)


def test_pike_many_groups_on_the_table_tier(cu):
    """a 14-group log regex (30 capture slots): full ovectors against the oracle,
    on the determinised Pike VM and on the closure-table kernel"""
    from test_lowering import WIDE_REGEX
    n = 512
    lines = corpus.log_lines(n, 1024)
    prog = cu.CudaProgram(WIDE_REGEX)
    _, want_rc, want_ov = baseline.run_lines("oracle", WIDE_REGEX, None, lines.numpy(), n, 1024, 1024,
                                             baseline.ENGINE_PIKE, nthreads=8, ovec_slots=prog.nslots)
    for mode, tier in ((0, 3), (3, 0)):
        prog.set_pike_tier(mode)
        rc, ov = prog.pike_lines(lines.cuda(), n, 1024, 1024)
        assert prog.last_pike_tier() == tier
        assert (rc.cpu().numpy() == want_rc).all() and int((rc == 0).sum()) == n
        assert (ov.cpu().numpy() == want_ov).all()


def test_big_regex_set_with_assertions_vs_oracle(cu):
    """a set of 40 random regexes rich in look-ahead / look-behind assertions
    (> 64 parked instructions: bit-set marks in shared memory, pending
    look-ahead closures, per-context closure tables) over random short lines:
    matched id + ovector == the oracle's Pike, on the default tier and on the table tier"""
    import random
    rng = random.Random(99)
    atoms = ["a", "b", "ab", " ", "_", "1", ".", "^", "$", "\\b", "\\B", "\\z", "\\A", "(a)", "(b+)", "(?:ab)*", "+", "?",
             "[ab]", "[^a ]", "\\w", "\\d", "(\\w+)", "|"]
    oracle = capi.load("oracle")
    pats = []
    while len(pats) < 40:
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(2, 7))).encode()
        try:
            oracle.compile(rx, 0).close()
        except capi.SreSyntaxError:
            continue
        pats.append(rx)
    prog = cu.CudaProgram(pats)
    assert prog.info.nregexes == 40
    n, linelen, pitch = 4096, 24, 32
    alphabet = np.frombuffer(b"ab _1.\n", dtype=np.uint8)
    rs = np.random.RandomState(7)
    lines = np.zeros((n, pitch), dtype=np.uint8)
    lines[:, :linelen] = alphabet[rs.randint(0, len(alphabet), size=(n, linelen))]
    _, want_rc, want_ov = baseline.run_lines("oracle", pats, None, lines, n, pitch, linelen,
                                             baseline.ENGINE_PIKE, nthreads=8, ovec_slots=prog.nslots)
    dev = torch.from_numpy(lines).cuda()
    # the default tier (the determinised Pike VM when the set has one, look-ahead threads parked
    # in its lists) and the closure-table tier
    for mode, tiers in ((0, (0, 3)), (3, (0,))):
        prog.set_pike_tier(mode)
        rc, ov = prog.pike_lines(dev, n, pitch, linelen)
        assert prog.last_pike_tier() in tiers
        assert (rc.cpu().numpy() == want_rc).all(), mode
        assert (ov.cpu().numpy() == want_ov).all(), mode
    assert len(set(want_rc.tolist())) > 5


def test_index_lines_and_ragged_grep(cu):
    """a '\\n'-delimited buffer -> device line index -> ragged Thompson + Pike:
    offsets against a host scan, verdicts / ovectors against the oracle per line"""
    rs = np.random.RandomState(11)
    oracle = capi.load("oracle")
    for trailing_newline in (True, False):
        src = corpus.log_lines(3000, 1024).numpy()
        lens = rs.randint(0, 300, size=3000)
        lens[::97] = 0                                  # empty lines
        parts = [bytes(src[i, 1024 - 60 - lens[i]: 1024 - 60]) + b"\n" for i in range(3000)]
        data = b"".join(parts)
        if not trailing_newline:
            data += b"GET /x/1 HTTP/1.1\" 503 "        # last line without terminator
        host = np.frombuffer(data, dtype=np.uint8)
        want_off = np.concatenate([[0], np.flatnonzero(host == 10) + 1]).astype(np.int64)
        if not trailing_newline:
            want_off = np.concatenate([want_off, [len(data)]])
        dev = torch.from_numpy(host.copy()).cuda()
        off = cu.index_lines(dev)
        assert np.array_equal(off.cpu().numpy(), want_off)
        # too small an array: the count is still reported, the prefix is right
        off_small = cu.index_lines(dev, max_lines=100)
        assert np.array_equal(off_small.cpu().numpy(), want_off[:101])
        n = len(want_off) - 1
        prog = cu.CudaProgram(corpus.C2_REGEX)
        got = prog.thompson_ragged(dev, off).cpu().numpy()
        po = oracle.compile(corpus.C2_REGEX, 0)
        want = np.array([oracle.thompson(po, data[want_off[i]:want_off[i + 1]]) for i in range(n)])
        assert np.array_equal(got, want)
        assert (want == 0).sum() > 10
        po.close()
    empty = cu.index_lines(torch.empty(0, dtype=torch.uint8, device="cuda"))
    assert empty.cpu().tolist() == [0]


def test_long_lines_32bit_captures(cu):
    """lines of 40,000 bytes (beyond the 16-bit capture offsets of the table
    kernel): Thompson verdicts and full Pike ovectors against the oracle"""
    n, linelen = 48, 40000
    rs = np.random.RandomState(3)
    alphabet = np.frombuffer(b"abcdefghij0123456789/_-. ", dtype=np.uint8)
    lines = alphabet[rs.randint(0, len(alphabet), size=(n, linelen))].copy()
    needle = np.frombuffer(b' GET /x/42 HTTP/1.1" 503 ', dtype=np.uint8)
    for i in range(0, n, 2):                                   # every other line matches, far into the line
        at = 33000 + 97 * i
        lines[i, at:at + len(needle)] = needle
    for rx in (corpus.C2_REGEX, corpus.C3_REGEX):
        prog = cu.CudaProgram(rx)
        _, want_t, _ = baseline.run_lines("oracle", rx, None, lines, n, linelen, linelen, baseline.ENGINE_THOMPSON)
        _, want_rc, want_ov = baseline.run_lines("oracle", rx, None, lines, n, linelen, linelen,
                                                 baseline.ENGINE_PIKE, ovec_slots=prog.nslots)
        dev = torch.from_numpy(lines).cuda()
        assert (prog.thompson_lines(dev, n, linelen, linelen).cpu().numpy() == want_t).all()
        rc, ov = prog.pike_lines(dev, n, linelen, linelen)
        assert (rc.cpu().numpy() == want_rc).all() and (ov.cpu().numpy() == want_ov).all()
        assert int((rc == 0).sum()) >= n // 2
        assert int(ov.max()) > 32767
        # the same lines as a ragged batch with a (wrong) caller's bound of 500 bytes: lines
        # beyond what 16-bit capture offsets can hold are handed to the next tier, not mangled
        off = torch.arange(0, (n + 1) * linelen, linelen, dtype=torch.int64, device="cuda")
        rc2, ov2 = prog.pike_lines(dev.view(-1), n, 0, 500, offsets=off)
        assert torch.equal(rc, rc2) and torch.equal(ov, ov2)


@pytest.mark.parametrize("rx", [corpus.C2_REGEX, rb'[Hh]TTP/1\.[01]" (5\d\d) ', rb'(GET|PUT) (/x/\d+) '])
def test_pike_with_word_skipping_hint_pass(cu, rx):
    """literal-prefixed regexes (start state left by 1 or 2 byte values; the third
    one by more, so it takes the plain hint kernel): the gate + start hint pass
    skips words, and rc + ovector still equal the oracle's on 1 KB lines"""
    n = 4096
    lines = corpus.log_lines(n, 1024)
    prog = cu.CudaProgram(rx)
    _, want_rc, want_ov = baseline.run_lines("oracle", rx, None, lines.numpy(), n, 1024, 1024,
                                             baseline.ENGINE_PIKE, nthreads=8, ovec_slots=prog.nslots)
    rc, ov = prog.pike_lines(lines.cuda(), n, 1024, 1024)
    assert (rc.cpu().numpy() == want_rc).all()
    assert (ov.cpu().numpy() == want_ov).all()
    assert 0 < int((rc == 0).sum()) < n


def test_pike_tier_selection(cu):
    """the configurations the bench reports must run on the fast tier: C3 (4
    groups, 10 slots, 1 KB lines) on the determinised Pike VM"""
    n = 256
    dev = corpus.log_lines(n, 1024).cuda()
    prog = cu.CudaProgram(corpus.C3_REGEX)
    rc, _ = prog.pike_lines(dev, n, 1024, 1024)
    assert prog.last_pike_tier() == 3
    assert int((rc == 0).sum()) == n
    # assertions are part of the determinised Pike VM: look-behind ones through the start list and
    # the kind of the byte in front of a state, look-ahead ones as threads parked on the assertion
    # (resolved by the next byte or at the end of the line).  Same rows as the oracle, on lines
    # with inner newlines.
    o = capi.load("oracle")
    lines = corpus.log_lines(n, 1024).numpy().copy()
    lines[:, 200] = 10
    lines[::2, 201:204] = np.frombuffer(b"GET", dtype=np.uint8)
    dev = torch.from_numpy(lines).cuda()
    la = 3 if os.environ.get("SRE_PDFA_LOOKAHEAD", "1") != "0" else 0     # (see sre_cuda_pike_exec_lines)
    for rx, tier in ((rb"^(\d+)\.(\d+)", 3), (rb"^(GET|\d+)(.)", 3), (rb"\A(\d+)", 3), (rb"(\w+)$", la), (rb"\b(GET)\b", la),
                     (rb"(\d+)\.*$", la), (rb"^(\S+) .*\B(\.)\z", la), (rb"(\w+)\b (\S+)$", la)):
        prog = cu.CudaProgram(rx)
        rc, ov = prog.pike_lines(dev, n, 1024, 1024)
        assert prog.last_pike_tier() == tier, (rx, prog.last_pike_tier())
        _, wrc, wov = baseline.run_lines("oracle", rx, None, lines, n, 1024, 1024, baseline.ENGINE_PIKE,
                                         ovec_slots=prog.nslots)
        assert (rc.cpu().numpy() == wrc).all() and (ov.cpu().numpy() == wov).all(), rx
    assert o is not None


def test_pike_lineage_long_matches_fall_through(cu):
    """matches longer than the lineage kernel's 64-position ring: the lane runs the automaton
    forward again up to the position its walk back needs (up to 64 times, i.e. lineages of ~4 KB),
    beyond that the line is handed to the closure-table kernel -- same rows as the oracle either
    way, for a small automaton (byte ring) and a larger one (16-bit ring)"""
    big = rb"(x+)(y*) (\d+)|(GET|HEAD|POST|PUT|DELETE|OPTIONS|PATCH|TRACE|CONNECT) (\S+) HTTP/(\d)|(\w+)=(\d+);"
    for rx, n, linelen, longest in ((rb"(x+)(y*) (\d+)", 512, 256, 200), (rb"(x+)(y*) (\d+)", 96, 6000, 5900),
                                    (big, 256, 1024, 1000), (rb"\.(x+)(y*)\b.*?(\d+)\b\.*$", 256, 512, 400)):
        rs = np.random.RandomState(9)
        lines = np.full((n, linelen), ord("."), dtype=np.uint8)
        for i in range(n):
            k = int(rs.randint(1, longest))
            at = int(rs.randint(0, 20))
            body = b"x" * k + b"y" * int(rs.randint(0, 10)) + b" " + b"7" * int(rs.randint(1, 5))
            body = body[: linelen - at]
            lines[i, at:at + len(body)] = np.frombuffer(body, dtype=np.uint8)
        prog = cu.CudaProgram(rx)
        _, want_rc, want_ov = baseline.run_lines("oracle", rx, None, lines, n, linelen, linelen, baseline.ENGINE_PIKE,
                                                 nthreads=8, ovec_slots=prog.nslots)
        rc, ov = prog.pike_lines(torch.from_numpy(lines).cuda(), n, linelen, linelen)
        assert prog.last_pike_tier() == 3, rx
        assert (rc.cpu().numpy() == want_rc).all() and (ov.cpu().numpy() == want_ov).all(), (rx, linelen)
        assert int((want_ov[:, 1] - want_ov[:, 0] > 70).sum()) > n // 5 and int((want_rc >= 0).sum()) > n // 2, rx


def test_stream_scan_vs_oracle(cu):
    """bench/gen-data.pl buffer (scaled down) fed in chunks: same rc sequence as
    the oracle's sre_vm_thompson_exec with SRE_AGAIN carry."""
    buf = corpus.gen_data_buffer(40000)          # 200,008 bytes, match at the tail
    data = bytes(buf.numpy())
    prog = cu.CudaProgram(corpus.BENCH_REGEX)
    o = capi.load("oracle")
    po = o.compile(corpus.BENCH_REGEX)
    dev = buf.cuda()
    for chunk in (65536, 4096, 1000):
        pieces = [(data[i:i + chunk], i + chunk >= len(data)) for i in range(0, len(data), chunk)]
        want = o.thompson(po, data, pieces)
        rc, state, mchunk = prog.thompson_stream(dev, len(data), chunk, True)
        assert rc == want[-1] == capi.SRE_OK
        assert mchunk == len(want) - 1, (chunk, mchunk, len(want))
        # carried across two calls
        cut = (len(data) // 2) // 16 * 16
        rc1, st1, _ = prog.thompson_stream(dev, cut, chunk, False)
        assert rc1 == capi.SRE_AGAIN
        rc2, _, _ = prog.thompson_stream(dev[cut:], len(data) - cut, chunk, True, state=st1)
        assert rc2 == capi.SRE_OK
    # no-match variant
    nomatch = corpus.gen_data_buffer(40000)[:-8].contiguous()
    rc, _, _ = prog.thompson_stream(nomatch.cuda(), nomatch.numel(), 65536, True)
    assert rc == capi.SRE_DECLINED == o.thompson(po, bytes(nomatch.numpy()))
    # the same stream from host memory, in slices that are copied while the previous one is scanned
    for slice_bytes in (0, 65536, 8192):
        rc, _, mchunk = prog.thompson_stream_host(buf, len(data), 4096, True, slice_bytes=slice_bytes)
        assert rc == capi.SRE_OK and mchunk == (len(data) - 1) // 4096
        rc, _, _ = prog.thompson_stream_host(nomatch, nomatch.numel(), 4096, True, slice_bytes=slice_bytes)
        assert rc == capi.SRE_DECLINED
    early = torch.cat([torch.tensor(list(b"xx aaabbccbx"), dtype=torch.uint8), buf])
    rc, _, mchunk = prog.thompson_stream_host(early, early.numel(), 4096, True, slice_bytes=8192)
    assert rc == capi.SRE_OK and mchunk == 0            # stops at the first slice
    # multi-GPU building blocks: reduce a part to a record, then resolve it
    # (the match ends on the last byte, so it is the EOF step that sees it:
    #  the exit state is not ACC and there is no in-stream match offset)
    acc = prog.info.dfa_acc
    scan = prog.stream_reduce(dev, len(data), entry_state=cu.STATE_INIT)
    ex, off = scan.resolve(cu.STATE_INIT)
    assert cu.fn_apply(scan.fn, prog.info.dfa_start) == ex != acc and off == -1
    scan.close()
    longer = torch.cat([buf, torch.tensor(list(b"xyz"), dtype=torch.uint8)]).cuda()
    scan = prog.stream_reduce(longer, longer.numel(), entry_state=cu.STATE_INIT)
    ex, off = scan.resolve(cu.STATE_INIT)
    assert cu.fn_apply(scan.fn, prog.info.dfa_start) == ex == acc and off == len(data)   # the step on 'x' sees the MATCH thread
    scan.close()
    # a part whose entry state is not known: candidates from the halo in front of it
    cut = 30 * 4096 + 5 * 16
    halo = dev[cut - cu.STREAM_HALO:cut].clone()
    first = prog.stream_reduce(dev, cut, entry_state=cu.STATE_INIT)
    second = prog.stream_reduce(longer[cut:], longer.numel() - cut, halo=halo)
    mid = cu.fn_apply(first.fn, prog.info.dfa_start)
    assert mid not in (acc, cu.STATE_UNKNOWN)
    assert cu.fn_apply(second.fn, mid) == acc
    ex2, off2 = second.resolve(mid)
    assert ex2 == acc and cut + off2 == len(data)
    first.close()
    second.close()


import re as _re
_LOOKAHEAD = _re.compile(rb"\$|\\[bBz]")         # assertions that wait for the next byte
_LOOKBEHIND = _re.compile(rb"\^|\\[AbB]")        # assertions that look at the previous one


def _chunked_oracle(o, po, data, chunk):
    pieces = [(data[i:i + chunk], i + chunk >= len(data)) for i in range(0, len(data), chunk)] or [(b"", True)]
    rcs = o.thompson(po, data, pieces)
    return rcs[-1], len(rcs) - 1


@pytest.mark.parametrize("name", ["c3", "multi8", "multi64"])
def test_stream_scan_large_automata(cu, name):
    """the stream scan on programs beyond the old 32-state limit: C3's regex, the
    8-pattern set and the 64-pattern set (2107 DFA states), over log text in 64 KB,
    4 KB and 1000-byte chunks: last rc and the chunk in which the reference's call
    sequence returns SRE_OK (oracle: sre_vm_thompson_exec fed the same chunks)"""
    rx = {"c3": corpus.C3_REGEX, "multi8": MULTI, "multi64": corpus.multi_pattern_set(64)}[name]
    prog = cu.CudaProgram(rx)
    assert prog.info.dfa_states > (32 if name != "c3" else 16) and prog.info.image_states > 0
    o = capi.load("oracle")
    po = o.compile(rx)
    lines = corpus.log_lines(320, 1024, hit_rate=0.0).numpy()
    quiet = lines.tobytes()
    for w in (b"GET", b"POST", b"PUT", b"HEAD", b"HTTP", b"08:47:0", b"10", b"11", b"12", b"13", b"14"):
        quiet = quiet.replace(w, b"~" * len(w))
    texts = [lines.tobytes(), quiet, quiet[:200000] + b' GET /x/123456 HTTP/1.1" 503 ' + quiet[200000:210000]]
    for data in texts:
        dev = torch.frombuffer(bytearray(data) + bytearray(16), dtype=torch.uint8).cuda()
        for chunk in (65536, 4096, 1000):
            want_rc, want_idx = _chunked_oracle(o, po, data, chunk)
            rc, state, mchunk = prog.thompson_stream(dev, len(data), chunk, True)
            assert rc == want_rc, (name, chunk)
            if rc == capi.SRE_OK:
                assert mchunk == want_idx, (name, chunk, mchunk, want_idx)
        # carried across two calls at an odd (16-byte aligned) cut
        cut = (len(data) // 3) // 16 * 16
        rc1, st1, _ = prog.thompson_stream(dev, cut, 4096, False)
        if rc1 == capi.SRE_AGAIN:
            rc2, _, _ = prog.thompson_stream(dev[cut:], len(data) - cut, 4096, True, state=st1)
            assert rc2 == o.thompson(po, data)


def test_stream_scan_random_regex_fuzz(cu):
    """>= 150 random regexes (with ^ $ \\b \\B \\A \\z) over random text of a few
    pieces: last rc and match chunk for 64 KB, 4 KB and 1000-byte chunks against the
    oracle's sre_vm_thompson_exec fed the same chunks, chunk edges falling anywhere"""
    import random
    rng = random.Random(4242)
    atoms = ["a", "b", "ab", " ", "\\n", "_", ".", "^", "$", "\\b", "\\B", "\\A", "\\z", "|", "(", ")", "(?:", "*",
             "+", "?", "*?", "+?", "??", "{2}", "{0,2}", "{1,}", "[ab]", "[^a]", "\\w", "\\W", "\\s", "\\d", "1"]
    alphabet = b"ab \n_1."
    o = capi.load("oracle")
    done = late = 0
    diverged = set()
    while done < 160:
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(2, 9))).encode()
        try:
            po = o.compile(rx, 0)
        except capi.SreSyntaxError:
            continue
        prog = cu.CudaProgram(rx)
        if not prog.info.dfa_states:
            po.close()
            continue
        done += 1
        n = rng.choice([0, 1, 100, 4096, 4097, 9000, 20000, 70000])
        # text that does not match early: of a few reduced alphabets keep the one whose first
        # match comes latest; then plant bytes of the full alphabet late in the text
        best = None
        for _ in range(4 if n > 200 else 1):
            alpha = bytes(rng.sample(list(alphabet), rng.randrange(1, 5)))
            body = bytearray(rng.choice(alpha) for _ in range(n))
            rc0, idx0 = _chunked_oracle(o, po, bytes(body), 1000)
            when = idx0 if rc0 == capi.SRE_OK and idx0 < n // 1000 else 1 << 30
            if best is None or when > best[0]:
                best = (when, body)
        body = best[1]
        if n > 200 and rng.random() < 0.7:
            at = rng.randrange(n // 2, n - 30)
            body[at:at + 24] = bytes(rng.choice(alphabet) for _ in range(24))
        data = bytes(body)
        dev = torch.frombuffer(bytearray(data) + bytearray(16), dtype=torch.uint8).cuda()
        for chunk in (65536, 4096, 1000):
            want_rc, want_idx = _chunked_oracle(o, po, data, chunk)
            rc, state, mchunk = prog.thompson_stream(dev, len(data), chunk, True)
            if (rc, mchunk if rc == capi.SRE_OK else 0) != (want_rc, want_idx if want_rc == capi.SRE_OK else 0):
                # The one documented divergence: the reference's
                # interpreter tests `sp == ctx->buffer` per CHUNK (sre_vm_thompson.c:79, 303,
                # 312, 321), so a look-ahead assertion held over a chunk edge and followed by a
                # look-behind one sees "start of input" at offset 0 of a later chunk.  The GPU
                # path (like the reference's JIT) behaves as on one big buffer.
                assert _LOOKAHEAD.search(rx) and _LOOKBEHIND.search(rx), (rx, n, chunk, rc, want_rc)
                one_rc, one_idx = _chunked_oracle(o, po, data, max(len(data), 1))
                assert rc == one_rc, (rx, n, chunk, rc, one_rc)
                diverged.add(rx)
                continue
            if rc == capi.SRE_OK:
                late += want_idx > 0
        # without eof: AGAIN or OK, and the carried state continues exactly
        if n >= 4096:
            cut = rng.randrange(1, n // 16) * 16
            rc1, st1, _ = prog.thompson_stream(dev, cut, 4096, False)
            want1 = o.thompson(po, data, [(data[:cut], False)])[-1]
            assert rc1 == want1, (rx, n, cut)
            if rc1 == capi.SRE_AGAIN:
                rc2, _, _ = prog.thompson_stream(dev[cut:], n - cut, 4096, True, state=st1)
                assert rc2 == o.thompson(po, data), (rx, n, cut)
        prog.program.close()
        po.close()
    assert late >= 10, late
    assert len(diverged) <= 8, diverged


def test_stream_scan_unresolved_pieces_are_repaired(cu):
    """a regex whose automaton does not forget ([ab]{5} after an a) over text made of
    a and b only: no window narrows the candidates, every piece is unresolved and
    goes through the repair path -- still exact"""
    rx = rb"a[ab]{5}c"
    prog = cu.CudaProgram(rx)
    o = capi.load("oracle")
    po = o.compile(rx)
    rs = np.random.RandomState(5)
    body = rs.choice(np.frombuffer(b"ab", dtype=np.uint8), size=40000)
    for plant in (None, 33333):
        data = bytearray(body.tobytes())
        if plant:
            data[plant:plant + 7] = b"abbabac"
        data = bytes(data)
        dev = torch.frombuffer(bytearray(data) + bytearray(16), dtype=torch.uint8).cuda()
        want_rc, want_idx = _chunked_oracle(o, po, data, 4096)
        rc, _, mchunk = prog.thompson_stream(dev, len(data), 4096, True)
        assert rc == want_rc and (rc != capi.SRE_OK or mchunk == want_idx)


def test_classic_exec_large_buffer(cuda):
    """sre_vm_thompson_exec on a 5 MB buffer (the reference's bench config):
    takes the chunk-parallel path inside the drop-in entry point."""
    data = bytes(corpus.gen_data_buffer(1048576).numpy())
    p = cuda.compile(corpus.BENCH_REGEX)
    assert cuda.thompson(p, data) == capi.SRE_OK
    assert cuda.thompson(p, data[:-8]) == capi.SRE_DECLINED
    half = len(data) // 2
    assert cuda.thompson(p, data, [(data[:half], False), (data[half:], True)]) == [capi.SRE_AGAIN, capi.SRE_OK]
    # sre_vm_pike_exec over the whole buffer: the scan locates the match, the Pike VM runs from
    # the last position before it at which no thread was alive (bench/sregex.c:330-334)
    rc, ov = cuda.pike(p, data)
    assert (rc, ov) == (0, [len(data) - 8, len(data)])
    assert cuda.pike(p, data[:-8])[0] == capi.SRE_DECLINED
    mid = data[:3000000] + b"baabbccaxyz" + data[3000000:]        # an earlier match in the middle
    assert cuda.pike(p, mid) == (0, [3000000, 3000008])
    o = capi.load("oracle")
    for rx, text in ((rb"(c+)(a|b)(\w*)", data[:300000]), (rb"^(abc+)+$", data[:100000 - 3]),
                     (rb"\b(\w+)\b", b"." * 200000 + b"word" + b"." * 100000), (rb"x*", data[:70000])):
        po, pc = o.compile(rx), cuda.compile(rx)
        assert cuda.pike(pc, text) == o.pike(po, text), rx
    # a program beyond the old 32-state limit takes the same path (C3's regex: no match in this text)
    p3 = cuda.compile(corpus.C3_REGEX)
    assert cuda.thompson(p3, data) == capi.SRE_DECLINED
    assert cuda.thompson(p3, data + b" GET /x/1 HTTP/1.1 ") == capi.SRE_OK


def test_full_size_properties(cu):
    """BASELINE config 2 at full size (1M x 1 KB): the verdict of every line
    must equal what the generator planted (5xx status <=> match), and every
    tier must agree."""
    n = 1 << 20
    dev = corpus.log_lines(n, 1024, device="cuda")
    prog = cu.CudaProgram(corpus.C2_REGEX)
    rc = prog.thompson_lines(dev, n, 1024, 1024)
    # status field: 3 digits before the final space of the request
    is5 = torch.zeros(n, dtype=torch.bool, device="cuda")
    # find '" ' + '5' pattern directly with tensor ops (independent of the library)
    q = (dev[:, :-3] == ord('"')) & (dev[:, 1:-2] == ord(' ')) & (dev[:, 2:-1] == ord('5'))
    is5 = q.any(dim=1)
    assert torch.equal(rc == 0, is5)
    assert 0.08 < float(is5.float().mean()) < 0.12
    rc2 = prog.thompson_lines(dev, n, 1024, 1024, engine=cu.ENGINE_DFA_GENERIC)
    assert torch.equal(rc, rc2)
    sub = 1 << 16
    rc3 = prog.thompson_lines(dev, sub, 1024, 1024, engine=cu.ENGINE_NFA)
    assert torch.equal(rc[:sub], rc3)


def test_random_regex_fuzz_gpu_vs_oracle(cu):
    """Seeded random regexes (assertions, nested repetition, classes) x batches of
    random short lines: every Thompson tier and the Pike path (rc + ovector)
    against the CPU oracle."""
    import random
    rng = random.Random(99)
    atoms = ["a", "b", "ab", " ", "\\n", "_", ".", "^", "$", "\\b", "\\B", "\\A", "\\z", "|", "(", ")", "(?:", "*",
             "+", "?", "*?", "+?", "??", "{2}", "{0,2}", "{1,}", "[ab]", "[^a]", "\\w", "\\W", "\\s", "\\d", "1"]
    alphabet = list(b"ab \n_1.")
    o = capi.load("oracle")
    done = 0
    nlines, pitch = 96, 32
    while done < 150:
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(1, 8))).encode()
        try:
            po = o.compile(rx, 0)
        except capi.SreSyntaxError:
            continue
        done += 1
        linelen = rng.choice([0, 1, 7, 16, 17, 31, 32])
        host = np.array([rng.choice(alphabet) for _ in range(nlines * pitch)], dtype=np.uint8).reshape(nlines, pitch)
        prog = cu.CudaProgram(rx)
        _, want_t, _ = baseline.run_lines("oracle", rx, None, host, nlines, pitch, linelen, baseline.ENGINE_THOMPSON)
        _, want_rc, want_ov = baseline.run_lines("oracle", rx, None, host, nlines, pitch, linelen,
                                                 baseline.ENGINE_PIKE, ovec_slots=prog.nslots)
        dev = torch.from_numpy(host).cuda()
        engines = [cu.ENGINE_AUTO, cu.ENGINE_NFA, cu.ENGINE_NFA_WARP]
        if prog.info.dfa_states:
            engines += [cu.ENGINE_DFA_TILED, cu.ENGINE_DFA_GENERIC]
        if prog.info.dfa_leave_bytes:
            engines.append(cu.ENGINE_DFA_SKIP)
        for e in engines:
            got = prog.thompson_lines(dev, nlines, pitch, linelen, engine=e).cpu().numpy()
            assert (got == want_t).all(), (rx, linelen, e)
        rc, ov = prog.pike_lines(dev, nlines, pitch, linelen)
        assert (rc.cpu().numpy() == want_rc).all(), (rx, linelen)
        assert (ov.cpu().numpy() == want_ov).all(), (rx, linelen)
        prog.program.close()
        po.close()


def test_unaligned_and_padded_lines(cu):
    """base pointer off by one (generic tier), pitch > linelen with junk between lines"""
    n = 200
    lines = corpus.log_lines(n, 1024)
    prog = cu.CudaProgram(corpus.C2_REGEX)
    _, want, _ = baseline.run_lines("oracle", corpus.C2_REGEX, None, lines.numpy(), n, 1024, 1000,
                                    baseline.ENGINE_THOMPSON)
    shifted = torch.cat([torch.zeros(1, dtype=torch.uint8), lines.view(-1)]).cuda()
    got = prog.thompson_lines(shifted[1:], n, 1024, 1000)
    assert (got.cpu().numpy() == want).all()
    got = prog.thompson_lines(lines.cuda(), n, 1024, 1000)
    assert (got.cpu().numpy() == want).all()


def test_global_scan_classic_and_batch(golden, cuda, cu):
    """all non-overlapping matches: the classic ctx continuation on the GPU and the
    batch entry point sre_cuda_pike_exec_lines_all, against the oracle"""
    o = capi.load("oracle")
    for b in runnable(golden)[::4]:
        po = o.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        pc = cuda.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        s = b["subject_b"] * 3
        assert capi.pike_all(cuda, pc, s) == capi.pike_all(o, po, s), (b["file"], b["name"])
        po.close()
        pc.close()
    # batch: a few lines of log text, several matches per line
    rx = rb"(\d+)\.|HTTP"
    n, K = 64, 12
    lines = corpus.log_lines(n, 1024)
    prog = cu.CudaProgram(rx)
    count, spans, ids = prog.pike_lines_all(lines.cuda(), n, 1024, 1024, K)
    po = o.compile(rx)
    for i in range(n):
        want = capi.pike_all(o, po, bytes(lines[i].numpy()), limit=K)
        assert int(count[i]) == len(want), i
        got = [(int(ids[i, k]), int(spans[i, k, 0]), int(spans[i, k, 1])) for k in range(len(want))]
        assert got == want, i
    assert int(count.max()) >= 5


def test_one_program_from_two_threads_and_two_streams(cu):
    """re-entrancy at the boundary (the reference's JIT run time is re-entrant): ONE
    lowered program driven from 2 host threads x 2 CUDA streams at once -- Thompson lines, Pike
    lines and the stream scan -- with different inputs per thread; every result bit-exact"""
    import threading
    n = 8192
    prog2 = cu.CudaProgram(corpus.C2_REGEX)
    prog3 = cu.CudaProgram(corpus.C3_REGEX)
    progs = cu.CudaProgram(corpus.BENCH_REGEX)
    inputs, want = [], []
    for t in range(4):
        lines = corpus.log_lines(n, 1024, first_line=t * n)
        host = lines.numpy()
        _, w2, _ = baseline.run_lines("oracle", corpus.C2_REGEX, None, host, n, 1024, 1024, baseline.ENGINE_THOMPSON,
                                      nthreads=8)
        _, w3rc, w3ov = baseline.run_lines("oracle", corpus.C3_REGEX, None, host, n, 1024, 1024, baseline.ENGINE_PIKE,
                                           nthreads=8, ovec_slots=prog3.nslots)
        stream = corpus.gen_data_buffer(30000 + 777 * t)
        if t % 2:
            stream = stream[:-8].contiguous()
        inputs.append((lines.cuda(), stream.cuda(), stream.numel()))
        want.append((w2, w3rc, w3ov, capi.SRE_DECLINED if t % 2 else capi.SRE_OK))
    errors = []

    def worker(t):
        try:
            torch.cuda.set_device(0)
            st = torch.cuda.Stream()
            dev, sdev, slen = inputs[t]
            w2, w3rc, w3ov, wst = want[t]
            with torch.cuda.stream(st):
                for _ in range(25):
                    rc = prog2.thompson_lines(dev, n, 1024, 1024)
                    prc, pov = prog3.pike_lines(dev, n, 1024, 1024)
                    src, _, _ = progs.thompson_stream(sdev, slen, 4096, True)
                    st.synchronize()
                    assert (rc.cpu().numpy() == w2).all(), "thompson"
                    assert (prc.cpu().numpy() == w3rc).all() and (pov.cpu().numpy() == w3ov).all(), "pike"
                    assert src == wst, "stream"
        except Exception as e:      # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors


def test_text_grep_one_pass_vs_oracle(cu):
    """sre_cuda_thompson_exec_text: the verdict of every line of a '\\n'-delimited buffer in one pass
    (pieces through the TMA pipeline, line structure folded into the table) against the oracle run
    line by line -- ragged log lines, empty lines, lines longer than a piece, pieces with more lines
    than the staging holds, a last line with and without terminator, anchors at line ends"""
    rs = np.random.RandomState(123)
    oracle = capi.load("oracle")
    src = corpus.log_lines(6000, 1024).numpy()

    def build(kind, trailing_newline):
        parts = []
        for i in range(6000):
            if kind == "log":
                n = int(rs.randint(0, 400))
            elif kind == "tiny":
                n = int(rs.randint(0, 12))                  # > 64 lines per 4 KB piece
            else:
                n = int(rs.choice([0, 30, 200, 5000, 9000]))  # lines that span pieces
            row = src[i]
            body = bytes(np.resize(row[1024 - 60 - min(n, 900): 1024 - 60], n)) if n else b""
            parts.append(body.replace(b"\n", b" ") + b"\n")
        data = b"".join(parts)
        if not trailing_newline:
            data += b'GET /x/1 HTTP/1.1" 503 '
        return data

    regexes = [corpus.C2_REGEX, corpus.C3_REGEX, rb"5\d\d $", rb"^\w+[./-]", rb"^$", rb"x*"]
    for kind in ("log", "tiny", "long"):
        for trailing in (True, False):
            data = build(kind, trailing)
            host = np.frombuffer(data, dtype=np.uint8)
            want_off = np.concatenate([[0], np.flatnonzero(host == 10) + 1]).astype(np.int64)
            if not trailing:
                want_off = np.concatenate([want_off, [len(data)]])
            dev = torch.from_numpy(np.concatenate([host, np.zeros(16, np.uint8)])).cuda()
            for rx in regexes if kind == "log" else regexes[:3]:
                prog = cu.CudaProgram(rx)
                assert prog.info.dfa_states and prog.info.dfa_states <= 128
                rc, off = prog.thompson_text(dev, len(data))
                assert np.array_equal(off.cpu().numpy(), want_off), (kind, trailing, rx)
                # verdicts only (no offsets wanted): the count-only hot loop, same rows
                rc_only, none = prog.thompson_text(dev, len(data), want_offsets=False)
                assert none is None and torch.equal(rc_only, rc), (kind, trailing, rx)
                po = oracle.compile(rx, 0)
                n = len(want_off) - 1
                step = 1 if kind != "tiny" else 3
                idx = list(range(0, n, step))
                want = np.array([oracle.thompson(po, data[want_off[i]:want_off[i + 1]]) for i in idx])
                assert np.array_equal(rc.cpu().numpy()[idx], want), (kind, trailing, rx)
                po.close()
                # too small an output: the count is still reported, the prefix is right
                rc2, off2 = prog.thompson_text(dev, len(data), max_lines=100)
                assert torch.equal(rc2, rc[:100]) and torch.equal(off2, off[:101])
    # lines that end exactly at piece boundaries (4 KB) or one byte off, every line matching (more
    # matched lines per piece than the staging holds), partial matches hanging over piece boundaries,
    # very long lines spanning several pieces with a match before / after / across the boundaries
    hit = b'HTTP/1.1" 503 '
    cases = {
        "piece-aligned": b"".join((b"a" * (4095 - len(hit)) + hit + b"\n") if i % 3 else (b"b" * 4095 + b"\n")
                                  for i in range(40)),
        "off-by-one": b"".join(b"c" * int(n) + b"\n" for n in [4094, 4096, 0, 4095, 8191, 1, 4097, 12287, 5]),
        "all-match": (b"x" + hit + b"\n") * 3000,
        "straddle": b"".join(b"z" * int(4096 * (1 + i % 3) - 20 - k) + hit + b"q" * 3 + b"\n"
                             for i, k in enumerate(range(0, 30))),
        "long": b"".join(b"y" * 9000 + (hit if i % 2 else b"") + b"w" * 7000 + b"\n" for i in range(12))
                + b"tail without newline " + hit,
    }
    for name, data in cases.items():
        host = np.frombuffer(data, dtype=np.uint8)
        dev = torch.from_numpy(np.concatenate([host, np.zeros(16, np.uint8)])).cuda()
        lines = data.split(b"\n")
        if lines[-1] == b"":
            lines.pop()
        for rx in (corpus.C2_REGEX, rb"^[a-z]+HTTP", rb"5\d\d q*$", rb"[xz]+H"):
            prog = cu.CudaProgram(rx)
            po = oracle.compile(rx, 0)
            want = np.array([oracle.thompson(po, ln + b"\n") for ln in lines[:-1]]
                            + [oracle.thompson(po, lines[-1] + (b"\n" if data.endswith(b"\n") else b""))])
            po.close()
            # (the verdict path fits its piece size to the input; force a few, the default last)
            for piece in ("512", "2176", "3584", "4096", None):
                if piece is None:
                    os.environ.pop("SRE_CUDA_TEXT_PIECE", None)
                else:
                    os.environ["SRE_CUDA_TEXT_PIECE"] = piece
                rc_only, _ = prog.thompson_text(dev, len(data), want_offsets=False)
                assert np.array_equal(rc_only.cpu().numpy(), want), (name, rx, piece)
            rc, _ = prog.thompson_text(dev, len(data))
            assert np.array_equal(rc.cpu().numpy(), want), (name, rx)
    # empty buffer, and a buffer that is a single unterminated line
    prog = cu.CudaProgram(corpus.C2_REGEX)
    rc, off = prog.thompson_text(torch.zeros(16, dtype=torch.uint8, device="cuda"), 0)
    assert rc.numel() == 0 and off.cpu().tolist() == [0]
    one = b'zz HTTP/1.1" 503 '
    rc, off = prog.thompson_text(torch.frombuffer(bytearray(one) + bytearray(16), dtype=torch.uint8).cuda(), len(one))
    assert rc.cpu().tolist() == [0] and off.cpu().tolist() == [0, len(one)]
    # a larger automaton (64 patterns, 2107 states) takes the index + ragged route: same answers
    pats = corpus.multi_pattern_set(64)
    data = build("log", True)
    host = np.frombuffer(data, dtype=np.uint8)
    dev = torch.from_numpy(np.concatenate([host, np.zeros(16, np.uint8)])).cuda()
    progm = cu.CudaProgram(pats)
    rc, off = progm.thompson_text(dev, len(data))
    want = progm.thompson_ragged(dev, off)
    assert torch.equal(rc, want) and 0 < int((rc == 0).sum()) < rc.numel()


def test_nfa_tier_on_a_regex_that_defeats_determinisation(cu):
    """/[ab]*a[ab]{15}c/: 20-odd NFA states, 2^16 subsets -- no DFA, so AUTO takes the
    bit-parallel NFA (thread per line, 64-bit set); same verdicts as the warp-per-line kernel
    and as the oracle, on text over {a, b, c} where partial matches are alive all the time"""
    rx = rb"[ab]*a[ab]{15}c"
    prog = cu.CudaProgram(rx)
    assert prog.info.dfa_states == 0 and prog.info.nfa_states <= 64
    n, linelen = 3000, 256
    rs = np.random.RandomState(17)
    lines = np.frombuffer(b"abc", dtype=np.uint8)[rs.choice(3, size=(n, linelen), p=[0.48, 0.48, 0.04])].copy()
    lines[::3] = np.frombuffer(b"ab", dtype=np.uint8)[rs.randint(0, 2, size=(len(lines[::3]), linelen))]   # no 'c': no match
    _, want, _ = baseline.run_lines("oracle", rx, None, lines, n, linelen, linelen, baseline.ENGINE_THOMPSON, nthreads=8)
    dev = torch.from_numpy(lines).cuda()
    for engine in (cu.ENGINE_AUTO, cu.ENGINE_NFA, cu.ENGINE_NFA_WARP):
        got = prog.thompson_lines(dev, n, linelen, linelen, engine=engine).cpu().numpy()
        assert (got == want).all(), engine
    assert 0.2 < (want == 0).mean() < 0.8
    # log text: the same through AUTO on 1 KB lines
    log = corpus.log_lines(2048, 1024)
    _, want, _ = baseline.run_lines("oracle", rx, None, log.numpy(), 2048, 1024, 1024, baseline.ENGINE_THOMPSON,
                                    nthreads=8)
    got = prog.thompson_lines(log.cuda(), 2048, 1024, 1024).cpu().numpy()
    assert (got == want).all()
    # the packed-entry kernel in its other shapes: 33..64 lowered states (64-bit sets), several loop
    # states (non-shift movers), assertions (more than one kind of byte), more than four movers
    # (the looping form); ragged line lengths exercise the byte-wise last tile
    rs = np.random.RandomState(5)
    sigma = np.frombuffer(b"abcde x\n", dtype=np.uint8)
    n = 1500
    for rx in (rb"[ab]*a[ab]{40}c", rb"(a|b)*x[ab]{35}(c|d)+e", rb"\b[ab]+ [ab]{20}\b", rb"^[ab]*a[ab]{12}c$",
               rb"(a*b*c*d*e*x)+[ab]{30}", rb"[ab]*a[ab]{15}c|[cd]*c[cd]{15}a|x+e+x"):
        prog = cu.CudaProgram(rx)
        assert prog.info.nfa_states <= 64, (rx, prog.info.nfa_states)
        for linelen in (16, 100, 256):
            lines = sigma[rs.choice(len(sigma), size=(n, 256), p=[0.3, 0.3, 0.1, 0.08, 0.08, 0.06, 0.06, 0.02])].copy()
            _, want, _ = baseline.run_lines("oracle", rx, None, lines, n, 256, linelen, baseline.ENGINE_THOMPSON,
                                            nthreads=8)
            dev = torch.from_numpy(lines).cuda()
            for engine in (cu.ENGINE_NFA, cu.ENGINE_NFA_WARP):
                got = prog.thompson_lines(dev, n, 256, linelen, engine=engine).cpu().numpy()
                assert (got == want).all(), (rx, linelen, engine, int((got != want).sum()))


def test_pike_reproduces_the_reference_prefilter_misfire(cu):
    """The reference's first-byte prefilter is not result neutral: after a match reported in the
    step a prefilter jump landed on, survivors that look like the initial list are dropped and a
    LATER match overwrites the leftmost one (sre_vm_pike.c:262-274; tests/test_oracle.py pins the
    oracle to the live reference on QUIRK_CASES).  Every Pike tier and the classic
    sre_vm_pike_exec give the reference's answer: the fast tiers compute the leftmost-first match,
    k_pike_quirk_mark finds the lines on which the misfire is possible, and the general kernel
    replays those to the letter (pike_exec(faithful))."""
    import random
    from test_oracle import QUIRK_CASES
    o = capi.load("oracle")
    cuda_lib = capi.load("cuda")
    pitch = 32
    for rx, s, want, _leftmost in QUIRK_CASES:
        prog = cu.CudaProgram(rx)
        po = o.compile(rx, 0)
        assert o.pike(po, s) == want
        host = np.zeros((1, pitch), dtype=np.uint8)
        host[0, : len(s)] = np.frombuffer(s, dtype=np.uint8)
        dev = torch.from_numpy(host).cuda()
        for tier in (0, 1, 2, 3):
            prog.set_pike_tier(tier)
            rc, ov = prog.pike_lines(dev, 1, pitch, len(s))
            got = (int(rc[0]), ov[0].cpu().tolist() if int(rc[0]) >= 0 else None)
            assert got == want, (rx, s, tier, got)
        pc = cuda_lib.compile(rx, 0)
        assert cuda_lib.pike(pc, s) == want, (rx, s, "classic")
        pc.close()
        po.close()
        prog.program.close()
    # the family at large: regexes that can match one byte, lines full of isolated candidates
    rng = random.Random(2718)
    heads = [rb"a+", rb"\w+", rb"\d+", rb"[ab]+", rb"(a+)", rb"(\w)+", rb"(?:a|b)+", rb"^a+", rb"\ba+", rb"(?:^|\b)a+",
             rb"\B?a+", rb"^?[ab]+"]
    tails = [rb"b?", rb"x?", rb"\.?", rb"(\d)?", rb" ?", rb"(?:ab)?", rb"b*", rb"$", rb"", rb"(?:b|$)?", rb"b?\b", rb"\B?x?"]
    nlines, pitch = 512, 64
    alphabet = list(b"ab1. x\n")
    total = differ = 0
    for _ in range(40):
        rx = rng.choice(heads) + rng.choice(tails)
        linelen = rng.choice([5, 16, 33, 64])
        host = np.array([rng.choice(alphabet) for _ in range(nlines * pitch)], dtype=np.uint8).reshape(nlines, pitch)
        prog = cu.CudaProgram(rx)
        _, want_rc, want_ov = baseline.run_lines("oracle", rx, None, host, nlines, pitch, linelen,
                                                 baseline.ENGINE_PIKE, ovec_slots=prog.nslots)
        o.pike_prefilter(False)
        try:
            _, _, clean_ov = baseline.run_lines("oracle", rx, None, host, nlines, pitch, linelen,
                                                baseline.ENGINE_PIKE, ovec_slots=prog.nslots)
        finally:
            o.pike_prefilter(True)
        total += nlines
        differ += int((clean_ov != want_ov).any(axis=1).sum())
        dev = torch.from_numpy(host).cuda()
        for tier in (0, 1, 2, 3):
            prog.set_pike_tier(tier)
            rc, ov = prog.pike_lines(dev, nlines, pitch, linelen)
            assert (rc.cpu().numpy() == want_rc).all(), (rx, linelen, tier)
            assert (ov.cpu().numpy() == want_ov).all(), (rx, linelen, tier)
        prog.program.close()
    assert differ > 50, (differ, total)     # the misfire really is exercised


def test_batched_streaming_pike_contexts(cu):
    """sre_cuda_pike_streams_*: many persistent Pike contexts fed chunk by chunk at the same time
    (one ctx per connection, the ngx_replace_filter shape): per stream the whole trace of
    (rc, temp capture, pending match) and the final rc + ovector equal the oracle's
    sre_vm_pike_exec driven with the same chunks, continuation after a match included"""
    import random
    rng = random.Random(2024)
    o = capi.load("oracle")
    for rx in (rb"(a+)(b|c)", rb"(\w+) (\S+) HTTP/(\d)\.(\d)", [rb"ab+c", rb"(x|y)z", rb"\bGET\b"]):
        prog = cu.CudaProgram(rx)
        po = o.compile(rx)
        n = 200
        subjects = []
        for i in range(n):
            k = rng.randrange(0, 60)
            if isinstance(rx, list) or rx.startswith(b"(a+)"):
                body = bytes(rng.choice(b"abcxyz GET") for _ in range(k))
            else:
                body = bytes(corpus.log_lines(1, 1024, first_line=i).numpy()[0]).rstrip(b".")[-(20 + k):]
            subjects.append(body)
        # per stream: chunks of random sizes (some empty), eof on the last
        plans = []
        for sbj in subjects:
            cuts, at = [], 0
            while at < len(sbj):
                m = rng.choice([0, 1, 1, 2, 5, 13])
                cuts.append(sbj[at:at + m])
                at += m
            cuts.append(b"")
            plans.append([(c, j == len(cuts) - 1) for j, c in enumerate(cuts)])
        want = [o.pike(po, sbj, plan) for sbj, plan in zip(subjects, plans)]     # (trace, rc, ov)
        streams = cu.PikeStreams(prog, n)
        traces = [[] for _ in range(n)]
        final = [None] * n
        rounds = max(len(p) for p in plans)
        for r in range(rounds):
            parts, off, eofs = [], [0], []
            for i in range(n):
                if final[i] is None and r < len(plans[i]):
                    chunk, eof = plans[i][r]
                else:
                    chunk, eof = b"", False          # nothing more for this stream in this round
                parts.append(chunk)
                off.append(off[-1] + len(chunk))
                eofs.append(1 if eof else 0)
            blob = b"".join(parts) + bytes(16)
            out = streams.exec(torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda(),
                               torch.tensor(off, dtype=torch.int64, device="cuda"),
                               eof=torch.tensor(eofs, dtype=torch.uint8, device="cuda")).cpu().tolist()
            for i in range(n):
                if final[i] is not None or r >= len(plans[i]):
                    continue
                row = out[i]
                if row[0] == capi.SRE_AGAIN:
                    traces[i].append((capi.SRE_AGAIN, row[4], row[5], row[2] if row[1] else None,
                                      row[3] if row[1] else None))
                else:
                    final[i] = (row[0], row[4:4 + prog.nslots] if row[0] >= 0 else None)
        streams.close()
        for i in range(n):
            wt, wrc, wov = want[i]
            assert final[i] == (wrc, wov), (rx, subjects[i], final[i], wrc, wov)
            assert traces[i] == [tuple(t) for t in wt], (rx, subjects[i])
        assert sum(1 for f in final if f[0] >= 0) > 20


def test_pike_lookahead_assertions_on_long_lines(cu):
    """Programs with look-ahead assertions (`$ \\z \\b \\B`) over lines long enough for the start
    hint to land anywhere (the search then begins from the start list of the byte in front: nothing
    / newline / word byte / other): rc and the whole ovector against the oracle, through the
    default tier (the determinised Pike VM when SRE_PDFA_LOOKAHEAD allows it, else the
    closure-table kernel) and through the closure-table tier."""
    import random
    rng = random.Random(4242)
    words = [b"GET", b"POST", b"a", b"ab", b"x1", b"_", b"77", b"index.html", b"b", b"."]
    seps = [b" ", b" ", b" ", b"\n", b".", b"-", b"  ", b""]
    regexes = [rb"(\w+)$", rb"\b(GET)\b (\w+)", rb"(\d+)\.*$", rb"(\w+)\b (\S+)$", rb"(a|ab)\b(.)", rb"(\w+)\B(\w)$",
               rb"^(\w+)\b.*?(\w+)$", rb"(\S+)\z", rb"\B(\d)(\d)\b", rb"(x\d)?\b(\w+)\b\.$", rb"(?:(a)|b)+\b",
               rb"(\w*)\b(\W+)\b(\w*)$", rb"\b(\w+) \1?$", rb"(a+)$|(b+)\b", rb"()\b(\w)"]
    nlines, pitch = 384, 208
    for rx in regexes:
        try:
            prog = cu.CudaProgram(rx)
        except Exception:
            continue            # (a back-reference: not in the reference's syntax either)
        for linelen in (200, 61):
            host = np.zeros((nlines, pitch), dtype=np.uint8)
            for i in range(nlines):
                row = bytearray()
                # a head that cannot match most of the regexes, so that the hint moves into the line
                row += b"-" * rng.randrange(0, 120)
                while len(row) < pitch:
                    row += rng.choice(words) + rng.choice(seps)
                host[i] = np.frombuffer(bytes(row[:pitch]), dtype=np.uint8)
            _, want_rc, want_ov = baseline.run_lines("oracle", rx, None, host, nlines, pitch, linelen,
                                                     baseline.ENGINE_PIKE, ovec_slots=prog.nslots)
            dev = torch.from_numpy(host).cuda()
            for tier in (0, 3):
                prog.set_pike_tier(tier)
                rc, ov = prog.pike_lines(dev, nlines, pitch, linelen)
                assert (rc.cpu().numpy() == want_rc).all(), (rx, linelen, tier)
                assert (ov.cpu().numpy() == want_ov).all(), (rx, linelen, tier)
        prog.program.close()
