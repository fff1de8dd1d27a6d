"""CPU tier: the product's lowering pass (sregex_b200/csrc/lower), executed by
the test-only host executor oracle/lower_check.cpp, against the reference's
golden vectors -- before and independently of any kernel."""
import ctypes as C
import os

import pytest

from conftest import runnable
from sregex_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lc():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liblowercheck.so"))
    lib.lc_create.restype = C.c_void_p
    lib.lc_create.argtypes = [C.c_void_p, C.c_uint]
    lib.lc_destroy.argtypes = [C.c_void_p]
    lib.lc_reset.argtypes = [C.c_void_p]
    lib.lc_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
    lib.lc_nfa_exec.restype = C.c_long
    lib.lc_nfa_exec.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint]
    lib.lc_dfa_exec.restype = C.c_long
    lib.lc_dfa_exec.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint, C.c_int]
    lib.lc_hint.restype = C.c_long
    lib.lc_hint.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.lc_hint_cls.restype = C.c_long
    lib.lc_hint_cls.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    lib.lc_table_pike.restype = C.c_long
    lib.lc_table_pike.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_long, C.POINTER(C.c_int64)]
    return lib


def _table_pike(lc, prog, s, start=0):
    """the closure-table Pike (CPU model of k_pike_table) -> (rc, ovector) or None"""
    ov = (C.c_int64 * prog.nslots)()
    rc = lc.lc_table_pike(prog.prog, s, len(s), start, ov)
    if rc == -1000:
        return None
    return rc, (list(ov) if rc >= 0 else None)


def _stream(fn, chunks):
    out = []
    for chunk, eof in chunks:
        rc = fn(chunk, len(chunk), int(eof))
        out.append(rc)
        if rc != capi.SRE_AGAIN:
            break
    return out


def test_lowered_tables_against_golden(golden, oracle, lc):
    """NFA tables, DFA (class table and byte table): single buffer verdicts and
    the rc sequence under 1-byte chunks (exact SRE_OK timing)."""
    ndfa = 0
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        h = lc.lc_create(p.prog, 4096)
        assert h
        s = b["subject_b"]
        info = (C.c_uint * 6)()
        lc.lc_info(h, info)
        tag = (b["file"], b["name"])
        assert lc.lc_nfa_exec(h, s, len(s), 1) == b["thompson"], tag
        lc.lc_reset(h)
        assert _stream(lambda c, n, e: lc.lc_nfa_exec(h, c, n, e), capi.split_chunks(s)) == b["thompson_split"], tag
        if info[3]:
            ndfa += 1
            for t256 in (0, 1):
                lc.lc_reset(h)
                assert lc.lc_dfa_exec(h, s, len(s), 1, t256) == b["thompson"], tag
            lc.lc_reset(h)
            assert _stream(lambda c, n, e: lc.lc_dfa_exec(h, c, n, e, 1), capi.split_chunks(s)) == b["thompson_split"], tag
        lc.lc_destroy(h)
        p.close()
    assert ndfa >= 1900


def test_pike_start_hint_never_passes_the_match_start(golden, oracle, lc):
    """The restart hint (offset after the last byte that only the `.*?` thread
    consumed, frozen once the automaton is in its absorbing ACC state, whose row
    carries no flags) must be <= the start of the leftmost-first match the
    reference's Pike VM reports, for every matching block of the corpus."""
    checked = 0
    for b in runnable(golden):
        if b["pike"]["rc"] < 0:
            continue
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        h = lc.lc_create(p.prog, 4096)
        s = b["subject_b"]
        hint = lc.lc_hint(h, s, len(s))
        if hint >= 0:
            assert hint <= b["pike"]["ov"][0], (b["file"], b["name"], hint, b["pike"]["ov"])
            checked += 1
        hint2 = lc.lc_hint_cls(h, s, len(s))
        if hint2 >= 0:
            assert hint2 <= b["pike"]["ov"][0], (b["file"], b["name"], hint2, b["pike"]["ov"])
            assert hint < 0 or hint2 == hint
        lc.lc_destroy(h)
        p.close()
    assert checked > 1000


def test_random_regex_fuzz_lowering_vs_live_reference(oracle, ref, lc):
    """Random regexes (heavy on assertions and nested repetition) x random
    subjects: the lowered NFA / DFA verdicts and the oracle's Pike results
    against the reference itself, single buffer and split at a random point."""
    import random
    rng = random.Random(2026)
    atoms = ["a", "b", "ab", " ", "\\n", "_", ".", "^", "$", "\\b", "\\B", "\\A", "\\z", "|", "(", ")", "(?:", "*",
             "+", "?", "*?", "+?", "??", "{2}", "{0,2}", "{1,}", "[ab]", "[^a]", "\\w", "\\W", "\\s", "\\d", "1"]
    alphabet = b"ab \n_1."
    tried = 0
    for _ in range(1500):
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(1, 8))).encode()
        try:
            pr = ref.compile(rx, 0)
        except capi.SreSyntaxError:
            continue
        po = oracle.compile(rx, 0)
        h = lc.lc_create(po.prog, 4096)
        info = (C.c_uint * 6)()
        lc.lc_info(h, info)
        tried += 1
        for _ in range(6):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 12)))
            want = ref.thompson(pr, s)
            lc.lc_reset(h)
            assert lc.lc_nfa_exec(h, s, len(s), 1) == want, (rx, s)
            if info[3]:
                lc.lc_reset(h)
                assert lc.lc_dfa_exec(h, s, len(s), 1, 1) == want, (rx, s)
            prc, pov = ref.pike(pr, s)
            assert oracle.pike(po, s) == (prc, pov), (rx, s)
            assert oracle.thompson(po, s) == want, (rx, s)
            if prc >= 0:
                hint = lc.lc_hint_cls(h, s, len(s))
                assert hint <= pov[0], (rx, s, hint, pov)
            cut = rng.randrange(0, len(s) + 1)
            chunks = [(s[:cut], False), (s[cut:], True)]
            lc.lc_reset(h)
            assert _stream(lambda c, n, e: lc.lc_nfa_exec(h, c, n, e), chunks)[-1] == want, (rx, s, cut)
        lc.lc_destroy(h)
        po.close()
        pr.close()
    assert tried > 600


def test_closure_table_pike_golden(golden, oracle, lc):
    """the closure tables of sre_closure.cpp, run by the CPU model of
    k_pike_table, reproduce the reference's Pike rc (regex id) + ovector on
    every golden block, multi-regex sets included -- from offset 0 and from the
    DFA start hint"""
    checked = hinted = multi = 0
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        multi += bool(b["multi"])
        s = b["subject_b"]
        got = _table_pike(lc, p, s)
        if got is not None:
            assert got == (b["pike"]["rc"], b["pike"]["ov"]), (b["file"], b["name"])
            checked += 1
            h = lc.lc_create(p.prog, 4096)
            hint = lc.lc_hint_cls(h, s, len(s)) if h else -1
            if hint > 0:
                assert _table_pike(lc, p, s, hint) == (b["pike"]["rc"], b["pike"]["ov"]), (b["file"], b["name"])
                hinted += 1
            if h:
                lc.lc_destroy(h)
        p.close()
    assert checked > 1500 and hinted > 100 and multi > 5, (checked, hinted, multi)


def test_closure_table_pike_multi_regex_fuzz(ref, oracle, lc):
    """random regex SETS (2-5 members, assertions and captures) x random
    subjects: regex id and ovector slice of the closure-table Pike against the
    reference itself (sre_vm_pike.c:945-989 prepare_matched_captures)"""
    import random
    rng = random.Random(4242)
    atoms = ["a", "b", "ab", " ", "_", ".", "^", "$", "\\b", "\\B", "\\A", "\\z", "|", "(a)", "(b*)", "(?:ab)+",
             "*", "+", "?", "+?", "{1,2}", "[ab]", "[^a]", "\\w", "\\d", "1", "(\\w+)", "\\n"]
    alphabet = b"ab \n_1."
    tried = 0
    for _ in range(1500):
        n = rng.randrange(2, 6)
        rxs = ["".join(rng.choice(atoms) for _ in range(rng.randrange(1, 6))).encode() for _ in range(n)]
        try:
            pr = ref.compile(rxs, 0, multi=True)
        except capi.SreSyntaxError:
            continue
        po = oracle.compile(rxs, 0, multi=True)
        if _table_pike(lc, po, b"") is None:
            po.close()
            pr.close()
            continue
        tried += 1
        for _ in range(6):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 14)))
            assert _table_pike(lc, po, s) == ref.pike(pr, s), (rxs, s)
        po.close()
        pr.close()
    assert tried > 500


def test_closure_table_pike_fuzz_vs_live_reference(ref, oracle, lc):
    """random regexes rich in assertions, nested repetition and captures x
    random subjects: the closure-table Pike against the reference itself"""
    import random
    rng = random.Random(777)
    atoms = ["a", "b", "ab", " ", "\\n", "_", ".", "^", "$", "\\b", "\\B", "\\A", "\\z", "|", "(", ")", "(?:",
             "*", "+", "?", "*?", "+?", "??", "{2}", "{0,2}", "{1,}", "[ab]", "[^a]", "\\w", "\\W", "\\s", "\\d",
             "1", "(a|b)", "(a*)", "(\\w+)"]
    alphabet = b"ab \n_1."
    tried = 0
    for _ in range(2500):
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode()
        try:
            pr = ref.compile(rx, 0)
        except capi.SreSyntaxError:
            continue
        po = oracle.compile(rx, 0)
        if _table_pike(lc, po, b"") is None:
            po.close()
            pr.close()
            continue
        tried += 1
        for _ in range(8):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 14)))
            assert _table_pike(lc, po, s) == ref.pike(pr, s), (rx, s)
        po.close()
        pr.close()
    assert tried > 1000


def test_closure_table_pike_64_pattern_set(oracle, lc):
    """the C4 pattern set (64 regexes, bucketed start closure) over log lines:
    matched id + ovector of the closure-table Pike == the oracle's Pike"""
    from sregex_b200 import corpus
    po = oracle.compile(corpus.multi_pattern_set(64), 0, multi=True)
    lines = corpus.log_lines(300, 1024).numpy()
    hits = 0
    for i in range(300):
        s = bytes(lines[i])
        want = oracle.pike(po, s)
        assert _table_pike(lc, po, s) == want, i
        hits += want[0] >= 0
    assert hits > 50
    po.close()


WIDE_REGEX = rb'(\d+)\.(\d+)\.(\d+)\.(\d+) - - \[(\d+)/(\w+)/(\d+):(\d+):(\d+):(\d+) ([+-]\d+)\] "(\S*) (\w+) (\S+) HTTP'


def test_closure_table_pike_many_groups(oracle, lc):
    """14 capture groups (30 slots, beyond the 16-bit save masks of narrow
    programs): the closure-table Pike == the oracle's Pike on log lines"""
    from sregex_b200 import corpus
    po = oracle.compile(WIDE_REGEX, 0)
    assert po.ncaps == 14
    lines = corpus.log_lines(200, 1024).numpy()
    hits = 0
    for i in range(200):
        s = bytes(lines[i])
        want = oracle.pike(po, s)
        assert _table_pike(lc, po, s) == want, i
        hits += want[0] >= 0
    assert hits == 200
    po.close()


def _pdfa_pike(lc, prog, s, start=0, ring=4096):
    """the P-DFA Pike (CPU model of k_pike_lineage) -> (rc, ovector), None (no P-DFA) or "ring" """
    ov = (C.c_int64 * prog.nslots)()
    info = (C.c_uint * 2)()
    rc = lc.lc_pdfa_pike(prog.prog, s, len(s), start, ring, ov, info)
    if rc == -1000:
        return None
    if rc == -1001:
        return "ring"
    return rc, (list(ov) if rc >= 0 else None)


def _bind_pdfa(lc):
    lc.lc_pdfa_pike.restype = C.c_long
    lc.lc_pdfa_pike.argtypes = [C.c_void_p, C.c_char_p, C.c_long, C.c_long, C.c_long, C.POINTER(C.c_int64),
                                C.POINTER(C.c_uint)]


def test_pdfa_pike_against_golden(golden, oracle, lc, leftmost_first):
    """The determinised Pike VM (ordered thread lists as DFA states + lineage walk,
    lower/sre_pdfa.cpp) on every golden block it applies to (every block whose automaton stays small): rc and
    the whole ovector; with a ring of 8 positions the walk back refills it by running forward again
    (k_pike_lineage's answer to matches longer than its ring) and still agrees."""
    _bind_pdfa(lc)
    n = short = 0
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        s = b["subject_b"]
        got = _pdfa_pike(lc, p, s)
        if got is not None:
            n += 1
            assert got == (b["pike"]["rc"], b["pike"]["ov"]), (b["file"], b["name"], got, b["pike"])
            got8 = _pdfa_pike(lc, p, s, ring=8)
            assert got8 == got, (b["file"], b["name"])
            short += got[0] >= 0 and got[1][1] - got[1][0] > 8
        p.close()
    assert n > 1000 and short > 20, (n, short)


def test_pdfa_pike_fuzz_vs_oracle(oracle, lc, leftmost_first):
    """random regexes (nested groups, lazy and greedy repetition, alternation, empty
    loops, all six assertions over subjects with newlines) and sets of them x random
    subjects: rc + ovector == the oracle's Pike"""
    import random
    _bind_pdfa(lc)
    rng = random.Random(31337)
    atoms = ["a", "b", "ab", " ", "_", ".", "|", "(", ")", "(?:", "*", "+", "?", "*?", "+?", "??", "{2}", "{0,2}",
             "{1,}", "[ab]", "[^a]", "\\w", "\\W", "\\s", "\\d", "1", "(a)", "(b*)", "(a|ab)", "(\\w+)",
             "^", "\\A", "\\n", "(^a)", "(?:^|b)", "$", "\\z", "\\b", "\\B", "(a$)", "(\\bb)", "(?:$|a)",
             # optional look-ahead assertions: a held closure meets what the previous step tagged
             # without parking it (a MATCH already reported) -- the `odd` marks of a P-DFA state
             "\\B?", "\\b?", "$?", "(?:\\b|)", "(\\B)?", "x?", "a+"]
    alphabet = b"ab _1.\n\nx"
    done = applicable = 0
    while done < 2500:
        k = 1 if rng.random() < 0.7 else rng.randrange(2, 5)
        rxs = ["".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode() for _ in range(k)]
        try:
            p = oracle.compile(rxs if k > 1 else rxs[0], 0)
        except capi.SreSyntaxError:
            continue
        done += 1
        for _ in range(3):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
            got = _pdfa_pike(lc, p, s)
            if got is None:
                break
            applicable += 1
            assert got == oracle.pike(p, s), (rxs, s, got)
        p.close()
    assert applicable > 2000


def test_pdfa_held_closures_see_the_previous_steps_tags(oracle, lc, leftmost_first):
    """A look-ahead thread that holds appends its closure under the PREVIOUS step's tag
    (sre_vm_pike.c:484-509), and that step tagged more than it parked: in /a\\B?x?/ on "ax" the
    step on `a` reports a match (MATCH tagged, not parked) and parks [\\B, x]; on `x` the \\B
    thread holds, its closure finds `x` and MATCH tagged already and adds nothing, and the `x`
    thread goes on to the longer match.  A P-DFA state that forgot the MATCH tag reported (0, 1)."""
    _bind_pdfa(lc)
    for rx, s in [(rb"a\B?x?", b"ax"), (rb"a(?:\B|)x?", b"ax"), (rb"a+\B?x?", b".aax."), (rb"a\B?(x)?", b"ax"),
                  (rb"a\B?x?", b"axx"), (rb"^a+\B?x?", b"\nax"), (rb"a\b?x?", b"a x"), (rb"a$?x?", b"ax")]:
        p = oracle.compile(rx, 0)
        got = _pdfa_pike(lc, p, s)
        assert got is not None and got == leftmost_first.pike(p, s) == _table_pike(lc, p, s), (rx, s, got)
        p.close()


def test_dfa_minimisation_counts_its_initial_blocks(oracle, lc, leftmost_first):
    """Moore refinement stops when a round adds no block.  When every state has the start state's
    EOF verdict (a set with a member that always matches at the end, /$?$\\z/) there are two
    initial blocks, not three; counted as three, a first round that ended with three blocks was
    taken for stable and distinguishable states stayed merged: the DFA of this set never entered
    ACC (late SRE_OK when streaming) and its restart flags let the Pike search begin behind the
    match."""
    rxs = [rb"(\B)?[^a] ", rb"$?$\z"]
    s = b"x1 ax\n\n\n\nx\n .1x\nb_ "
    p = oracle.compile(rxs, 0)
    h = lc.lc_create(p.prog, 4096)
    for n in range(len(s) + 1):
        chunk = s[:n]
        want = oracle.thompson(p, chunk, [(chunk, False)])[0]
        for mode in (0, 1):
            lc.lc_reset(h)
            assert lc.lc_dfa_exec(h, chunk, n, 0, mode) == want, (n, mode)
        lc.lc_reset(h)
        assert lc.lc_nfa_exec(h, chunk, n, 0) == want, n
    want = leftmost_first.pike(p, s)
    assert want == (0, [1, 3, -1, -1])
    hint = lc.lc_hint_cls(h, s, len(s))
    assert 0 <= hint <= 1, hint
    assert lc.lc_hint(h, s, len(s)) == hint
    _bind_pdfa(lc)
    assert _pdfa_pike(lc, p, s, hint) == want and _table_pike(lc, p, s, hint) == want
    lc.lc_destroy(h)
    p.close()


SKIPPABLE_LOOKAHEAD = [      # (regex, subject): the closure-table model or the P-DFA model differ from the reference
    (rb"x?(?:\B|x)+", b"ax A\n A1B\n1BxB"),
    (rb"(\w+)*?(\B)?(a)", b"1BBA1x \n.A\n Bbb Ba.x"),
    (rb"(\B)?(a)?(\w+)(?:\B|x)+", b" A1aAx_ A\n \nB\n\n\n\na_B1"),
    (rb"(a)??(\w+)(\B)?(a|ab)", b"xax \nbbA._1B.\n_a.Ba1"),
    (rb"\b?(\w+)+(?:\b|)*(?:\B|x)+", b"\nabx .B.\na\n\nxbb"),
    (rb"\s(?:\B|x)+B*?(?:\b|)", b"_\nxb\n.bBa_x"),
]


def test_lookahead_overlap_flags_every_program_the_tables_get_wrong(oracle, lc, leftmost_first):
    """The closure tables and the P-DFA deduplicate closures by parked instruction, the reference
    by one tag word per instruction, and a held look-ahead closure runs under the PREVIOUS step's
    tag (sre_vm_pike.c:484-509).  The two can differ when one closure parks a look-ahead assertion
    and also visits what lies behind it -- an assertion that can be skipped, /(\\B)?x/ -- which
    oracle/lower_check.cpp (sre_lookahead_overlap) reports as level 2.  This is the one program
    shape on which the batch Pike tiers are known to differ from the reference (DESIGN.md 3.3,
    INTEGRATION.md): wherever a CPU model of the fast tiers differs from the oracle the program
    is level 2 (63 of 30,000 fuzzed regex sets when this was written), and ordinary patterns --
    an assertion next to a literal, at the end of an alternative, around a group -- are not."""
    import random
    from sregex_b200 import corpus
    _bind_pdfa(lc)
    lc.lc_lookahead_overlap.argtypes = [C.c_void_p]
    for rx in (corpus.C2_REGEX, corpus.C3_REGEX, corpus.BENCH_REGEX, rb"\bGET\b", rb"(\w+)$", rb"^(\d+)\.(\d+)",
               rb"foo\b|bar", rb"^$|^#", rb"(GET|POST)\b", rb'"\s(\d+)\b.*?(\.*)$', rb"(?:^|\s)foo\b", rb"\d+$"):
        p = oracle.compile(rx, 0)
        assert lc.lc_lookahead_overlap(p.prog) < 2, rx
        p.close()
    p = oracle.compile(corpus.multi_pattern_set(64), 0)
    assert lc.lc_lookahead_overlap(p.prog) == 0
    p.close()
    wrong = 0
    for rx, s in SKIPPABLE_LOOKAHEAD:
        p = oracle.compile(rx, 0)
        assert lc.lc_lookahead_overlap(p.prog) == 2, rx
        want = leftmost_first.pike(p, s)
        wrong += _table_pike(lc, p, s) != want or _pdfa_pike(lc, p, s) != want
        p.close()
    assert wrong >= 4, wrong
    rng = random.Random(2024)
    atoms = ["a", "b", "A", "ab", " ", "_", ".", "|", "(", ")", "(?:", "*", "+", "?", "*?", "+?", "{2}", "{0,2}", "[ab]",
             "[^a]", "\\w", "\\W", "\\d", "\\s", "1", "(a)", "(b*)", "(a|ab)", "(\\w+)", "()", "^", "\\A", "\\n", "$",
             "\\z", "\\b", "\\B", "\\b?", "\\B?", "$?", "^?", "(?:\\b|)", "(?:$|a)", "(\\B)?", "(?:\\B|x)+", "x?", "a+"]
    alphabet = b"abAB _1.\n\nx"
    done = differ = 0
    levels = [0, 0, 0]
    while done < 2000:
        k = 1 if rng.random() < 0.8 else rng.randrange(2, 4)
        rxs = ["".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode() for _ in range(k)]
        try:
            p = oracle.compile(rxs if k > 1 else rxs[0], rng.random() < 0.25)
        except capi.SreSyntaxError:
            continue
        done += 1
        level = lc.lc_lookahead_overlap(p.prog)
        levels[level] += 1
        for _ in range(4):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
            want = leftmost_first.pike(p, s)
            for got in (_table_pike(lc, p, s), _pdfa_pike(lc, p, s, ring=16)):
                if got is not None and got != "ring" and got != want:
                    differ += 1
                    assert level == 2, (rxs, s, got, want)
        p.close()
    assert differ >= 3 and levels[0] > 500, (differ, levels)


def test_lowering_models_broad_fuzz(oracle, lc, leftmost_first):
    """One fuzz over everything the lowering produces, on regex SETS (members that can match the
    empty string at the end included), with and without SRE_REGEX_CASELESS, over subjects with
    newlines: NFA and DFA verdicts (class table and byte table) == the oracle's Thompson; the two
    hint tables agree and never pass the match start; the closure-table Pike and the P-DFA Pike,
    from offset 0, from the hint and from the 16-byte boundary below it, == the oracle's Pike --
    except on programs with a look-ahead assertion that can be skipped (level 2 of
    sre_lookahead_overlap, the known divergence of DESIGN.md 3.3), where the Thompson side and the
    hint are still checked.  This is the fuzz that found the under-refined DFA minimisation and the
    missing tag state of the P-DFA."""
    import random
    _bind_pdfa(lc)
    lc.lc_lookahead_overlap.argtypes = [C.c_void_p]
    rng = random.Random(1234)
    atoms = ["a", "b", "A", "B", "ab", "Ab", " ", "_", ".", "|", "|", "(", ")", "(?:", "*", "+", "?", "*?", "+?", "??",
             "{2}", "{0,2}", "{2,}", "{1,3}?", "[ab]", "[^a]", "[A-Ca]", "[^\\n]", "\\w", "\\W", "\\d", "\\D", "\\s",
             "\\S", "1", "(a)", "(b*)", "(a|ab)", "(\\w+)", "()", "^", "\\A", "\\n", "$", "\\z", "\\b", "\\B", "\\b?",
             "$?", "^?", "(?:$|a)", "(^)*", "x?", "a+", "\\x41", "\\.", "[\\d_]"]
    alphabet = b"abAB _1.\n\nx"
    done = pike_checked = hinted = 0
    while done < 2500:
        k = 1 if rng.random() < 0.7 else rng.randrange(2, 5)
        rxs = ["".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode() for _ in range(k)]
        flags = capi.SRE_REGEX_CASELESS if rng.random() < 0.3 else 0
        try:
            p = oracle.compile(rxs if k > 1 else rxs[0], flags)
        except capi.SreSyntaxError:
            continue
        done += 1
        exotic = lc.lc_lookahead_overlap(p.prog) == 2
        h = lc.lc_create(p.prog, 4096)
        info = (C.c_uint * 6)()
        lc.lc_info(h, info)
        for _ in range(3):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
            tag = (rxs, flags, s)
            want_t = oracle.thompson(p, s)
            lc.lc_reset(h)
            assert lc.lc_nfa_exec(h, s, len(s), 1) == want_t, tag
            hint = 0
            if info[3]:
                for byte_table in (0, 1):
                    lc.lc_reset(h)
                    assert lc.lc_dfa_exec(h, s, len(s), 1, byte_table) == want_t, tag
                hint = lc.lc_hint_cls(h, s, len(s))
                if info[3] <= 128:
                    assert lc.lc_hint(h, s, len(s)) == hint, tag
            want = leftmost_first.pike(p, s)
            if want[0] >= 0:
                assert 0 <= hint <= want[1][0], (tag, hint, want)
            if exotic:
                continue
            hinted += hint > 0
            for start in {0, hint, hint & ~15}:
                got = _table_pike(lc, p, s, start)
                assert got is None or got == want, (tag, start, got, want)
                got = _pdfa_pike(lc, p, s, start, ring=16)
                assert got is None or got == want, (tag, start, got, want)
                pike_checked += got is not None
        lc.lc_destroy(h)
        p.close()
    assert pike_checked > 8000 and hinted > 400, (pike_checked, hinted)


def test_pdfa_built_from_the_bytecode_is_exact_on_skippable_lookahead(golden, oracle, lc, leftmost_first, monkeypatch):
    """SRE_PDFA_EXACT=1: sre_build_pdfa steps the bytecode itself under the reference's tag words
    (the tagged instructions are part of a state) instead of combining the closure tables.  On the
    CPU model of k_pike_lineage that P-DFA gives the oracle's rc and ovector on the programs the
    table-built one gets wrong (SKIPPABLE_LOOKAHEAD), on random regex sets full of such atoms from
    offset 0 / the hint / the 16-byte boundary below it, and on a sample of the golden blocks.  Off
    by default: the tables it emits have not been run on a GPU."""
    import random
    _bind_pdfa(lc)
    lc.lc_lookahead_overlap.argtypes = [C.c_void_p]
    monkeypatch.setenv("SRE_PDFA_EXACT", "1")
    for rx, s in SKIPPABLE_LOOKAHEAD:
        p = oracle.compile(rx, 0)
        assert _pdfa_pike(lc, p, s) == leftmost_first.pike(p, s), rx
        p.close()
    rng = random.Random(77)
    atoms = ["a", "b", "A", "ab", " ", "_", ".", "|", "(", ")", "(?:", "*", "+", "?", "*?", "+?", "{2}", "{0,2}", "[ab]",
             "[^a]", "\\w", "\\W", "\\d", "\\s", "1", "(a)", "(b*)", "(a|ab)", "(\\w+)", "()", "^", "\\A", "\\n", "$",
             "\\z", "\\b", "\\B", "\\b?", "\\B?", "$?", "^?", "(?:\\b|)", "(?:$|a)", "(\\B)?", "(?:\\B|x)+", "x?", "a+"]
    alphabet = b"abAB _1.\n\nx"
    done = skippable = checked = 0
    while done < 1500:
        k = 1 if rng.random() < 0.8 else rng.randrange(2, 4)
        rxs = ["".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode() for _ in range(k)]
        try:
            p = oracle.compile(rxs if k > 1 else rxs[0], rng.random() < 0.25)
        except capi.SreSyntaxError:
            continue
        done += 1
        skippable += lc.lc_lookahead_overlap(p.prog) == 2
        h = lc.lc_create(p.prog, 4096)
        for _ in range(3):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
            want = leftmost_first.pike(p, s)
            hint = max(0, lc.lc_hint_cls(h, s, len(s))) if h else 0
            for start in {0, hint, hint & ~15}:
                got = _pdfa_pike(lc, p, s, start, ring=16)
                assert got is None or got == want, (rxs, s, start, got, want)
                checked += got is not None
        if h:
            lc.lc_destroy(h)
        p.close()
    assert skippable > 300 and checked > 5000, (skippable, checked)
    for b in runnable(golden)[::5]:
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        got = _pdfa_pike(lc, p, b["subject_b"])
        assert got is None or got == (b["pike"]["rc"], b["pike"]["ov"]), (b["file"], b["name"])
        p.close()


def test_pdfa_pike_from_the_start_hint(golden, oracle, lc, leftmost_first):
    """k_pike_lineage begins at the 16-byte boundary below the DFA start hint, from the start list
    of the byte in front (nothing / newline / word byte / other -- what `^ \\A` look back at and
    what decides `\\b \\B` at the first byte).  Model: every start from the hint down to 16 bytes
    below it gives the oracle's rc and ovector, on the golden blocks and on random regexes with all
    six assertions over longer subjects."""
    import random
    _bind_pdfa(lc)

    def check(p, s, want, tag):
        if _pdfa_pike(lc, p, s) is None:
            return 0
        h = lc.lc_create(p.prog, 4096)
        hint = lc.lc_hint_cls(h, s, len(s)) if h else -1
        if h:
            lc.lc_destroy(h)
        if hint <= 0:
            return 0
        for start in range(max(0, hint - 16), hint + 1):
            assert _pdfa_pike(lc, p, s, start) == want, (tag, s, hint, start)
        return 1

    hinted = 0
    for b in runnable(golden)[::4]:
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        hinted += check(p, b["subject_b"], (b["pike"]["rc"], b["pike"]["ov"]), (b["file"], b["name"]))
        p.close()
    assert hinted > 20, hinted

    rng = random.Random(777)
    atoms = ["a", "b", "ab", " ", "_", ".", "|", "(", ")", "(?:", "*", "+", "?", "+?", "{2}", "[ab]", "[^a]", "\\w",
             "\\W", "\\d", "1", "(a)", "(a|ab)", "(\\w+)", "^", "\\A", "\\n", "$", "\\z", "\\b", "\\B", "(a$)",
             "(\\bb)", "(?:$|a)", "\\b", "$", "\\B"]
    alphabet = b"ab _1.\n--"
    done = fuzz_hinted = 0
    while done < 400:
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(2, 8))).encode()
        try:
            p = oracle.compile(rx, 0)
        except capi.SreSyntaxError:
            continue
        done += 1
        for _ in range(2):
            s = b"-" * rng.randrange(0, 30) + bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
            fuzz_hinted += check(p, s, leftmost_first.pike(p, s), rx)
        p.close()
    assert fuzz_hinted > 100, fuzz_hinted


def test_prefilter_misfire_marking_rule(oracle, lc):
    """k_pike_quirk_mark's rule is a superset of the lines on which the reference's first-byte
    prefilter misfires (lower/sre_quirk.cpp): wherever the oracle with the prefilter (== the
    reference) and without it (the leftmost-first match the fast tiers compute) disagree, the
    leftmost-first match starts at offset s >= 1 on a byte of the program's `single` set with a
    non-leading byte on either side.  Programs that cannot match one byte have an empty set."""
    import random
    from sregex_b200 import corpus
    lc.lc_quirk_bytes.restype = C.c_int
    lc.lc_quirk_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]

    def sets(p):
        single, lead = (C.c_uint32 * 8)(), (C.c_uint32 * 8)()
        possible = lc.lc_quirk_bytes(p.prog, single, lead)
        bit = lambda st, b: (st[b >> 5] >> (b & 31)) & 1
        return possible, (lambda b: bit(single, b)), (lambda b: bit(lead, b))

    for rx in (corpus.C2_REGEX, corpus.C3_REGEX, corpus.BENCH_REGEX):
        p = oracle.compile(rx, 0)
        assert sets(p)[0] == 0, rx
        p.close()
    p = oracle.compile(corpus.multi_pattern_set(64), 0)
    assert sets(p)[0] == 0
    p.close()

    rng = random.Random(1618)
    heads = [rb"a+", rb"\w+", rb"\d+", rb"[ab]+", rb"(a+)", rb"(\w)+", rb"(?:a|b)+", rb"^a+", rb"\ba+", rb"a", rb"(a|ab)+"]
    tails = [rb"b?", rb"x?", rb"\.?", rb"(\d)?", rb" ?", rb"(?:ab)?", rb"b*", rb"$", rb"", rb"\b", rb"(b|c)?"]
    alphabet = b"ab1. x\n"
    differ = 0
    for _ in range(300):
        rx = rng.choice(heads) + rng.choice(tails)
        p = oracle.compile(rx, 0)
        possible, single, lead = sets(p)
        for _ in range(40):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 24)))
            on = oracle.pike(p, s)
            oracle.pike_prefilter(False)
            try:
                off = oracle.pike(p, s)
            finally:
                oracle.pike_prefilter(True)
            if on != off:
                differ += 1
                assert possible and off[0] >= 0, (rx, s, on, off)
                st = off[1][0]
                assert st >= 1 and st + 1 < len(s) and single(s[st]) and not lead(s[st - 1]) \
                    and not lead(s[st + 1]), (rx, s, on, off)
        p.close()
    assert differ > 100, differ

    # programs with assertions: the analysis replays the landing step in each look-behind context
    # (the byte in front a newline / a word byte / neither) with the look-ahead threads parked as
    # the reference parks them -- random regexes over all six assertions
    pre = [rb"", rb"", rb"^", rb"\b", rb"\B", rb"(?:^|\b)", rb"\A|", rb"(?:\b|\B)", rb"^?"]
    mid = [rb"", rb"", rb"\b", rb"\B", rb"$?", rb"(?:\b|x)"]
    post = [rb"", rb"", rb"\b", rb"\B", rb"$", rb"\z", rb"(?:$|\b)", rb"|\bb"]
    alphabet = b"ab1. x\n_"
    tried = differ2 = flagged = 0
    while tried < 1200:
        rx = rng.choice(pre) + rng.choice(heads) + rng.choice(mid) + rng.choice(tails) + rng.choice(post)
        try:
            p = oracle.compile(rx, 0)
        except capi.SreSyntaxError:
            continue
        tried += 1
        possible, single, lead = sets(p)
        flagged += possible
        for _ in range(30):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 28)))
            on = oracle.pike(p, s)
            oracle.pike_prefilter(False)
            try:
                off = oracle.pike(p, s)
            finally:
                oracle.pike_prefilter(True)
            if on != off:
                differ2 += 1
                assert possible and off[0] >= 0, (rx, s, on, off)
                st = off[1][0]
                assert st >= 1 and st + 1 < len(s) and single(s[st]) and not lead(s[st - 1]) \
                    and not lead(s[st + 1]), (rx, s, on, off)
        p.close()
    assert differ2 > 100 and flagged < tried, (differ2, flagged, tried)
    # the shapes real patterns have: a word boundary or an anchor around a literal does not make
    # every matched line a candidate
    for rx in (rb'\b(GET|HEAD|POST|PUT) (\S+) HTTP/(\d)\.(\d)\b', rb'(\w+)$', rb'^(\d+)\.(\d+)', rb'\bERROR\b', rb'\d+$'):
        p = oracle.compile(rx, 0)
        assert sets(p)[0] == 0, rx
        p.close()
