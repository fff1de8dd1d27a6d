"""The CPU oracle (oracle/sre_oracle.c + the host front end) against the
reference's golden vectors: every runnable block of the reference's t/ suite,
answers recorded from the unmodified reference by tests/golden/gen_golden.py."""
import random

from conftest import runnable
from sregex_b200 import capi


def test_golden_has_whole_corpus(golden):
    st = golden["stats"]
    assert st["blocks"] == 1999 and st["runnable"] >= 1918 and st["explicit_checked"] >= 70
    assert all(b.get("modes_agree", True) for b in golden["blocks"])


def test_thompson_against_golden(golden, oracle):
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        assert oracle.thompson(p, b["subject_b"]) == b["thompson"], (b["file"], b["name"])
        p.close()


def test_thompson_streaming_against_golden(golden, oracle):
    """1-byte chunks: the whole rc sequence, incl. the interpreter's
    one-step-late SRE_OK (sre_vm_thompson.c:233-235)."""
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        got = oracle.thompson(p, b["subject_b"], capi.split_chunks(b["subject_b"]))
        assert got == b["thompson_split"], (b["file"], b["name"])
        p.close()


def test_pike_against_golden(golden, oracle):
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        rc, ov = oracle.pike(p, b["subject_b"])
        assert (rc, ov) == (b["pike"]["rc"], b["pike"]["ov"]), (b["file"], b["name"])
        p.close()


def test_pike_streaming_against_golden(golden, oracle):
    """1-byte chunks: final rc/ovector plus the temp-capture / pending-match
    trace after every call (sre_vm_pike.c:640-658, 692-735)."""
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        trace, rc, ov = oracle.pike(p, b["subject_b"], capi.split_chunks(b["subject_b"]))
        want = b["pike_split"]
        assert (rc, ov) == (want["rc"], want["ov"]), (b["file"], b["name"])
        assert [list(t) for t in trace] == want["trace"], (b["file"], b["name"])
        p.close()


def _random_subject(rng, alphabet, n):
    return bytes(rng.choice(alphabet) for _ in range(n))


def test_oracle_against_live_reference_fuzz(golden, oracle, ref):
    """Seeded random subjects over the corpus regexes, oracle vs the reference
    itself (only where oracle/_ref exists, i.e. not required on the GPU box)."""
    rng = random.Random(0x5EED)
    blocks = runnable(golden)
    for b in rng.sample(blocks, 400):
        po = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        pr = ref.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        alphabet = list(set(b["subject_b"]) | set(b"ab \n_")) or list(b"ab")
        for _ in range(6):
            s = _random_subject(rng, alphabet, rng.randrange(0, 40))
            assert oracle.thompson(po, s) == ref.thompson(pr, s), (b["name"], s)
            assert oracle.pike(po, s) == ref.pike(pr, s), (b["name"], s)
            cuts = sorted(rng.sample(range(len(s) + 1), min(len(s) + 1, 2)))
            chunks = [(s[:cuts[0]], False), (s[cuts[0]:cuts[-1]], False), (s[cuts[-1]:], True)]
            assert oracle.thompson(po, s, chunks) == ref.thompson(pr, s, chunks), (b["name"], s)
            assert oracle.pike(po, s, chunks) == ref.pike(pr, s, chunks), (b["name"], s, chunks)
        po.close()
        pr.close()


def test_oracle_against_live_reference_on_random_regex_sets(oracle, ref):
    """Random regexes and regex SETS -- assertions that can be skipped or looped ((\\B)?, (?:\\B|x)+,
    $?), members that match the empty string, SRE_REGEX_CASELESS -- oracle vs the reference itself:
    Pike rc + ovector (prefilter and all), Thompson rc, and both fed in 3-byte chunks.  The family
    on which the fast Pike tiers were found to differ from the reference: the
    oracle does not."""
    from test_lowering import SKIPPABLE_LOOKAHEAD
    for rx, s in SKIPPABLE_LOOKAHEAD:
        po, pr = oracle.compile(rx, 0), ref.compile(rx, 0)
        assert oracle.pike(po, s) == ref.pike(pr, s), rx
        po.close()
        pr.close()
    rng = random.Random(31)
    atoms = ["a", "b", "A", "ab", " ", "_", ".", "|", "(", ")", "(?:", "*", "+", "?", "*?", "+?", "{2}", "{0,2}", "[ab]",
             "[^a]", "\\w", "\\W", "\\d", "\\s", "1", "(a)", "(b*)", "(a|ab)", "(\\w+)", "()", "^", "\\A", "\\n", "$",
             "\\z", "\\b", "\\B", "\\b?", "\\B?", "$?", "^?", "(?:\\b|)", "(?:$|a)", "(\\B)?", "(?:\\B|x)+", "x?", "a+"]
    alphabet = b"abAB _1.\n\nx"
    done = 0
    while done < 1500:
        k = 1 if rng.random() < 0.8 else rng.randrange(2, 4)
        rxs = ["".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode() for _ in range(k)]
        flags = capi.SRE_REGEX_CASELESS if rng.random() < 0.25 else 0
        try:
            pr = ref.compile(rxs if k > 1 else rxs[0], flags)
        except capi.SreSyntaxError:
            continue
        po = oracle.compile(rxs if k > 1 else rxs[0], flags)
        done += 1
        for _ in range(4):
            s = bytes(rng.choice(alphabet) for _ in range(rng.randrange(0, 40)))
            assert oracle.pike(po, s) == ref.pike(pr, s), (rxs, flags, s)
            assert oracle.thompson(po, s) == ref.thompson(pr, s), (rxs, flags, s)
            chunks = [(s[i:i + 3], i + 3 >= len(s)) for i in range(0, len(s), 3)] or [(b"", True)]
            assert oracle.thompson(po, s, chunks) == ref.thompson(pr, s, chunks), (rxs, flags, s)
            assert oracle.pike(po, s, chunks) == ref.pike(pr, s, chunks), (rxs, flags, s)
        po.close()
        pr.close()


def test_oracle_against_reference_answers_on_fuzzed_regex_sets(oracle):
    """tests/golden/fuzz_sets.json.gz (tests/golden/gen_fuzz_sets.py): the unmodified reference's
    answers on 4000 seeded (regex set, subject) cases outside its own t/ corpus -- assertions that
    can be skipped or looped, members that match the empty string, SRE_REGEX_CASELESS, the
    prefilter-misfire family -- Thompson rc, Pike rc + ovector, and both fed in 3-byte chunks
    (Pike with its temp-capture / pending-match trace).  Runs wherever the fixture is, i.e. also on
    the GPU box where the reference is not."""
    import gzip
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fuzz_sets.json.gz")
    with gzip.open(path) as f:
        cases = json.load(f)["cases"]
    assert len(cases) == 4000
    plain = lambda x: json.loads(json.dumps(x))
    multi = matched = 0
    for c in cases:
        rxs = [bytes.fromhex(r) for r in c["regexes"]]
        s = bytes.fromhex(c["subject"])
        p = oracle.compile(rxs if len(rxs) > 1 else rxs[0], c["flags"])
        chunks = [(s[i:i + 3], i + 3 >= len(s)) for i in range(0, len(s), 3)] or [(b"", True)]
        assert oracle.thompson(p, s) == c["thompson"], (rxs, s)
        assert plain(oracle.pike(p, s)) == c["pike"], (rxs, s)
        assert plain(oracle.thompson(p, s, chunks)) == c["thompson_chunked"], (rxs, s)
        assert plain(oracle.pike(p, s, chunks)) == c["pike_chunked"], (rxs, s)
        multi += len(rxs) > 1
        matched += c["pike"][0] >= 0
        p.close()
    assert multi > 200 and 1000 < matched < 3900, (multi, matched)


def test_post_match_continuation_against_live_reference(golden, oracle, ref):
    """global scan through the classic API: after each match the ctx is given the
    rest of the data (sre_vm_pike.c:624-635, empty-match skip :179-193)"""
    n = 0
    for b in runnable(golden)[::2]:
        po = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        pr = ref.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        s = b["subject_b"] * 2
        assert capi.pike_all(oracle, po, s) == capi.pike_all(ref, pr, s), (b["file"], b["name"])
        n += 1
        po.close()
        pr.close()
    assert n > 900


QUIRK_CASES = [
    # regex, subject, the reference's answer, the leftmost-first answer
    (rb"a+b?", b".a.a", (0, [3, 4]), (0, [1, 2])),
    (rb"a+b?", b" ab cd", (0, [1, 2]), (0, [1, 3])),
    (rb"\w+x?", b".b.a", (0, [3, 4]), (0, [1, 2])),
    (rb"\d+\.?", b"-1-2", (0, [3, 4]), (0, [1, 2])),
    (rb"\d+\.?", b"x.b.1.b a", (0, [4, 5]), (0, [4, 6])),
    (rb"(\w+) ?", b".b.a", (0, [3, 4, 3, 4]), (0, [1, 2, 1, 2])),
    (rb"(\w+)+(\w+)?", b".b.1.b a", (0, [7, 8, 7, 8, -1, -1]), (0, [1, 2, 1, 2, -1, -1])),
    # ... and where it does not misfire both agree
    (rb"a+b?", b"b.a", (0, [2, 3]), (0, [2, 3])),
    (rb"\w+x?", b" ab cd", (0, [1, 3]), (0, [1, 3])),
]


def test_first_byte_prefilter_misfire_is_the_references(golden, oracle, ref):
    """The reference's first-byte prefilter is not result neutral (sre_vm_pike.c:262-274 compares
    the thread count and every pc but the last: after a match has cut the ".*?" thread, survivors
    can pass for the initial list, get dropped, and a LATER match overwrites the leftmost one).
    The oracle restates that bit for bit (== the live reference on these cases); with the
    prefilter left out it gives the leftmost-first match, which is what the CUDA tiers implement
    (tests/test_gpu_parity.py).  On the whole golden corpus the two coincide: the reference's own
    test suite never meets the misfire."""
    for rx, s, want_ref, want_clean in QUIRK_CASES:
        po, pr = oracle.compile(rx, 0), ref.compile(rx, 0)
        assert ref.pike(pr, s) == want_ref, (rx, s)
        assert oracle.pike(po, s) == want_ref, (rx, s)
        oracle.pike_prefilter(False)
        try:
            assert oracle.pike(po, s) == want_clean, (rx, s)
        finally:
            oracle.pike_prefilter(True)
        po.close()
        pr.close()
    differ = 0
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        on = oracle.pike(p, b["subject_b"])
        oracle.pike_prefilter(False)
        try:
            off = oracle.pike(p, b["subject_b"])
        finally:
            oracle.pike_prefilter(True)
        differ += on != off
        p.close()
    assert differ == 0
