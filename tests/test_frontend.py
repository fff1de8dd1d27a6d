"""Host front end (parser + compiler) parity: program dump text and syntax
error offsets against the reference, over every regex of its t/ suite."""
from conftest import runnable
from sregex_b200 import capi


def test_program_dumps_match_reference(golden, oracle):
    n = 0
    for b in runnable(golden):
        p = oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        assert p.ncaps == b["ncaps"], (b["file"], b["name"])
        assert p.dump() == b["dump"], (b["file"], b["name"])
        p.close()
        n += 1
    assert n >= 1918


def test_syntax_errors_match_reference(golden, oracle):
    n = 0
    for b in golden["blocks"]:
        if "error" not in b:
            continue
        try:
            oracle.compile(b["regexes_b"], b["flags"], multi=b["multi"])
        except capi.SreSyntaxError as e:
            assert {"offset": e.offset, "regex_id": e.regex_id} == b["error"], (b["file"], b["name"])
            n += 1
        else:
            raise AssertionError(("no error", b["file"], b["name"]))
    assert n >= 75


def test_flags_newline_and_counted(oracle, ref):
    cases = [(rb"a.c", capi.SRE_REGEX_NEWLINE), (rb"\C[^a]x", capi.SRE_REGEX_NEWLINE),
             (rb"[a-c]{2,4}?x{0}y{1}z{3,}", 0), (rb"(a|b)*?c+?d??", capi.SRE_REGEX_CASELESS),
             (rb"\x41\x{42}\o{103}\101\cA\e[\d\W-x]", capi.SRE_REGEX_CASELESS),
             (rb"(?:a|)(|b)\b\B\A\z^$", 0), (rb"[]a][^]a][a\-z][-a][a-]", 0)]
    for rx, fl in cases:
        assert oracle.compile(rx, fl).dump() == ref.compile(rx, fl).dump(), rx


def test_random_regex_fuzz_against_live_reference(oracle, ref):
    """Seeded random regex strings over the metacharacter alphabet: same program
    dump, same capture count, or the same syntax-error offset as the reference
    (needs oracle/_ref, i.e. runs in the build container)."""
    import random
    rng = random.Random(0x5EED)
    atoms = ["a", "b", "c", "Z", "0", "9", " ", "\n", ".", "^", "$", "|", "(", ")", "(?:", "[", "]", "[^", "-",
             "*", "+", "?", "{2}", "{1,3}", "{0,}", "{,2}", "{3,1}", "{", "}", ":", "\\d", "\\D", "\\w", "\\W", "\\s",
             "\\S", "\\b", "\\B", "\\A", "\\z", "\\Z", "\\h", "\\H", "\\v", "\\V", "\\N", "\\C", "\\n", "\\t", "\\e",
             "\\x41", "\\x{4a}", "\\x{123}", "\\o{101}", "\\o{777}", "\\101", "\\0", "\\1", "\\8", "\\cA", "\\c",
             "\\", "\\\\", "\\.", "\\[", "\\-", "\\#", "\\p", "é", "{500}", "{499}", "a{2}{3}", "x**"]
    same = errs = 0
    for i in range(4000):
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode("latin-1")
        if b"\0" in rx:
            continue
        flags = rng.choice([0, 0, capi.SRE_REGEX_CASELESS, capi.SRE_REGEX_NEWLINE, 3])
        try:
            pr = ref.compile(rx, flags)
            want = ("ok", pr.ncaps, pr.dump())
            pr.close()
        except capi.SreSyntaxError as e:
            want = ("err", e.offset)
            errs += 1
        try:
            po = oracle.compile(rx, flags)
            got = ("ok", po.ncaps, po.dump())
            po.close()
        except capi.SreSyntaxError as e:
            got = ("err", e.offset)
        assert got == want, (rx, flags, got[:2], want[:2])
        same += 1
    assert same > 3500 and 500 < errs < 3500


def test_random_multi_regex_fuzz_against_live_reference(oracle, ref):
    import random
    rng = random.Random(7)
    pool = [rb"a(b)c", rb"x|y(z)", rb"\d+", rb"(", rb"[a-", rb"(?:q)*?", rb"^w$", rb"(a)(b)(c)", rb"", rb"a{2,1}"]
    for _ in range(300):
        pats = [rng.choice(pool) for _ in range(rng.randrange(1, 5))]
        flags = [rng.choice([0, 1]) for _ in pats]
        try:
            pr = ref.compile(pats, flags, multi=True)
            want = ("ok", pr.ncaps, pr.dump())
        except capi.SreSyntaxError as e:
            want = ("err", e.offset, e.regex_id)
        try:
            po = oracle.compile(pats, flags, multi=True)
            got = ("ok", po.ncaps, po.dump())
        except capi.SreSyntaxError as e:
            got = ("err", e.offset, e.regex_id)
        assert got == want, (pats, flags)
