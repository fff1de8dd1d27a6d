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
