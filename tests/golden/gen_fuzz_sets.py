#!/usr/bin/env python3
"""Generate tests/golden/fuzz_sets.json.gz: what the UNMODIFIED reference answers on seeded random
regexes and regex SETS that its own t/ corpus does not hold -- look-ahead assertions that can be
skipped or looped ((\\B)?, (?:\\B|x)+, $?), members that match the empty string, SRE_REGEX_CASELESS,
and the first-byte prefilter misfire family (a+b?, \\d+\\.?) -- so that the oracle stays pinned on
them where /root/reference does not exist (the GPU box).

Run in the build container only (needs oracle/_ref/libsregex_ref.so, `make -C oracle ref`):

    python tests/golden/gen_fuzz_sets.py

Per case: regexes, flags, subject, Thompson rc, Pike (rc, ovector), and the rc sequences of both
fed in 3-byte chunks (Pike: the last rc and its ovector).
"""
import gzip
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: F401,E402  (registers the checker libraries with capi)
from sregex_b200 import capi  # noqa: E402

ATOMS = ["a", "b", "A", "ab", " ", "_", ".", "|", "(", ")", "(?:", "*", "+", "?", "*?", "+?", "{2}", "{0,2}", "[ab]",
         "[^a]", "\\w", "\\W", "\\d", "\\s", "1", "(a)", "(b*)", "(a|ab)", "(\\w+)", "()", "^", "\\A", "\\n", "$",
         "\\z", "\\b", "\\B", "\\b?", "\\B?", "$?", "^?", "(?:\\b|)", "(?:$|a)", "(\\B)?", "(?:\\B|x)+", "x?", "a+"]
HEADS = [b"a+", b"\\w+", b"\\d+", b"[ab]+", b"(a+)", b"(\\w)+", b"(?:a|b)+", b"^a+", b"\\ba+", b"(a|ab)+"]
TAILS = [b"b?", b"x?", b"\\.?", b"(\\d)?", b" ?", b"(?:ab)?", b"b*", b"$", b"", b"\\b", b"(b|c)?"]
ALPHABET = b"abAB _1.\n\nx"


def main():
    ref = capi.load("ref")
    rng = random.Random(20261019)
    cases = []
    while len(cases) < 4000:
        if rng.random() < 0.2:
            rxs = [rng.choice(HEADS) + rng.choice(TAILS)]
        else:
            k = 1 if rng.random() < 0.8 else rng.randrange(2, 4)
            rxs = ["".join(rng.choice(ATOMS) for _ in range(rng.randrange(1, 9))).encode() for _ in range(k)]
        flags = capi.SRE_REGEX_CASELESS if rng.random() < 0.25 else 0
        multi = len(rxs) > 1
        try:
            p = ref.compile(rxs if multi else rxs[0], flags)
        except capi.SreSyntaxError:
            continue
        for _ in range(2):
            s = bytes(rng.choice(ALPHABET) for _ in range(rng.randrange(0, 40)))
            chunks = [(s[i:i + 3], i + 3 >= len(s)) for i in range(0, len(s), 3)] or [(b"", True)]
            rc, ov = ref.pike(p, s)
            cases.append({"regexes": [r.hex() for r in rxs], "flags": flags, "subject": s.hex(),
                          "thompson": ref.thompson(p, s), "pike": [rc, ov],
                          "thompson_chunked": ref.thompson(p, s, chunks),
                          "pike_chunked": ref.pike(p, s, chunks)})
        p.close()
    out = os.path.join(ROOT, "tests", "golden", "fuzz_sets.json.gz")
    with gzip.open(out, "wt", compresslevel=9) as f:
        json.dump({"generator": "tests/golden/gen_fuzz_sets.py", "seed": 20261019, "cases": cases}, f,
                  separators=(",", ":"))
    print(f"{len(cases)} cases -> {out} ({os.path.getsize(out)} bytes)")


if __name__ == "__main__":
    main()
