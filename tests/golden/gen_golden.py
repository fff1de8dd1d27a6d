#!/usr/bin/env python3
"""Generate tests/golden/t_suite.json.gz from the reference's own test corpus.

Run in the build container only (needs /root/reference, perl, and
oracle/_ref/libsregex_ref.so built by `make -C oracle ref`):

    python tests/golden/gen_golden.py

For every block of /root/reference/t/*.t (format: t/SRegex.pm, Test::Base) it
records the inputs (regexes, flags, subject) and what the UNMODIFIED reference
library answers through its public API, in the six modes of the reference CLI
(src/sre_cli.c:299-660): Thompson, Thompson fed 1-byte chunks, JIT, JIT
chunked, Pike, Pike chunked (with temp captures / pending matches), plus the
program dump text and, for syntax-error blocks, the error offset / regex id.
Explicit expectations written in the .t files (--- cap, --- match_id,
--- temp_cap, --- no_match, --- err) are checked against the reference here
and stored too, so the fixture is pinned to the reference's golden vectors.
"""
import glob
import gzip
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: F401  (registers the checker libraries with capi)
from sregex_b200 import capi  # noqa: E402

REF_T = "/root/reference/t"


def read_blocks(path):
    text = open(path, "rb").read().decode("latin-1")
    data = text.split("__DATA__", 1)[1]
    blocks = []
    for chunk in re.split(r"^=== ", data, flags=re.M)[1:]:
        name, _, body = chunk.partition("\n")
        secs = {}
        for m in re.finditer(r"^--- (\w+)((?: \w+)*)(?::[ \t]*(.*))?\n((?:(?!--- ).*\n?)*)",
                             body, flags=re.M):
            key, filters, inline, multi = m.group(1), m.group(2).split(), m.group(3), m.group(4)
            if inline is not None:
                val = inline.strip()
            else:
                # Test::Base default `trim` filter: drop blank lines around the value
                val = re.sub(r"\A(?:[ \t]*\n)+", "", multi)
                val = re.sub(r"(?<=\n)\s*\Z", "", val)
                if "chop" in filters or "chomp" in filters:
                    val = val[:-1] if val.endswith("\n") else val
            secs[key] = (val, filters)
        blocks.append((name.strip(), secs))
    return blocks


def perl_eval_many(exprs):
    """Evaluate Perl expressions; returns for each a list of byte strings
    (array refs flattened) -- the `eval` filter of Test::Base."""
    script = r'''
use strict; use warnings; no warnings;
binmode(STDOUT);
local $/ = "\0";
while (my $e = <STDIN>) {
    chomp $e;
    my $v = eval $e;
    die "eval failed: $e: $@" if $@;
    my @v = ref $v eq 'ARRAY' ? @$v : ($v);
    print scalar(@v), "\n";
    for my $x (@v) { utf8::encode($x) if utf8::is_utf8($x) && $x =~ /[^\x00-\xff]/; utf8::downgrade($x, 1); print unpack("H*", $x), "\n"; }
}
'''
    inp = b"".join(e.encode("latin-1") + b"\0" for e in exprs)
    out = subprocess.run(["perl", "-e", script], input=inp, capture_output=True, check=True).stdout
    lines = out.decode().split("\n")
    res, k = [], 0
    for _ in exprs:
        n = int(lines[k]); k += 1
        res.append([bytes.fromhex(lines[k + j]) for j in range(n)])
        k += n
    return res


def fmt_cap(ov, nslots):
    return " ".join(f"({ov[i]}, {ov[i + 1]})" for i in range(0, nslots, 2))


def main():
    ref = capi.load("ref")
    files = sorted(glob.glob(os.path.join(REF_T, "*.t")))
    raw = []
    for f in files:
        for name, secs in read_blocks(f):
            raw.append((os.path.basename(f), name, secs))

    # batch all perl evals
    exprs, slots = [], []
    for bi, (_, _, secs) in enumerate(raw):
        for key in ("re", "s", "flags"):
            if key in secs and "eval" in secs[key][1]:
                exprs.append(secs[key][0])
                slots.append((bi, key))
    evald = dict(zip(slots, perl_eval_many(exprs)))

    out, stats = [], {"blocks": 0, "runnable": 0, "errors": 0, "explicit_checked": 0, "skipped": 0}
    for bi, (fname, name, secs) in enumerate(raw):
        stats["blocks"] += 1
        if "SKIP" in secs:
            stats["skipped"] += 1
            continue

        def get(key):
            if key not in secs:
                return None
            if (bi, key) in evald:
                return evald[(bi, key)]
            return [secs[key][0].encode("latin-1")]

        regexes = get("re")
        subject = get("s")[0]
        multi = "re" in secs and "eval" in secs["re"][1] and secs["re"][0].lstrip().startswith("[")
        fl = get("flags")
        flag_str = fl[0].decode() if fl else ""
        flags = [0] * len(regexes)
        i = 0
        for ch in flag_str:             # src/sre_cli.c:680-711
            if ch == " ":
                i += 1
            elif ch == "i":
                flags[i] |= capi.SRE_REGEX_CASELESS
        rec = {"file": fname, "name": name, "regexes": [r.hex() for r in regexes],
               "flags": flags, "multi": bool(multi), "subject": subject.hex()}
        expect = {k: secs[k][0] for k in ("cap", "match_id", "temp_cap", "err", "err_like")
                  if k in secs and "eval" not in secs[k][1]}
        if "no_match" in secs:
            expect["no_match"] = True
        if "fatal" in secs:
            expect["fatal"] = True
        rec["expect"] = expect

        if any(b"\0" in r for r in regexes):
            rec["skip"] = "NUL in regex"
            out.append(rec)
            continue
        try:
            prog = ref.compile(regexes, flags, multi=multi)
        except capi.SreSyntaxError as e:
            rec["error"] = {"offset": e.offset, "regex_id": e.regex_id}
            stats["errors"] += 1
            if "err" in expect:
                want = expect["err"].strip()
                got = (f"[error] regex {e.regex_id}: syntax error at pos {e.offset}" if multi
                       else f"[error] syntax error at pos {e.offset}")
                assert want == got, (fname, name, want, got)
                stats["explicit_checked"] += 1
            out.append(rec)
            continue
        assert "err" not in expect, (fname, name, "expected a syntax error")

        stats["runnable"] += 1
        rec["ncaps"] = prog.ncaps
        rec["dump"] = prog.dump()
        chunks = capi.split_chunks(subject)
        rec["thompson"] = ref.thompson(prog, subject)
        rec["thompson_split"] = ref.thompson(prog, subject, chunks)
        rec["jit"] = ref.thompson(prog, subject, jit=True)
        rec["jit_split"] = ref.thompson(prog, subject, chunks, jit=True)
        rc, ov = ref.pike(prog, subject)
        rec["pike"] = {"rc": rc, "ov": ov}
        trace, rc2, ov2 = ref.pike(prog, subject, chunks)
        rec["pike_split"] = {"rc": rc2, "ov": ov2, "trace": [list(t) for t in trace]}

        # pin against the explicit expectations of the .t file
        if "cap" in expect:
            assert rc >= 0 and fmt_cap(ov, prog.nslots).startswith(expect["cap"].strip()) \
                or fmt_cap(ov, prog.nslots) == expect["cap"].strip(), (fname, name, expect["cap"], ov)
            stats["explicit_checked"] += 1
        if "match_id" in expect:
            assert rc == int(expect["match_id"]), (fname, name)
            stats["explicit_checked"] += 1
        if expect.get("no_match"):
            assert rc == capi.SRE_DECLINED, (fname, name)
            stats["explicit_checked"] += 1
        if "temp_cap" in expect:
            # the CLI prints only after 1-byte chunks (every 2nd call)
            s = ""
            for k, t in enumerate(trace):
                if k % 2 == 1:
                    s += f"[({t[1]}, {t[2]})]" + (f"({t[3]}, {t[4]}) " if t[3] is not None else " ")
            assert s.strip() == expect["temp_cap"].strip(), (fname, name, s, expect["temp_cap"])
            stats["explicit_checked"] += 1
        # the reference's six modes agree on the final verdict
        verdicts = {rec["thompson"], rec["thompson_split"][-1], rec["jit"], rec["jit_split"][-1],
                    capi.SRE_OK if rc >= 0 else rc, capi.SRE_OK if rc2 >= 0 else rc2}
        rec["modes_agree"] = len(verdicts) == 1 and ov == ov2 and rc == rc2
        prog.close()
        out.append(rec)

    path = os.path.join(ROOT, "tests", "golden", "t_suite.json.gz")
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps({"stats": stats, "blocks": out}, separators=(",", ":")).encode())
    print(stats, "->", path, os.path.getsize(path), "bytes")
    print("blocks where the reference's own modes disagree:",
          [(r["file"], r["name"]) for r in out if r.get("modes_agree") is False])


if __name__ == "__main__":
    main()
