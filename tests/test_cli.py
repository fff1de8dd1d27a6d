"""tools/sregex-cli.c (interface + output format of the reference's src/sre_cli.c,
what t/SRegex.pm parses) linked against different implementations of the sregex
API: the CPU oracle here, libsregex_cuda on the GPU box."""
import os
import subprocess

import pytest

from conftest import runnable
from sregex_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOLS = os.path.join(ROOT, "tools")


def _build(target):
    subprocess.check_call(["make", "-C", TOOLS, target], stdout=subprocess.DEVNULL)
    return os.path.join(TOOLS, target)


def _expected_lines(b, jit):
    def verdict(rc):
        return {capi.SRE_OK: "match", capi.SRE_DECLINED: "no match"}[rc]

    def pike(rc, ov):
        if rc < 0:
            return "no match"
        return f"match {rc} " + " ".join(f"({ov[i]}, {ov[i + 1]})" for i in range(0, len(ov), 2))

    temp = ""
    for k, t in enumerate(b["pike_split"]["trace"]):
        if k % 2 == 1:      # the CLI prints after 1-byte chunks only
            temp += f"[({t[1]}, {t[2]})]" + (f"({t[3]}, {t[4]}) " if t[3] is not None else " ")
    out = [f"thompson {verdict(b['thompson'])}", f"splitted thompson {verdict(b['thompson_split'][-1])}"]
    if jit:
        out += [f"jitted thompson {verdict(b['jit'])}", f"splitted jitted thompson {verdict(b['jit_split'][-1])}"]
    else:
        out += ["jitted thompson disabled", "splitted jitted thompson disabled"]
    out += [f"pike {pike(b['pike']['rc'], b['pike']['ov'])}",
            f"splitted pike {temp}{pike(b['pike_split']['rc'], b['pike_split']['ov'])}"]
    return out


def _run_block(cli, b, jit):
    flags = []
    if any(b["flags"]):
        flags = ["--flags", " ".join("i" if f else "" for f in b["flags"])]
    args = [cli, "--stdin"] + flags + (["-n", str(len(b["regexes_b"]))] if b["multi"] else [])
    args += [r.decode("latin-1").encode("latin-1") for r in b["regexes_b"]]
    s = b["subject_b"]
    res = subprocess.run(args, input=str(len(s)).encode() + b"\n" + s, capture_output=True, timeout=120)
    assert res.returncode == 0, res.stderr
    text = res.stdout.decode("latin-1")
    head, _, tail = text.partition("\n## ")
    assert head.split("\n", 2)[2].rstrip("\n") + "\n" == b["dump"], b["name"]       # AST, captures, then the dump
    # the "## <subject> (len N)" header may itself contain newlines
    body = tail[tail.index(f"(len {len(s)})\nthompson ") + len(f"(len {len(s)})\n"):]
    got = body.rstrip("\n").split("\n")
    assert got == _expected_lines(b, jit), (b["file"], b["name"], got)


def _usable(b):
    return (b"\n" not in b["subject_b"] or True) and all(r and not r.startswith(b"-") for r in b["regexes_b"])


def test_cli_over_oracle_matches_reference_cli_output(golden):
    cli = _build("sregex-cli-oracle")
    blocks = [b for b in runnable(golden) if _usable(b)]
    for b in blocks[::9]:
        _run_block(cli, b, jit=False)


def test_cli_reports_syntax_errors_like_the_reference(golden):
    cli = _build("sregex-cli-oracle")
    n = 0
    for b in golden["blocks"]:
        if "error" not in b or b["multi"] or not _usable(b):
            continue
        res = subprocess.run([cli, b["regexes_b"][0], b"x"], capture_output=True)
        assert res.returncode == 1
        assert res.stderr.decode() == f"[error] syntax error at pos {b['error']['offset']}\n", b["name"]
        n += 1
    assert n > 50


@pytest.mark.gpu
def test_cli_over_libsregex_cuda(golden):
    """the same CLI source linked against the GPU library: six modes per subject"""
    cli = _build("sregex-cli-cuda")
    blocks = [b for b in runnable(golden) if _usable(b)]
    for b in blocks[::160]:
        _run_block(cli, b, jit=True)
