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


# ---- the reference's own callers, compiled unmodified against libsregex_cuda -----------------
# (oracle/Makefile `callers`: /root/reference/src/sre_cli.c and bench/sregex.c, binaries under
# the git-ignored oracle/_ref/; built in the build container, they travel to the GPU box)

REF_BIN = os.path.join(ROOT, "oracle", "_ref")


def _caller(name):
    path = os.path.join(REF_BIN, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not built (reference sources absent)")
    return path


@pytest.mark.gpu
def test_unmodified_reference_cli_over_libsregex_cuda(golden):
    """src/sre_cli.c itself (what t/SRegex.pm drives), re-linked with -lsregex_cuda: the six
    modes per subject print what the reference's own build prints"""
    cli = _caller("sre-cli-cuda")
    blocks = [b for b in runnable(golden) if _usable(b)]
    for b in blocks[::40]:
        _run_block(cli, b, jit=True)


def _gen_data(path, repeat):
    with open(path, "wb") as f:                 # bench/gen-data.pl:9
        f.write(b"abccc" * repeat + b"aaabbccb")
    return 5 * repeat + 8


def _bench_times(out):
    import re
    t = {}
    for line in out.splitlines():
        m = re.match(r"sregex (Thompson JIT|Thompson|Pike) (.*): ([0-9.]+) ms elapsed\.", line)
        assert m, line
        t[m.group(1)] = (m.group(2), float(m.group(3)))
    return t


@pytest.mark.gpu
def test_unmodified_reference_bench_over_libsregex_cuda(tmp_path):
    """bench/sregex.c itself (BASELINE configs[0]), re-linked with -lsregex_cuda, over the
    bench/gen-data.pl buffer at its shipped size and scaled to 16 MiB: same verdicts and the same
    Pike offsets as the reference; the times go to gpurun_out/ beside the CPU's"""
    import json
    bench, ref = _caller("bench-sregex-cuda"), _caller("bench-sregex-ref")
    regex = "(?:a|b)aa(?:aa|bb)cc(?:a|b)"       # bench/Makefile:62
    report = {}
    for repeat in (1048576, 3355443):
        n = _gen_data(tmp_path / "abc.txt", repeat)
        row = {}
        for name, exe in (("gpu", bench), ("cpu", ref)):
            res = subprocess.run([exe, "--thompson", "--thompson-jit", "--pike", regex, str(tmp_path / "abc.txt")],
                                 capture_output=True, timeout=600)
            assert res.returncode == 0, res.stderr
            t = _bench_times(res.stdout.decode())
            assert t["Thompson"][0] == "match" and t["Thompson JIT"][0] == "match", t
            assert t["Pike"][0] == f"match ({n - 8}, {n})", t
            row[name] = {k: {"ms": v[1], "MB_per_s": n / v[1] / 1e3} for k, v in t.items()}
        report[str(n)] = row
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "reference_bench_unmodified.json"), "w") as f:
        json.dump(report, f, indent=1)


def test_unmodified_reference_bench_builds_against_the_product_header():
    """CPU tier: the binaries exist after build() in the build container and are linked against
    libsregex_cuda.so (not against the reference library)"""
    bench = _caller("bench-sregex-cuda")
    out = subprocess.run(["ldd", bench], capture_output=True).stdout.decode()
    assert "libsregex_cuda.so" in out and "libsregex_ref" not in out
    # the same source against the reference library prints the reference's answer (tiny buffer)
    ref = _caller("bench-sregex-ref")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        n = _gen_data(os.path.join(d, "abc.txt"), 1000)
        res = subprocess.run([ref, "--thompson", "--pike", "(?:a|b)aa(?:aa|bb)cc(?:a|b)", os.path.join(d, "abc.txt")],
                             capture_output=True)
        assert res.returncode == 0 and f"match ({n - 8}, {n})" in res.stdout.decode()
