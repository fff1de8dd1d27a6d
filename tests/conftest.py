import gzip
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """Build the checkers (and the product library) if they are not there yet.
    On the GPU box the prebuilt .so files travel with the snapshot."""
    import oracle
    from sregex_b200 import capi
    if not os.path.exists(oracle.ORACLE_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    if not os.path.exists(oracle.REF_LIB) and os.path.isdir("/root/reference/src/sregex"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(capi.CUDA_LIB) and os.path.exists(
            os.path.join(ROOT, "sregex_b200", "csrc", "Makefile")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "sregex_b200", "csrc")])


@pytest.fixture(scope="session", autouse=True)
def built():
    _ensure_built()


@pytest.fixture(scope="session")
def golden():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "t_suite.json.gz")) as f:
        g = json.load(f)
    for b in g["blocks"]:
        b["regexes_b"] = [bytes.fromhex(r) for r in b["regexes"]]
        b["subject_b"] = bytes.fromhex(b["subject"])
    return g


@pytest.fixture(scope="session")
def oracle():
    from sregex_b200 import capi
    return capi.load("oracle")


@pytest.fixture(scope="session")
def ref():
    import oracle
    from sregex_b200 import capi
    if not os.path.exists(oracle.REF_LIB):
        pytest.skip("oracle/_ref not built (reference sources absent)")
    return capi.load("ref")


@pytest.fixture(scope="session")
def cuda():
    from sregex_b200 import capi
    return capi.load("cuda")


@pytest.fixture(scope="module")
def leftmost_first():
    """For the CPU models of the fast Pike tiers (closure tables, P-DFA), which compute the
    leftmost-first match: the oracle without the reference's first-byte prefilter, i.e. what the
    Pike VM computes when that shortcut does not misfire (oracle/sre_oracle.c:
    oracle_pike_prefilter).  The product adds the misfire back in a separate pass
    (k_pike_quirk_mark + the faithful general kernel); the GPU tests compare with the oracle WITH
    the prefilter, which is pinned to the reference in tests/test_oracle.py."""
    from sregex_b200 import capi
    o = capi.load("oracle")
    o.pike_prefilter(False)
    yield o
    o.pike_prefilter(True)


def runnable(golden):
    return [b for b in golden["blocks"] if "skip" not in b and "error" not in b]
