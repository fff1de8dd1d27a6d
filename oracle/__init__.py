"""Test infrastructure only: the CPU oracle (sre_oracle.c), the reference build
recipe (_ref/), the lowering checker (lower_check.cpp) and the line-batch CPU
runner (cpu_baseline.py).  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by sregex_b200."""
import os as _os

from sregex_b200 import capi as _capi

_HERE = _os.path.dirname(_os.path.abspath(__file__))
ORACLE_LIB = _os.path.join(_HERE, "liboracle.so")
REF_LIB = _os.path.join(_HERE, "_ref", "libsregex_ref.so")
# the checkers speak the same C API as the product: loadable through capi.load("oracle" / "ref")
_capi.register("oracle", ORACLE_LIB)
_capi.register("ref", REF_LIB)
