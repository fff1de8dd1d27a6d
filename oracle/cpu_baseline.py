"""CPU checkers over line batches: the oracle port (oracle/liboracle.so) or the
unmodified reference (oracle/_ref/libsregex_ref.so).  TEST / BASELINE USE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline and
`--impl reference` legs, never by the product path (sregex_b200/cuda.py)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from sregex_b200 import capi

from . import ORACLE_LIB, REF_LIB

ENGINE_THOMPSON, ENGINE_JIT, ENGINE_PIKE = 0, 1, 2


def _bind(which):
    sl = capi.load(which)
    fn = sl.L.ref_bench_lines_reps
    fn.restype = C.c_double
    fn.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                   C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
    return fn


def available(which: str) -> bool:
    return os.path.exists({"ref": REF_LIB, "oracle": ORACLE_LIB}[which])


def run_stream(which, regexes, flags, buf: np.ndarray, chunk, engine, reps=1):
    """one ctx fed buf in `chunk`-byte calls with SRE_AGAIN carry, on one core
    -> (seconds of the calls, rc of the last call, number of calls made)"""
    sl = capi.load(which)
    fn = sl.L.ref_bench_stream
    fn.restype = C.c_double
    fn.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                   C.c_size_t, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_long)]
    if isinstance(regexes, bytes):
        regexes = [regexes]
    flags = flags or [0] * len(regexes)
    arr = (C.c_char_p * len(regexes))(*regexes)
    fl = (C.c_int * len(regexes))(*flags)
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    rc, calls = C.c_int(-99), C.c_long(-1)
    secs = fn(arr, fl, len(regexes), engine, buf.ctypes.data, buf.size, chunk, reps, C.byref(rc), C.byref(calls))
    if secs < 0:
        raise RuntimeError("baseline stream run failed")
    return secs, rc.value, calls.value + 1


def run_lines(which, regexes, flags, buf: np.ndarray, nlines, pitch, linelen, engine, nthreads=1,
              ovec_slots=0, reps=1):
    """-> (seconds of the matching alone, over all reps; rc int32[nlines]; ovec int64[nlines, ovec_slots] or
    None).  Thread start, parse, compile and JIT happen before the clock starts."""
    fn = _bind(which)
    if isinstance(regexes, bytes):
        regexes = [regexes]
    flags = flags or [0] * len(regexes)
    arr = (C.c_char_p * len(regexes))(*regexes)
    fl = (C.c_int * len(regexes))(*flags)
    buf = np.ascontiguousarray(buf, dtype=np.uint8)
    rc = np.full(nlines, -99, dtype=np.int32)
    ov = np.full((nlines, ovec_slots), -99, dtype=np.int64) if ovec_slots else None
    secs = fn(arr, fl, len(regexes), engine, buf.ctypes.data, nlines, pitch, linelen, nthreads, reps,
              rc.ctypes.data, ov.ctypes.data if ov is not None else None, ovec_slots)
    if secs < 0:
        raise RuntimeError("baseline run failed")
    return secs, rc, ov
