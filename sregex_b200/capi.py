"""ctypes binding of the sregex C API (reference src/sregex/sregex.h:82-171).

The binding drives any shared library that exports that API.  The product
knows one: ``sregex_b200/libsregex_cuda.so`` (host front end + CUDA VMs).  The
test checkers (the CPU restatement and the unmodified reference) register their
own libraries from ``oracle/__init__.py``; with the same binding over all three
the parity tests read like the reference's own CLI driver
(src/sre_cli.c:299-660): same calls, same argument meaning, same status codes.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

SRE_OK, SRE_ERROR, SRE_AGAIN, SRE_BUSY, SRE_DONE, SRE_DECLINED = 0, -1, -2, -3, -4, -5
SRE_REGEX_CASELESS, SRE_REGEX_NEWLINE = 1, 2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_LIB = os.path.join(ROOT, "sregex_b200", "libsregex_cuda.so")
_paths = {"cuda": CUDA_LIB}


def register(name: str, path: str):
    """make another implementation of the sregex API loadable by name (test checkers)"""
    _paths[name] = path

_intp = C.POINTER(C.c_ssize_t)


class SreSyntaxError(Exception):
    def __init__(self, offset: int, regex_id: int):
        super().__init__(f"regex {regex_id}: syntax error at pos {offset}")
        self.offset, self.regex_id = offset, regex_id


@dataclass
class Program:
    lib: "SreLib"
    pool: int            # pool that owns the program
    prog: int
    ncaps: int           # max over the regexes of a set
    nregexes: int
    regexes: list = field(default_factory=list)
    flags: list = field(default_factory=list)

    @property
    def nslots(self) -> int:
        return 2 * (self.ncaps + 1)

    def dump(self) -> str:
        return self.lib.program_dump(self)

    def close(self):
        if self.pool:
            self.lib.L.sre_destroy_pool(self.pool)
            self.pool = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SreLib:
    """One loaded implementation of the sregex API."""

    def __init__(self, path: str):
        self.path = path
        self.L = L = C.CDLL(path)
        vp, cp, sz, i = C.c_void_p, C.c_char_p, C.c_size_t, C.c_ssize_t
        sig = {
            "sre_create_pool": (vp, [sz]),
            "sre_reset_pool": (None, [vp]),
            "sre_destroy_pool": (None, [vp]),
            "sre_regex_parse": (vp, [vp, cp, C.POINTER(sz), C.c_int, _intp]),
            "sre_regex_parse_multi": (vp, [vp, C.POINTER(cp), i, C.POINTER(sz),
                                           C.POINTER(C.c_int), _intp, _intp]),
            "sre_regex_dump": (None, [vp]),
            "sre_regex_compile": (vp, [vp, vp]),
            "sre_program_dump": (None, [vp]),
            "sre_vm_pike_create_ctx": (vp, [vp, vp, _intp, sz]),
            "sre_vm_pike_exec": (i, [vp, vp, sz, C.c_uint, C.POINTER(_intp)]),
            "sre_vm_thompson_create_ctx": (vp, [vp, vp]),
            "sre_vm_thompson_exec": (i, [vp, vp, sz, C.c_uint]),
            "sre_vm_thompson_jit_compile": (i, [vp, vp, C.POINTER(vp)]),
            "sre_vm_thompson_jit_create_ctx": (vp, [vp, vp]),
            "sre_vm_thompson_jit_get_handler": (vp, [vp]),
            "sre_vm_thompson_jit_free": (i, [vp]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        self.api_symbols = list(sig)
        # dump-to-string helper: ours or the reference shim's
        for name in ("sre_program_dump_str", "ref_program_dump_str"):
            if hasattr(L, name):
                self._dump = getattr(L, name)
                self._dump.restype, self._dump.argtypes = vp, [vp]
                break
        else:
            self._dump = None
        self._free = C.CDLL(None).free
        self._free.argtypes = [vp]

    # -- front end -----------------------------------------------------------
    def compile(self, regexes, flags=None, multi=None) -> Program:
        """regexes: bytes or list[bytes]; flags: int or list[int]."""
        if isinstance(regexes, (bytes, str)):
            regexes = [regexes]
        regexes = [r.encode("latin-1") if isinstance(r, str) else r for r in regexes]
        if flags is None:
            flags = [0] * len(regexes)
        elif isinstance(flags, int):
            flags = [flags] * len(regexes)
        if multi is None:
            multi = len(regexes) > 1
        L = self.L
        ppool = L.sre_create_pool(1024)
        ncaps, err = C.c_size_t(0), C.c_ssize_t(-1)
        try:
            if not multi:
                re = L.sre_regex_parse(ppool, regexes[0], C.byref(ncaps), flags[0], C.byref(err))
                if not re:
                    raise SreSyntaxError(err.value, 0)
            else:
                arr = (C.c_char_p * len(regexes))(*regexes)
                fl = (C.c_int * len(regexes))(*flags)
                eid = C.c_ssize_t(-1)
                re = L.sre_regex_parse_multi(ppool, arr, len(regexes), C.byref(ncaps), fl,
                                             C.byref(err), C.byref(eid))
                if not re:
                    raise SreSyntaxError(err.value, eid.value)
            cpool = L.sre_create_pool(1024)
            prog = L.sre_regex_compile(cpool, re)
            if not prog:
                L.sre_destroy_pool(cpool)
                raise RuntimeError("sre_regex_compile failed")
        finally:
            L.sre_destroy_pool(ppool)   # the parser pool may go after compile (sre_cli.c:198)
        return Program(self, cpool, prog, ncaps.value, len(regexes), list(regexes), list(flags))

    def program_dump(self, p: Program) -> str:
        s = self._dump(p.prog)
        try:
            return C.string_at(s).decode("latin-1")
        finally:
            self._free(s)

    # -- Thompson ------------------------------------------------------------
    def thompson(self, p: Program, data: bytes, chunks=None, jit=False):
        """Run the Thompson VM.  chunks=None: one exec(data, eof=1) -> rc.
        Otherwise `chunks` is a list of (bytes, eof) fed in order; returns the
        list of rcs up to and including the first rc != SRE_AGAIN."""
        L = self.L
        pool = L.sre_create_pool(1024)
        code = C.c_void_p()
        try:
            if jit:
                rc = L.sre_vm_thompson_jit_compile(pool, p.prog, C.byref(code))
                if rc != SRE_OK:
                    return None
                ctx = L.sre_vm_thompson_jit_create_ctx(pool, p.prog)
                fn = C.CFUNCTYPE(C.c_ssize_t, C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint)(
                    L.sre_vm_thompson_jit_get_handler(code))
            else:
                ctx = L.sre_vm_thompson_create_ctx(pool, p.prog)
                fn = L.sre_vm_thompson_exec
            if not ctx:
                raise RuntimeError("create_ctx failed")
            if chunks is None:
                buf = C.create_string_buffer(data, len(data)) if data else None
                return fn(ctx, C.cast(buf, C.c_void_p) if buf else None, len(data), 1)
            rcs = []
            for chunk, eof in chunks:
                buf = C.create_string_buffer(chunk, len(chunk)) if chunk else None
                rc = fn(ctx, C.cast(buf, C.c_void_p) if buf else None, len(chunk), int(eof))
                rcs.append(rc)
                if rc != SRE_AGAIN:
                    break
            return rcs
        finally:
            if code:
                L.sre_vm_thompson_jit_free(code)
            L.sre_destroy_pool(pool)

    def pike_prefilter(self, on: bool):
        """checker libraries only (oracle): run the Pike VM with / without the reference's
        first-byte prefilter (oracle/sre_oracle.c: oracle_pike_prefilter)"""
        self.L.oracle_pike_prefilter.argtypes = [C.c_int]
        self.L.oracle_pike_prefilter.restype = None
        self.L.oracle_pike_prefilter(1 if on else 0)

    # -- Pike ----------------------------------------------------------------
    def pike(self, p: Program, data: bytes, chunks=None):
        """chunks=None: one exec(data, eof=1) -> (rc, ovector list).
        Otherwise returns a list of per-call records
        (rc, ov0, ov1, pending0, pending1) (pending = None when not reported)
        followed by the final (rc, ovector)."""
        L = self.L
        pool = L.sre_create_pool(1024)
        n = p.nslots
        ov = (C.c_ssize_t * n)(*([-99] * n))
        try:
            ctx = L.sre_vm_pike_create_ctx(pool, p.prog, ov, n * C.sizeof(C.c_ssize_t))
            if not ctx:
                raise RuntimeError("create_ctx failed")
            if chunks is None:
                buf = C.create_string_buffer(data, len(data)) if data else None
                rc = L.sre_vm_pike_exec(ctx, C.cast(buf, C.c_void_p) if buf else None,
                                        len(data), 1, None)
                return rc, (list(ov) if rc >= 0 else None)
            trace = []
            for chunk, eof in chunks:
                pend = _intp()
                buf = C.create_string_buffer(chunk, len(chunk)) if chunk else None
                rc = L.sre_vm_pike_exec(ctx, C.cast(buf, C.c_void_p) if buf else None,
                                        len(chunk), int(eof), C.byref(pend))
                if rc == SRE_AGAIN:
                    trace.append((rc, ov[0], ov[1],
                                  pend[0] if pend else None, pend[1] if pend else None))
                    continue
                return trace, rc, (list(ov) if rc >= 0 else None)
            return trace, SRE_AGAIN, None
        finally:
            L.sre_destroy_pool(pool)


def pike_all(lib: "SreLib", p: Program, data: bytes, limit: int = 64):
    """All non-overlapping matches through the classic API: after a match the
    same ctx is given the rest of the data from ovector[1] on (post-match
    continuation, sre_vm_pike.c:624-635).  -> list of (rc, start, end)"""
    L = lib.L
    pool = L.sre_create_pool(1024)
    n = p.nslots
    ov = (C.c_ssize_t * n)(*([-99] * n))
    out = []
    try:
        ctx = L.sre_vm_pike_create_ctx(pool, p.prog, ov, n * C.sizeof(C.c_ssize_t))
        at = 0
        buf = C.create_string_buffer(data, len(data) + 1)
        while len(out) < limit:
            rc = L.sre_vm_pike_exec(ctx, C.cast(C.addressof(buf) + at, C.c_void_p), len(data) - at, 1, None)
            if rc < 0:
                break
            out.append((rc, ov[0], ov[1]))
            at = ov[1]
    finally:
        L.sre_destroy_pool(pool)
    return out


def split_chunks(data: bytes):
    """The reference CLI's "splitted" feeding pattern (src/sre_cli.c:369-385):
    an empty non-eof chunk before every 1-byte chunk, then an empty eof chunk."""
    out = []
    for i in range(len(data)):
        out.append((b"", False))
        out.append((data[i:i + 1], False))
    out.append((b"", True))
    return out


_cache: dict = {}


def load(which: str) -> SreLib:
    if which not in _paths:
        raise KeyError(f"no sregex library registered as {which!r} (the checkers register theirs on "
                       f"`import oracle`)")
    path = _paths[which]
    if path not in _cache:
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        _cache[path] = SreLib(path)
    return _cache[path]
