"""Batch / device-pointer entry points of libsregex_cuda (include/sregex_cuda.h)
for torch tensors.  torch is used for device memory and streams only; the
matching is done by the library's own CUDA kernels."""
from __future__ import annotations

import ctypes as C

import torch

from . import capi

ENGINE_AUTO, ENGINE_DFA_TILED, ENGINE_DFA_GENERIC, ENGINE_NFA, ENGINE_DFA_SKIP, ENGINE_NFA_WARP = 0, 1, 2, 3, 4, 5
STATE_INIT = 0xFFFFFFFF
STATE_UNKNOWN = 0xFFFFFFFE
STREAM_FN_BYTES = 32
STREAM_HALO = 256


def engine_variant(engine: int, variant: int) -> int:
    """launch-shape variant of DFA_TILED / DFA_SKIP (tuning): engine | variant << 8"""
    return engine | ((variant & 0xFF) << 8)


class Info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "prog_len", "nfa_states", "nfa_classes", "nfa_kinds", "nfa_shift_states", "dfa_states",
        "dfa_classes", "dfa_byte_table", "dfa_leave_bytes", "nregexes", "pike_slots")] + [
            ("pike_ctx_bytes", C.c_uint64)] + [(n, C.c_uint32) for n in (
                "dfa_start", "dfa_acc", "image_states", "reserved")]


_lib = None


def lib():
    """The product library.  Loading never needs a GPU; compute calls do."""
    global _lib
    if _lib is None:
        sl = capi.load("cuda")
        L = sl.L
        vp, sz, i32p, i64p = C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p
        sig = {
            "sre_cuda_program_create": (vp, [vp]),
            "sre_cuda_program_info": (C.c_int, [vp, C.POINTER(Info)]),
            "sre_cuda_thompson_exec_lines": (C.c_int, [vp, vp, sz, sz, sz, i32p, C.c_int, vp]),
            "sre_cuda_thompson_exec_ragged": (C.c_int, [vp, vp, i64p, sz, i32p, C.c_int, vp]),
            "sre_cuda_pike_exec_lines": (C.c_int, [vp, vp, i64p, sz, sz, sz, i32p, i32p, i64p, sz, vp]),
            "sre_cuda_pike_exec_lines_all": (C.c_int, [vp, vp, i64p, sz, sz, sz, sz, i32p, i64p, i32p, vp]),
            "sre_cuda_thompson_exec_stream": (C.c_int, [vp, vp, sz, sz, C.c_uint, C.POINTER(C.c_uint32),
                                                        C.POINTER(C.c_int64), vp]),
            "sre_cuda_thompson_stream_reduce": (vp, [vp, vp, sz, vp, C.c_uint32, C.c_char_p, vp]),
            "sre_cuda_thompson_stream_resolve": (C.c_int, [vp, C.c_uint32, C.POINTER(C.c_uint32),
                                                           C.POINTER(C.c_int64), C.c_char_p]),
            "sre_cuda_thompson_stream_free": (None, [vp]),
            "sre_cuda_thompson_exec_stream_host": (C.c_int, [vp, vp, sz, sz, C.c_uint, C.POINTER(C.c_uint32),
                                                             C.POINTER(C.c_int64), sz]),
            "sre_cuda_stream_fn_apply": (C.c_uint32, [C.c_char_p, C.c_uint32]),
            "sre_cuda_dfa_fin": (C.c_int, [vp, C.c_uint32]),
            "sre_cuda_thompson_exec_lines_host": (C.c_int, [vp, vp, sz, sz, sz, i32p, C.c_int]),
            "sre_cuda_pike_exec_lines_host": (C.c_int, [vp, vp, sz, sz, sz, C.c_int, i32p, i64p, sz]),
            "sre_cuda_program_set_pike_tier": (None, [vp, C.c_int]),
            "sre_cuda_program_last_pike_tier": (C.c_int, [vp]),
            "sre_cuda_pike_streams_create": (vp, [vp, sz, vp]),
            "sre_cuda_pike_streams_exec": (C.c_int, [vp, vp, i64p, vp, C.c_uint, i64p, sz, vp]),
            "sre_cuda_pike_streams_free": (None, [vp]),
            "sre_cuda_thompson_exec_text": (C.c_int, [vp, vp, sz, i64p, i32p, sz, C.POINTER(C.c_size_t), vp]),
            "sre_cuda_index_lines": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                               C.POINTER(C.c_size_t), C.c_void_p]),
            "sre_cuda_launch_count": (C.c_long, [C.c_int]),
            "sre_cuda_device_available": (C.c_int, []),
            "sre_cuda_last_error": (C.c_char_p, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        sl.ext_symbols = list(sig)
        _lib = sl
    return _lib


class SreCudaError(RuntimeError):
    pass


def _check(rc):
    if rc != capi.SRE_OK:
        raise SreCudaError(lib().L.sre_cuda_last_error().decode() or f"rc={rc}")


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class CudaProgram:
    """A compiled regex (set) lowered to GPU tables."""

    def __init__(self, regexes, flags=None, multi=None):
        self.lib = lib()
        if not self.lib.L.sre_cuda_device_available():
            raise SreCudaError("no usable CUDA device (libsregex_cuda has no CPU fallback)")
        self.program = self.lib.compile(regexes, flags, multi)
        self.cp = self.lib.L.sre_cuda_program_create(self.program.prog)
        if not self.cp:
            raise SreCudaError(self.lib.L.sre_cuda_last_error().decode())
        self.info = Info()
        _check(self.lib.L.sre_cuda_program_info(self.cp, C.byref(self.info)))

    @property
    def nslots(self):
        return self.program.nslots

    def thompson_lines(self, buf: torch.Tensor, nlines: int, pitch: int, linelen: int,
                       engine: int = ENGINE_AUTO, out: torch.Tensor | None = None) -> torch.Tensor:
        assert buf.is_cuda and buf.dtype == torch.uint8
        rc = out if out is not None else torch.empty(nlines, dtype=torch.int32, device=buf.device)
        _check(self.lib.L.sre_cuda_thompson_exec_lines(self.cp, buf.data_ptr(), nlines, pitch, linelen,
                                                       rc.data_ptr(), engine, _stream_ptr()))
        return rc

    def thompson_ragged(self, buf: torch.Tensor, offsets: torch.Tensor, engine: int = ENGINE_AUTO):
        assert buf.is_cuda and offsets.is_cuda and offsets.dtype == torch.int64
        n = offsets.numel() - 1
        rc = torch.empty(n, dtype=torch.int32, device=buf.device)
        _check(self.lib.L.sre_cuda_thompson_exec_ragged(self.cp, buf.data_ptr(), offsets.data_ptr(), n,
                                                        rc.data_ptr(), engine, _stream_ptr()))
        return rc

    def pike_lines(self, buf: torch.Tensor, nlines: int, pitch: int, linelen: int,
                   offsets: torch.Tensor | None = None, select: torch.Tensor | None = None,
                   out_rc: torch.Tensor | None = None, out_ovec: torch.Tensor | None = None):
        n = self.nslots
        if offsets is not None and linelen == 0 and nlines > 0:
            # an upper bound on the line lengths lets the kernels pack capture offsets
            linelen = int((offsets[1:nlines + 1] - offsets[:nlines]).max())
        rc = out_rc if out_rc is not None else torch.empty(nlines, dtype=torch.int32, device=buf.device)
        ov = out_ovec if out_ovec is not None else torch.empty((nlines, n), dtype=torch.int64,
                                                               device=buf.device)
        _check(self.lib.L.sre_cuda_pike_exec_lines(
            self.cp, buf.data_ptr(), offsets.data_ptr() if offsets is not None else None, nlines, pitch,
            linelen, select.data_ptr() if select is not None else None, rc.data_ptr(), ov.data_ptr(), n,
            _stream_ptr()))
        return rc, ov

    def pike_lines_all(self, buf: torch.Tensor, nlines: int, pitch: int, linelen: int, max_matches: int,
                       offsets: torch.Tensor | None = None):
        """-> (count int32[n], spans int64[n, max_matches, 2], ids int32[n, max_matches])"""
        count = torch.zeros(nlines, dtype=torch.int32, device=buf.device)
        spans = torch.full((nlines, max_matches, 2), -1, dtype=torch.int64, device=buf.device)
        ids = torch.full((nlines, max_matches), -1, dtype=torch.int32, device=buf.device)
        _check(self.lib.L.sre_cuda_pike_exec_lines_all(
            self.cp, buf.data_ptr(), offsets.data_ptr() if offsets is not None else None, nlines, pitch,
            linelen, max_matches, count.data_ptr(), spans.data_ptr(), ids.data_ptr(), _stream_ptr()))
        return count, spans, ids

    def thompson_stream(self, buf: torch.Tensor, length: int, chunk_bytes: int, eof: bool,
                        state: int = STATE_INIT):
        """-> (rc, new_state, match_chunk)"""
        st, mc = C.c_uint32(state), C.c_int64(-1)
        rc = self.lib.L.sre_cuda_thompson_exec_stream(self.cp, buf.data_ptr(), length, chunk_bytes,
                                                      int(eof), C.byref(st), C.byref(mc), _stream_ptr())
        if rc == capi.SRE_ERROR:
            raise SreCudaError(self.lib.L.sre_cuda_last_error().decode())
        return rc, st.value, mc.value

    def stream_reduce(self, buf: torch.Tensor, length: int, halo: torch.Tensor | None = None,
                      entry_state: int = STATE_UNKNOWN) -> "StreamScan":
        """one part of a sharded stream -> scan handle with .fn, the part's record"""
        fn = C.create_string_buffer(STREAM_FN_BYTES)
        if halo is not None:
            assert halo.is_cuda and halo.dtype == torch.uint8 and halo.numel() == STREAM_HALO
        h = self.lib.L.sre_cuda_thompson_stream_reduce(
            self.cp, buf.data_ptr(), length, halo.data_ptr() if halo is not None else None, entry_state, fn,
            _stream_ptr())
        if not h:
            raise SreCudaError(self.lib.L.sre_cuda_last_error().decode())
        return StreamScan(self, h, fn.raw, (buf, halo))

    def thompson_text(self, buf: torch.Tensor, length: int | None = None, max_lines: int | None = None,
                      want_offsets: bool = True):
        """grep: the verdict of every line of a '\\n'-delimited device buffer in one pass
        -> (rc int32[nlines], offsets int64[nlines + 1] or None)"""
        length = buf.numel() if length is None else length
        cap = max_lines if max_lines is not None else max(1024, length // 48)
        while True:
            rc = torch.empty(cap, dtype=torch.int32, device=buf.device)
            off = torch.empty(cap + 1, dtype=torch.int64, device=buf.device) if want_offsets else None
            n = C.c_size_t(0)
            _check(self.lib.L.sre_cuda_thompson_exec_text(self.cp, buf.data_ptr(), length,
                                                          off.data_ptr() if off is not None else None,
                                                          rc.data_ptr(), cap, C.byref(n), _stream_ptr()))
            if n.value <= cap or max_lines is not None:
                k = min(n.value, cap)
                return rc[:k], (off[: k + 1] if off is not None else None)
            cap = n.value

    def dfa_fin(self, state: int) -> bool:
        """does the EOF step of the lowered DFA see a match in `state`"""
        return bool(self.lib.L.sre_cuda_dfa_fin(self.cp, state))

    def set_pike_tier(self, mode: int):
        self.lib.L.sre_cuda_program_set_pike_tier(self.cp, mode)

    def last_pike_tier(self) -> int:
        return self.lib.L.sre_cuda_program_last_pike_tier(self.cp)

    # host-buffer (end-to-end) forms: H2D + kernels + D2H inside the call
    def thompson_lines_host(self, host_buf: torch.Tensor, nlines, pitch, linelen, host_rc: torch.Tensor,
                            engine: int = ENGINE_AUTO):
        assert not host_buf.is_cuda and not host_rc.is_cuda
        _check(self.lib.L.sre_cuda_thompson_exec_lines_host(self.cp, host_buf.data_ptr(), nlines, pitch,
                                                            linelen, host_rc.data_ptr(), engine))
        return host_rc

    def thompson_stream_host(self, host_buf: torch.Tensor, length: int, chunk_bytes: int, eof: bool,
                             state: int = STATE_INIT, slice_bytes: int = 0):
        """a stream in host memory -> (rc, new_state, match_chunk)"""
        assert not host_buf.is_cuda
        st, mc = C.c_uint32(state), C.c_int64(-1)
        rc = self.lib.L.sre_cuda_thompson_exec_stream_host(self.cp, host_buf.data_ptr(), length, chunk_bytes,
                                                           int(eof), C.byref(st), C.byref(mc), slice_bytes)
        if rc == capi.SRE_ERROR:
            raise SreCudaError(self.lib.L.sre_cuda_last_error().decode())
        return rc, st.value, mc.value

    def pike_lines_host(self, host_buf, nlines, pitch, linelen, host_rc, host_ovec, gate=True):
        _check(self.lib.L.sre_cuda_pike_exec_lines_host(self.cp, host_buf.data_ptr(), nlines, pitch,
                                                        linelen, int(gate), host_rc.data_ptr(),
                                                        host_ovec.data_ptr(), self.nslots))
        return host_rc, host_ovec


class PikeStreams:
    """nstreams persistent Pike contexts on the device (sre_cuda_pike_streams_*): each exec() call
    feeds every stream its next chunk, like sre_vm_pike_exec on one ctx per connection"""

    def __init__(self, prog: "CudaProgram", nstreams: int):
        self.prog, self.n = prog, nstreams
        self.handle = prog.lib.L.sre_cuda_pike_streams_create(prog.cp, nstreams, _stream_ptr())
        if not self.handle:
            raise SreCudaError(prog.lib.L.sre_cuda_last_error().decode())

    def exec(self, buf: torch.Tensor, offsets: torch.Tensor, eof=None, eof_all: bool = False) -> torch.Tensor:
        """-> int64[nstreams, 4 + nslots]: rc, pending flag, pending span, ovector"""
        out = torch.full((self.n, 4 + self.prog.nslots), -99, dtype=torch.int64, device=buf.device)
        _check(self.prog.lib.L.sre_cuda_pike_streams_exec(
            self.handle, buf.data_ptr(), offsets.data_ptr(), eof.data_ptr() if eof is not None else None,
            int(eof_all), out.data_ptr(), self.prog.nslots, _stream_ptr()))
        return out

    def close(self):
        if self.handle:
            self.prog.lib.L.sre_cuda_pike_streams_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class StreamScan:
    """a reduced stream part (sre_cuda_thompson_stream_reduce): its records stay
    on the device until resolve() is given the part's true entry state"""

    def __init__(self, prog, handle, fn: bytes, keep):
        self.prog, self.handle, self.fn, self._keep = prog, handle, fn, keep

    def resolve(self, entry_state: int):
        """-> (exit state, offset of the first step that sees a match or -1); .fn is updated"""
        ex, off = C.c_uint32(0), C.c_int64(-1)
        fn = C.create_string_buffer(STREAM_FN_BYTES)
        _check(self.prog.lib.L.sre_cuda_thompson_stream_resolve(self.handle, entry_state, C.byref(ex),
                                                                C.byref(off), fn))
        self.fn = fn.raw
        return ex.value, off.value

    def close(self):
        if self.handle:
            self.prog.lib.L.sre_cuda_thompson_stream_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fn_apply(fn: bytes, state: int) -> int:
    """state a part's record leads to from `state`; STATE_UNKNOWN when the record cannot tell"""
    return lib().L.sre_cuda_stream_fn_apply(fn, state)


def index_lines(buf: torch.Tensor, length: int | None = None, max_lines: int | None = None) -> torch.Tensor:
    """Offsets (int64, on the device) of the '\\n'-delimited lines of a device byte
    buffer: line i = buf[off[i], off[i+1]), terminator included; feed them to
    thompson_ragged / pike_lines(offsets=...)."""
    length = buf.numel() if length is None else length
    cap = max_lines if max_lines is not None else max(1024, length // 64)
    while True:
        off = torch.empty(cap + 1, dtype=torch.int64, device=buf.device)
        n = C.c_size_t(0)
        _check(lib().L.sre_cuda_index_lines(buf.data_ptr(), length, off.data_ptr(), cap, C.byref(n), _stream_ptr()))
        if n.value <= cap or max_lines is not None:
            return off[: min(n.value, cap) + 1]
        cap = n.value


def launch_count(reset=False) -> int:
    return lib().L.sre_cuda_launch_count(int(reset))
