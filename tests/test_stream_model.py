"""CPU tier: the image automaton (sregex_b200/csrc/lower/sre_image.cpp) and the
candidate-based stream scan built on it, as a host model
(oracle/lower_check.cpp: lc_stream_model), against the oracle's
sre_vm_thompson_exec fed the same stream in chunks.  What is checked is the
claim the stream kernels rest on: the state carried into a piece is always among
the candidates the image automaton names for it."""
import ctypes as C
import os
import random

import pytest

from sregex_b200 import capi, corpus

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lc():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "liblowercheck.so"))
    lib.lc_create.restype = C.c_void_p
    lib.lc_create.argtypes = [C.c_void_p, C.c_uint]
    lib.lc_destroy.argtypes = [C.c_void_p]
    lib.lc_info.argtypes = [C.c_void_p, C.POINTER(C.c_uint)]
    lib.lc_dfa_fin.argtypes = [C.c_void_p, C.c_uint]
    lib.lc_stream_model.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint,
                                    C.c_uint, C.POINTER(C.c_longlong)]
    lib.lc_image_info.argtypes = [C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]
    return lib


def _oracle_chunked(o, p, data, chunk):
    """rc sequence of the oracle's Thompson VM fed `chunk`-byte chunks -> (last rc, index of that call)"""
    pieces = [(data[i:i + chunk], i + chunk >= len(data)) for i in range(0, len(data), chunk)] or [(b"", True)]
    rcs = o.thompson(p, data, pieces)
    return rcs[-1], len(rcs) - 1


def _model(lc, h, data, piece, window, K=8):
    out = (C.c_longlong * 4)()
    rc = lc.lc_stream_model(h, data, len(data), piece, window, K, 0, out)
    assert rc == 0, "carried state not among a piece's candidates"
    return list(out)


def _verdict(lc, h, exit_state):
    return capi.SRE_OK if exit_state == 1 or lc.lc_dfa_fin(h, exit_state) else capi.SRE_DECLINED


def test_image_automaton_sizes(oracle, lc):
    """the automata the bench uses stay small, and most of their states are narrow"""
    for rx, limit in ((corpus.BENCH_REGEX, 200), (corpus.C2_REGEX, 200), (corpus.C3_REGEX, 200),
                      (corpus.multi_pattern_set(64), 6000)):
        p = oracle.compile(rx)
        h = lc.lc_create(p.prog, 16384)
        info = (C.c_uint * 2)()
        assert lc.lc_image_info(h, 8, info) == 0
        assert 2 <= info[0] <= limit and info[1] >= info[0] // 2, (rx, list(info))
        lc.lc_destroy(h)
        p.close()


@pytest.mark.parametrize("name", ["bench", "c2", "c3", "multi8", "multi64"])
def test_stream_model_on_bench_programs(oracle, lc, name):
    from test_gpu_parity import MULTI
    rx = {"bench": corpus.BENCH_REGEX, "c2": corpus.C2_REGEX, "c3": corpus.C3_REGEX, "multi8": MULTI,
          "multi64": corpus.multi_pattern_set(64)}[name]
    p = oracle.compile(rx)
    h = lc.lc_create(p.prog, 16384)
    texts = [bytes(corpus.gen_data_buffer(3000).numpy()),
             bytes(corpus.log_lines(24, 1024).numpy().tobytes()),
             bytes(corpus.log_lines(24, 1024, hit_rate=0.0).numpy().tobytes()).replace(b"GET", b"get")
                 .replace(b"POST", b"post").replace(b"PUT", b"put").replace(b"HEAD", b"head")]
    for data in texts:
        want_rc = oracle.thompson(p, data)
        for piece, window in ((4096, 64), (256, 32), (64, 16)):
            ex, first, unres, maxc = _model(lc, h, data, piece, window)
            assert _verdict(lc, h, ex) == want_rc
            if first >= 0:
                # the step that enters ACC is the one in which the reference's loop sees the
                # MATCH thread: its chunked call sequence returns SRE_OK in that chunk
                for chunk in (1000, 4096):
                    rc, idx = _oracle_chunked(oracle, p, data, chunk)
                    assert rc == capi.SRE_OK and idx == first // chunk, (name, chunk, idx, first)
            assert unres <= 2 and maxc <= 8
    lc.lc_destroy(h)
    p.close()


def test_stream_model_random_regex_fuzz(oracle, lc):
    """random regexes (assertions included) over random text in small pieces, so
    that many pieces start inside partial matches"""
    rng = random.Random(77)
    atoms = ["a", "b", "ab", " ", "\\n", "_", ".", "^", "$", "\\b", "\\B", "\\A", "\\z", "|", "(", ")", "(?:", "*",
             "+", "?", "*?", "+?", "??", "{2}", "{0,2}", "{1,}", "[ab]", "[^a]", "\\w", "\\W", "\\s", "\\d", "1"]
    alphabet = b"ab \n_1."
    done = unresolved = pieces = 0
    while done < 300:
        rx = "".join(rng.choice(atoms) for _ in range(rng.randrange(1, 9))).encode()
        try:
            p = oracle.compile(rx, 0)
        except capi.SreSyntaxError:
            continue
        h = lc.lc_create(p.prog, 16384)
        info = (C.c_uint * 6)()
        lc.lc_info(h, info)
        if not info[3]:
            lc.lc_destroy(h)
            p.close()
            continue
        done += 1
        n = rng.choice([0, 1, 15, 64, 300, 1000])
        # mostly text that does not match early: drop the rarer letters now and then
        alpha = alphabet if rng.random() < 0.5 else bytes(rng.sample(list(alphabet), 4))
        data = bytes(rng.choice(alpha) for _ in range(n))
        want_rc = oracle.thompson(p, data)
        for piece, window in ((32, 8), (16, 16), (100, 4)):
            ex, first, unres, maxc = _model(lc, h, data, piece, window)
            assert _verdict(lc, h, ex) == want_rc, (rx, data, piece)
            unresolved += unres
            pieces += max(1, (n + piece - 1) // piece)
            if first >= 0:
                rc, idx = _oracle_chunked(oracle, p, data, 7)
                assert rc == capi.SRE_OK and idx == first // 7, (rx, data, idx, first)
        lc.lc_destroy(h)
        p.close()
    assert unresolved < pieces // 4
