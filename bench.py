#!/usr/bin/env python3
"""bench.py -- benchmark of the sregex match-execution hot path on B200.

BASELINE.json's metric is "input GB/s scanned (Thompson bool, Pike+captures) at
1/2/4/8 B200 vs host-CPU sregex".  The JSON line's headline (`value`, `roofline`,
`e2e`, `cpu_baseline`) is BASELINE configs[1] (C2); `extra` carries the other
GPU configs, each measured the same way (device-timed value, roofline, e2e from
host buffers, the reference's CPU path beside it), at every N:

  c2  Thompson boolean, one regex, 1,048,576 x 1 KB log lines per GPU      (weak)
  c3  Pike VM, 4 capture groups, the same lines, rc + ovector per line     (weak)
  c4  64-pattern sre_regex_parse_multi set over an 8 GiB corpus sharded
      across the GPUs: which pattern matched (Thompson gate + Pike id)     (strong)
  c5  one 32 GiB stream in 64 KB chunks with SRE_AGAIN state carry,
      chunk-parallel scan; sharded: halo + record all-gather over NCCL     (strong)
  text  Thompson boolean per line of newline-delimited text (ragged lines,
      no line index beforehand), 1 GiB per GPU, one pass                   (weak)
  nfa   a regex whose subset construction blows up: the bit-parallel NFA tier
      over the C2 lines                                                    (weak)

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference ...                      (reference CPU arm)
  python bench.py --config c3 ...                           (one config as the headline)

A "step" is one pass of the hot path over the config's batch.  `value` = input
GB/s with the data resident in HBM (CUDA events around the K steps, max over
ranks); `e2e` = the same through the host-buffer C-ABI call (H2D of the input
and D2H of the results inside the timed region); `roofline` = algorithmic bytes
/ event-timed duration vs the measured HBM peak.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NLINES = 1 << 20
PITCH = 1024
C4_TOTAL_LINES = 8 << 20            # 8 GiB of 1 KB lines
C5_TOTAL_BYTES = 32 << 30
CHUNK = 65536
REGEX_NAME = 'HTTP/1\\.[01]" 5\\d\\d '
METRICS = {
    "c2": "input GB/s scanned (Thompson boolean, 1 regex, 1M x 1 KB lines per GPU)",
    "c3": "input GB/s scanned (Pike + 4 capture groups, rc + ovector per line, 1M x 1 KB lines per GPU)",
    "c4": "input GB/s scanned (64-pattern set, matched id per line, 8 GiB corpus sharded over the GPUs)",
    "c5": "input GB/s scanned (one 32 GiB stream, 64 KB chunks with SRE_AGAIN carry, sharded over the GPUs)",
    "text": "input GB/s scanned (Thompson boolean per line of newline-delimited text, 1 GiB per GPU, one pass)",
    "nfa": "input GB/s scanned (Thompson boolean, a regex with no DFA: bit-parallel NFA tier, 1M x 1 KB lines per GPU)",
}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed regions (NVML, 4 ms period).
    Started before the warm-up steps -- the first NVML queries take the driver
    for milliseconds and would stall kernel launches inside a short timed region
    -- and told with mark() where a timed region begins."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        self.first = 0
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def mark(self):
        """a timed region starts now: earlier samples (but the last one) and
        earlier throttle reasons do not count"""
        self.first = max(0, len(self.samples) - 1)
        self.reasons = set()

    def result(self):
        s = sorted(self.samples[self.first:])
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}

    def cpu_affinity(self):
        """CPUs near this GPU (NVML), for the pinned host buffers of the e2e legs"""
        try:
            words = self.nv.nvmlDeviceGetCpuAffinity(self.h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
            return cpus & set(os.sched_getaffinity(0))
        except Exception:
            return set()


# ---- the reference's CPU path --------------------------------------------------------

def ref_kind():
    from oracle import cpu_baseline as baseline
    return "ref" if baseline.available("ref") else "oracle"


def cpu_c2(lines_host, cores, reps=1):
    """Thompson JIT (+ interpreter) over 1 KB lines, all cores -> dict"""
    from oracle import cpu_baseline as baseline
    from sregex_b200 import corpus
    which = ref_kind()
    n = lines_host.shape[0]
    out = {}
    for name, eng in (("jit", baseline.ENGINE_JIT), ("interp", baseline.ENGINE_THOMPSON)):
        if which == "oracle" and name == "jit":
            continue
        secs, rc, _ = baseline.run_lines(which, corpus.C2_REGEX, None, lines_host, n, PITCH, PITCH, eng,
                                         nthreads=cores, reps=reps)
        out[name] = n * PITCH * reps / secs / 1e9
        out["hits"] = int((rc == 0).sum())
    return out


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on the host cores: the
    Thompson JIT over the full C2 batch per step (threads, parse, compile and JIT
    outside the timed region), and the other configs' CPU paths on bounded samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as baseline
    from sregex_b200 import corpus
    cores = host_cores()
    which = ref_kind()
    engine = baseline.ENGINE_JIT if which == "ref" else baseline.ENGINE_THOMPSON
    n = args.lines
    lines = corpus.log_lines(n, PITCH).numpy()
    # bounded: at ~3 GB/s a 1 GiB step takes 0.35 s; keep the whole run within ~2 minutes
    steps = max(1, min(args.steps, 240))
    if args.warmup:
        baseline.run_lines(which, corpus.C2_REGEX, None, lines, n, PITCH, PITCH, engine, nthreads=cores,
                           reps=min(args.warmup, 3))
    secs, rc, _ = baseline.run_lines(which, corpus.C2_REGEX, None, lines, n, PITCH, PITCH, engine,
                                     nthreads=cores, reps=steps)
    value = n * PITCH * steps / secs / 1e9
    what = "sre_vm_thompson_jit handler" if which == "ref" else "oracle port of sre_vm_thompson_exec"
    sample = (f"all {n} lines per step ({n * PITCH >> 20} MiB), {what}, one OS thread per core with a private "
              f"program, fresh ctx+pool per line; thread start, parse, compile and JIT outside the timed region")
    out = {
        "impl": "reference", "metric": METRICS["c2"], "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": secs / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "C2: Thompson boolean, regex " + REGEX_NAME + ", 1,048,576 x 1 KB log lines per GPU",
                   "lines_per_step": n, "line_bytes": PITCH},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores,
                         "kind": "reference" if which == "ref" else "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "hits": int((rc == 0).sum()),
    }
    if not args.no_extras:
        out["extra"] = cpu_extras(lines, cores)
    print(json.dumps(out))


def cpu_extras(lines, cores):
    """the reference's CPU paths of C3 / C4 / C5 on bounded samples (rates)"""
    from oracle import cpu_baseline as baseline
    from sregex_b200 import corpus
    which = ref_kind()
    ex = {}
    try:
        ns = min(lines.shape[0], 2048 * cores)
        secs, rc, _ = baseline.run_lines(which, corpus.C3_REGEX, None, lines[:ns], ns, PITCH, PITCH,
                                         baseline.ENGINE_PIKE, nthreads=cores, ovec_slots=10)
        ex["c3"] = {"value": ns * PITCH / secs / 1e9, "unit": "GB/s", "cores": cores,
                    "sample": f"sre_vm_pike_exec, first {ns} lines", "matched": int((rc >= 0).sum())}
        ns4 = min(lines.shape[0], 128 * cores)
        secs, rc, _ = baseline.run_lines(which, corpus.multi_pattern_set(64), None, lines[:ns4], ns4, PITCH, PITCH,
                                         baseline.ENGINE_PIKE, nthreads=cores, ovec_slots=2)
        ex["c4"] = {"value": ns4 * PITCH / secs / 1e9, "unit": "GB/s", "cores": cores,
                    "sample": f"sre_vm_pike_exec with the 64-pattern set, first {ns4} lines",
                    "matched": int((rc >= 0).sum())}
        sbytes = 64 << 20
        stream = corpus.gen_data_buffer(sbytes // 5).numpy()
        eng = baseline.ENGINE_JIT if which == "ref" else baseline.ENGINE_THOMPSON
        secs, last_rc, calls = baseline.run_stream(which, corpus.BENCH_REGEX, None, stream, CHUNK, eng)
        ex["c5"] = {"value": stream.size / secs / 1e9, "unit": "GB/s", "cores": 1,
                    "sample": f"{'Thompson JIT' if which == 'ref' else 'oracle port'}, one ctx fed "
                              f"{stream.size >> 20} MiB of the stream in 64 KB chunks with SRE_AGAIN carry "
                              f"(a single stream is sequential on the CPU: one core)",
                    "last_rc": last_rc, "calls": calls}
    except Exception as e:
        ex["error"] = repr(e)
    return ex


# ---- our arm ---------------------------------------------------------------------------

class Harness:
    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.sampler = ClockSampler(self.local)
        self.sampler.start()
        self.peak, self.peak_src = measured_hbm_peak()
        # pinned host buffers on the NUMA node next to this rank's GPU
        near = self.sampler.cpu_affinity() if self.sampler.nv is not None else set()
        self.numa_note = "no NUMA control"
        if near and self.world > 1:
            try:
                os.sched_setaffinity(0, near)
                self.numa_note = f"rank pinned to the {len(near)} CPUs next to its GPU (first-touch pinned buffers)"
            except Exception:
                pass

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.int64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t)
        return int(t.item())

    def timed(self, step, steps, warmup, closing=None):
        """W warm-up steps, then exactly K steps between barriers: (total ms, max over
        ranks; mean per-step ms on this rank; clocks during the region; library launches)"""
        from sregex_b200 import cuda
        torch = self.torch
        for _ in range(max(warmup, 3)):
            step()
        if closing:
            closing()
        self.barrier()
        while self.sampler.nv is not None and len(self.sampler.samples) < 3:
            time.sleep(0.002)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cuda.launch_count(reset=True)
        self.barrier()
        self.sampler.mark()
        t0.record()
        for a, b in evs:
            a.record()
            step()
            b.record()
        if closing:
            closing()
        t1.record()
        self.barrier()
        launches = cuda.launch_count()
        clocks = self.sampler.result()
        total = self.max_over_ranks(t0.elapsed_time(t1))
        per_step = sum(a.elapsed_time(b) for a, b in evs) / steps
        return total, per_step, clocks, launches

    def timed_host(self, step, steps):
        """wall clock around host-buffer calls (each ends with its own device sync), max over ranks"""
        step()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        self.torch.cuda.synchronize()
        return self.max_over_ranks(time.perf_counter() - t0)

    def roofline(self, algo_bytes, ms, kernel, traffic=None, note=None):
        achieved = algo_bytes / (ms * 1e-3) / 1e9
        r = {"bound": "hbm", "achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak,
             "traffic": traffic, "peak_source": self.peak_src, "kernel": kernel, "kernel_ms": ms,
             "algorithmic_bytes": algo_bytes}
        if note:
            r["note"] = note
        return r

    def log_corpus(self, nlines, first_line):
        from sregex_b200 import corpus
        torch = self.torch
        dev = torch.empty((nlines, PITCH), dtype=torch.uint8, device="cuda")
        blk = 1 << 17
        for i in range(0, nlines, blk):
            m = min(blk, nlines - i)
            dev[i:i + m] = corpus.log_lines(m, PITCH, device="cuda", first_line=first_line + i)
        return dev

    def pinned_copy(self, dev):
        host = self.torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
        host.copy_(dev)
        return host


def traffic_of(kernel):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(p))
        return t.get(kernel, {}).get("bytes_per_launch") if isinstance(t.get(kernel), dict) else (
            t.get("bytes_per_launch") if kernel == "k_dfa_lines_skipw" else None)
    except Exception:
        return None


def bench_c2(h, steps, warmup):
    from sregex_b200 import corpus, cuda
    torch, args = h.torch, h.args
    n = args.lines
    # corpus shard of this rank, resident in HBM (1 GiB > 126 MB of L2, so every
    # step streams it from HBM again)
    dev = h.log_corpus(n, h.rank * n)
    prog = cuda.CudaProgram(corpus.C2_REGEX)
    rc = torch.empty(n, dtype=torch.int32, device="cuda")
    info = prog.info
    engine = {"auto": cuda.ENGINE_AUTO, "tiled": cuda.ENGINE_DFA_TILED, "skip": cuda.ENGINE_DFA_SKIP,
              "generic": cuda.ENGINE_DFA_GENERIC, "nfa": cuda.ENGINE_NFA}[args.engine]
    engine_name = args.engine
    if engine_name == "auto":
        engine_name = ("nfa" if not info.dfa_states else
                       "dfa_skip" if 1 <= info.dfa_leave_bytes <= 2 else "dfa_tiled")
    engine = cuda.engine_variant(engine, args.variant)
    hits = [None]

    def step():
        prog.thompson_lines(dev, n, PITCH, PITCH, engine=engine, out=rc)

    def closing():
        # the closing match count (and, at N > 1, its NCCL reduce) is part of the job
        t = (rc == 0).sum()
        if h.world > 1:
            h.dist.all_reduce(t)
        hits[0] = t

    total_ms, kern_ms, clocks, launches = h.timed(step, steps, warmup, closing)
    bytes_per_step = n * PITCH
    kernel = {"dfa_skip": "k_dfa_lines_skipw", "dfa_tiled": "k_dfa_lines_tma_early"}.get(engine_name, engine_name)
    out = {
        "metric": METRICS["c2"], "value": h.world * bytes_per_step * steps / (total_ms * 1e-3) / 1e9,
        "unit": "GB/s", "n_gpus": h.world, "steps": steps, "warmup": max(warmup, 3),
        "ms_per_step": total_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": "C2: Thompson boolean, regex " + REGEX_NAME + ", 1,048,576 x 1 KB log lines per GPU",
            "lines_per_gpu": n, "line_bytes": PITCH,
            "sharding": f"lines x {h.world} ranks, no data-path collective (one all_reduce of the hit count)",
            "engine": engine_name, "dfa_states": info.dfa_states, "nfa_states": info.nfa_states,
            "variant": args.variant,
            "l2_policy": "input (1 GiB per GPU) larger than L2 (126 MB); no flush needed",
        },
        # algorithmic bytes per launch = 1 B read per input byte + 4 B verdict per line
        "roofline": h.roofline(bytes_per_step + 4 * n, kern_ms, kernel, traffic_of(kernel)),
        "gpu_launches": launches, "clocks": clocks, "hits": int(hits[0].item()),
    }
    # e2e through the host-buffer C-ABI call (every rank, its own shard)
    e2e_steps = max(2, min(5, steps))
    host = h.pinned_copy(dev)
    host_rc = torch.empty(n, dtype=torch.int32, pin_memory=True)
    dt = h.timed_host(lambda: prog.thompson_lines_host(host, n, PITCH, PITCH, host_rc), e2e_steps)
    assert torch.equal(host_rc, rc.cpu()), "host-buffer path disagrees with the device-resident path"
    out["e2e"] = {"value": h.world * bytes_per_step * e2e_steps / dt / 1e9, "unit": "GB/s",
                  "h2d_bytes_per_step": bytes_per_step, "d2h_bytes_per_step": 4 * n, "steps": e2e_steps,
                  "call": "sre_cuda_thompson_exec_lines_host",
                  "note": "pinned host buffers, all ranks concurrently, max over ranks; " + h.numa_note}
    if h.rank == 0 and h.world == 1:
        cores = host_cores()
        which = ref_kind()
        ns = min(n, 16384 * cores)
        cb = cpu_c2(host[:ns].numpy(), cores)
        out["cpu_baseline"] = {
            "value": cb.get("jit", cb.get("interp")), "unit": "GB/s", "cores": cores,
            "kind": "reference" if which == "ref" else "port",
            "sample": f"first {ns} lines ({ns * PITCH >> 20} MiB) of the workload, one OS thread per core, "
                      f"private program per thread, fresh ctx+pool per line; set-up outside the timed region",
            "thompson_jit_gbs": cb.get("jit"), "thompson_interp_gbs": cb.get("interp"),
        }
        assert cb["hits"] == int((host_rc[:ns] == 0).sum()), "CPU baseline disagrees with GPU verdicts"
    return out, dev, host


def bench_c3(h, steps, warmup, dev, host):
    """Pike + 4 capture groups over the C2 lines: rc + 10 ovector slots per line"""
    from oracle import cpu_baseline as baseline
    from sregex_b200 import corpus, cuda
    torch = h.torch
    n = dev.shape[0]
    prog = cuda.CudaProgram(corpus.C3_REGEX)
    ns = prog.nslots
    prc = torch.empty(n, dtype=torch.int32, device="cuda")
    pov = torch.empty((n, ns), dtype=torch.int64, device="cuda")
    total_ms, call_ms, clocks, launches = h.timed(
        lambda: prog.pike_lines(dev, n, PITCH, PITCH, out_rc=prc, out_ovec=pov), steps, warmup)
    matched = int((prc >= 0).sum())
    bytes_per_step = n * PITCH
    out = {
        "metric": METRICS["c3"], "value": h.world * bytes_per_step * steps / (total_ms * 1e-3) / 1e9, "unit": "GB/s",
        "steps": steps, "ms_per_step": total_ms / steps, "scaling": "weak",
        "config": {"workload": "C3: sre_vm_pike_exec per line, regex (\\w+) (\\S+) HTTP/(\\d)\\.(\\d), "
                               "1,048,576 x 1 KB log lines per GPU",
                   "lines_per_gpu": n, "ovector_slots": ns, "matched_lines": matched,
                   "pike_tier": prog.last_pike_tier()},
        # 1 B per input byte (the gate reads every line once) + 8*(1+slots) B per matched line
        "roofline": h.roofline(bytes_per_step + 8 * (1 + ns) * matched, call_ms,
                               "sre_cuda_pike_exec_lines: k_dfa_lines_hint (packs the matching lines) + Pike kernel(s)",
                               note="whole call (3-4 launches); per-kernel shares in profiles/"),
        "gpu_launches": launches, "clocks": clocks,
    }
    e2e_steps = max(2, min(3, steps))
    host_rc = torch.empty(n, dtype=torch.int32, pin_memory=True)
    host_ov = torch.empty((n, ns), dtype=torch.int64, pin_memory=True)
    dt = h.timed_host(lambda: prog.pike_lines_host(host, n, PITCH, PITCH, host_rc, host_ov, gate=False), e2e_steps)
    assert torch.equal(host_rc, prc.cpu()) and torch.equal(host_ov, pov.cpu()), "C3: host path differs"
    out["e2e"] = {"value": h.world * bytes_per_step * e2e_steps / dt / 1e9, "unit": "GB/s",
                  "h2d_bytes_per_step": bytes_per_step, "d2h_bytes_per_step": n * (4 + 8 * ns), "steps": e2e_steps,
                  "call": "sre_cuda_pike_exec_lines_host"}
    if h.rank == 0 and h.world == 1:
        cores = host_cores()
        which = ref_kind()
        m = min(n, 2048 * cores)
        secs, crc, cov = baseline.run_lines(which, corpus.C3_REGEX, None, host[:m].numpy(), m, PITCH, PITCH,
                                            baseline.ENGINE_PIKE, nthreads=cores, ovec_slots=ns)
        assert (host_rc[:m].numpy() == crc).all() and (host_ov[:m].numpy() == cov).all(), \
            "C3: GPU rows differ from the CPU reference"
        out["cpu_baseline"] = {"value": m * PITCH / secs / 1e9, "unit": "GB/s", "cores": cores,
                               "kind": "reference" if which == "ref" else "port",
                               "sample": f"sre_vm_pike_exec over the first {m} lines; rc and all {ns} ovector "
                                         f"slots compared with the GPU rows one for one"}
    return out


def bench_c4(h, steps, warmup):
    """64 patterns, 8 GiB corpus strong-scaled: Thompson gate + matched id (Pike) per line"""
    from oracle import cpu_baseline as baseline
    from sregex_b200 import corpus, cuda
    from sregex_b200 import dist as sdist
    torch = h.torch
    total_lines = h.args.c4_lines
    first, n = sdist.shard_range(total_lines, h.rank, h.world)
    dev = h.log_corpus(n, first)
    pats = corpus.multi_pattern_set(64)
    prog = cuda.CudaProgram(pats)
    ns = prog.nslots
    gate = torch.empty(n, dtype=torch.int32, device="cuda")
    mrc = torch.empty(n, dtype=torch.int32, device="cuda")
    mov = torch.empty((n, ns), dtype=torch.int64, device="cuda")
    g_ms, g_call, _, _ = h.timed(lambda: prog.thompson_lines(dev, n, PITCH, PITCH, out=gate), steps, warmup)
    total_ms, call_ms, clocks, launches = h.timed(
        lambda: prog.pike_lines(dev, n, PITCH, PITCH, out_rc=mrc, out_ovec=mov), steps, warmup)
    assert torch.equal(gate == 0, mrc >= 0), "C4: gate and matched ids disagree"
    matched = h.sum_over_ranks(int((mrc >= 0).sum()))
    total_bytes = total_lines * PITCH
    out = {
        "metric": METRICS["c4"], "value": total_bytes * steps / (total_ms * 1e-3) / 1e9, "unit": "GB/s",
        "steps": steps, "ms_per_step": total_ms / steps, "scaling": "strong",
        "gate_only_gbs": total_bytes * steps / (g_ms * 1e-3) / 1e9,
        "config": {"workload": "C4: sre_regex_parse_multi set of 64 patterns, matched regex id + $0 per line "
                               "(Thompson gate, then Pike on the hits), 1 KB log lines",
                   "total_lines": total_lines, "lines_this_rank": n, "dfa_states": prog.info.dfa_states,
                   "matched_fraction": matched / total_lines, "pike_tier": prog.last_pike_tier(),
                   "sharding": f"contiguous line ranges over {h.world} ranks, no data-path collective"},
        "roofline": h.roofline(n * PITCH + 4 * n + 8 * (1 + ns) * int((mrc >= 0).sum()), call_ms,
                               "sre_cuda_pike_exec_lines: k_dfa_lines_big<hint> (2107-state DFA, table through L1; packs "
                               "the matching lines) + k_pike_lineage", note="whole call on this rank's shard"),
        "gate_roofline": h.roofline(n * PITCH + 4 * n, g_call, "k_dfa_lines_big (2107 states, TMA-tiled lines, table through L1)"),
        "gpu_launches": launches, "clocks": clocks,
    }
    e2e_steps = 2
    try:
        host = h.pinned_copy(dev)
        host_rc = torch.empty(n, dtype=torch.int32, pin_memory=True)
        host_ov = torch.empty((n, ns), dtype=torch.int64, pin_memory=True)
        dt = h.timed_host(lambda: prog.pike_lines_host(host, n, PITCH, PITCH, host_rc, host_ov, gate=False),
                          e2e_steps)
        assert torch.equal(host_rc, mrc.cpu()), "C4: host path differs"
        out["e2e"] = {"value": total_bytes * e2e_steps / dt / 1e9, "unit": "GB/s",
                      "h2d_bytes_per_step": n * PITCH, "d2h_bytes_per_step": n * (4 + 8 * ns), "steps": e2e_steps,
                      "call": "sre_cuda_pike_exec_lines_host"}
        if h.rank == 0 and h.world == 1:
            cores = host_cores()
            which = ref_kind()
            m = min(n, 128 * cores)
            secs, crc, cov = baseline.run_lines(which, pats, None, host[:m].numpy(), m, PITCH, PITCH,
                                                baseline.ENGINE_PIKE, nthreads=cores, ovec_slots=ns)
            assert (host_rc[:m].numpy() == crc).all() and (host_ov[:m].numpy() == cov).all(), \
                "C4: GPU matched ids / ovectors differ from the CPU reference"
            out["cpu_baseline"] = {"value": m * PITCH / secs / 1e9, "unit": "GB/s", "cores": cores,
                                   "kind": "reference" if which == "ref" else "port",
                                   "sample": f"sre_vm_pike_exec with the 64-pattern set over the first {m} lines; "
                                             f"matched ids and ovectors compared with the GPU rows"}
    except Exception as e:
        out["e2e"] = {"error": repr(e)}
    return out


def bench_text(h, steps, warmup, dev):
    """grep: every line of '\\n'-delimited text (ragged lines, no index beforehand) in one pass"""
    from oracle import cpu_baseline as baseline
    from sregex_b200 import capi, corpus, cuda
    torch = h.torch
    flat = dev.view(-1).clone()
    n = flat.numel()
    # newlines at pseudo-random places: line lengths geometric with mean ~180 bytes
    blk = 1 << 27
    for i in range(0, n, blk):
        m = min(blk, n - i)
        pos = torch.arange(i, i + m, dtype=torch.int64, device="cuda")
        hsh = (pos * -7046029254386353131) ^ (pos >> 13)
        flat[i:i + m][((hsh >> 17) & 0xFFFF) % 180 == 0] = 10
        del pos, hsh
    prog = cuda.CudaProgram(corpus.C2_REGEX)
    nl = int((flat == 10).sum()) + (1 if int(flat[-1]) != 10 else 0)
    rc = torch.empty(nl, dtype=torch.int32, device="cuda")
    off = torch.empty(nl + 1, dtype=torch.int64, device="cuda")
    found = [0]

    def step():
        cnt = __import__("ctypes").c_size_t(0)
        r = prog.lib.L.sre_cuda_thompson_exec_text(prog.cp, flat.data_ptr(), n, off.data_ptr(), rc.data_ptr(), nl,
                                                   __import__("ctypes").byref(cnt),
                                                   torch.cuda.current_stream().cuda_stream)
        assert r == capi.SRE_OK
        found[0] = cnt.value

    off_ms, off_call_ms, _, _ = h.timed(step, steps, warmup)
    assert found[0] == nl, (found[0], nl)
    rc_all = rc.clone()

    def step_rc():
        # verdicts only (what C2 delivers per line): no offsets wanted
        cnt = __import__("ctypes").c_size_t(0)
        r = prog.lib.L.sre_cuda_thompson_exec_text(prog.cp, flat.data_ptr(), n, None, rc.data_ptr(), nl,
                                                   __import__("ctypes").byref(cnt),
                                                   torch.cuda.current_stream().cuda_stream)
        assert r == capi.SRE_OK
        found[0] = cnt.value

    total_ms, call_ms, clocks, launches = h.timed(step_rc, steps, warmup)
    assert found[0] == nl and torch.equal(rc, rc_all), "text: verdict-only rows differ from the rows with offsets"
    # parity at full size: the line index + ragged route must give the same rows
    off2 = cuda.index_lines(flat)
    assert torch.equal(off2, off)
    rc2 = prog.thompson_ragged(flat, off2)
    assert torch.equal(rc2, rc), "text: one-pass rows differ from index + ragged"
    out = {
        "metric": METRICS["text"], "value": h.world * n * steps / (total_ms * 1e-3) / 1e9, "unit": "GB/s",
        "steps": steps, "ms_per_step": total_ms / steps, "scaling": "weak",
        "with_line_offsets_gbs": h.world * n * steps / (off_ms * 1e-3) / 1e9,
        "config": {"workload": "newline-delimited log text (ragged lines, mean ~180 B, no index beforehand): "
                               "sre_vm_thompson_exec per line, regex " + REGEX_NAME,
                   "bytes_per_gpu": n, "lines": nl, "matched_lines": int((rc == 0).sum())},
        # 1 B per input byte + 4 B verdict per line (+ 8 B offset per line in the second form)
        "roofline": h.roofline(n + 4 * nl, call_ms, "sre_cuda_thompson_exec_text: k_text_verdicts + k_text_finish",
                               note="whole call, verdict per line (2 launches + the read-back of the line count)"),
        "with_line_offsets_roofline": h.roofline(n + 12 * nl, off_call_ms,
                                                 "sre_cuda_thompson_exec_text: k_text_pieces + sums/scan/write"),
        "gpu_launches": launches, "clocks": clocks,
    }
    if h.rank == 0 and h.world == 1:
        which = ref_kind()
        m = 20000
        host_off = off[: m + 1].cpu().numpy()
        text = flat[: int(host_off[-1])].cpu().numpy().tobytes()
        o = capi.load("oracle" if which == "oracle" else "ref")
        po = o.compile(corpus.C2_REGEX)
        t0 = time.perf_counter()
        want = [o.thompson(po, text[host_off[i]:host_off[i + 1]], jit=(which == "ref")) for i in range(m)]
        dt = time.perf_counter() - t0
        assert want == rc[:m].cpu().tolist(), "text: GPU rows differ from the CPU reference"
        out["cpu_baseline"] = {"value": len(text) / dt / 1e9, "unit": "GB/s", "cores": 1,
                               "kind": "reference" if which == "ref" else "port",
                               "sample": f"first {m} lines, one exec per line through the C API from Python "
                                         f"(a parity check more than a timing)"}
    return out


NFA_REGEX = rb"[ab]*a[ab]{15}c"      # 2^16 subsets: the subset construction gives up, the NFA tier runs


def bench_nfa(h, steps, warmup, dev, host):
    """the general Thompson tier: a regex that defeats determinisation, over the C2 lines"""
    from oracle import cpu_baseline as baseline
    from sregex_b200 import cuda
    torch = h.torch
    n = dev.shape[0]
    prog = cuda.CudaProgram(NFA_REGEX)
    assert prog.info.dfa_states == 0
    rc = torch.empty(n, dtype=torch.int32, device="cuda")
    total_ms, call_ms, clocks, launches = h.timed(lambda: prog.thompson_lines(dev, n, PITCH, PITCH, out=rc), steps,
                                                  warmup)
    # the warp-per-line kernel (the tier for more than 64 lowered states) on a slice, for comparison
    m = min(n, 1 << 16)
    rcw = torch.empty(m, dtype=torch.int32, device="cuda")
    w_ms, _, _, _ = h.timed(lambda: prog.thompson_lines(dev, m, PITCH, PITCH, engine=cuda.ENGINE_NFA_WARP, out=rcw),
                            3, 3)
    assert torch.equal(rcw, rc[:m]), "NFA tiers disagree"
    bytes_per_step = n * PITCH
    out = {
        "metric": METRICS["nfa"], "value": h.world * bytes_per_step * steps / (total_ms * 1e-3) / 1e9, "unit": "GB/s",
        "steps": steps, "ms_per_step": total_ms / steps, "scaling": "weak",
        "warp_per_line_kernel_gbs": h.world * m * PITCH * 3 / (w_ms * 1e-3) / 1e9,
        "config": {"workload": "Thompson boolean, regex [ab]*a[ab]{15}c (no DFA within 16384 states), "
                               "1,048,576 x 1 KB log lines per GPU", "lines_per_gpu": n,
                   "nfa_states": prog.info.nfa_states, "dfa_states": prog.info.dfa_states,
                   "matched_lines": int((rc == 0).sum())},
        "roofline": h.roofline(bytes_per_step + 4 * n, call_ms, "k_nfa_packed<u32, K> (k_nfa64_lines beyond 4 non-shift movers)",
                               note="ALU-pipe-bound by design: ~12 thread-instructions per byte, 7 of them on the "
                                    "16-lane integer pipe"),
        "gpu_launches": launches, "clocks": clocks,
    }
    if h.rank == 0 and h.world == 1:
        cores = host_cores()
        which = ref_kind()
        ns = min(n, 2048 * cores)
        eng = baseline.ENGINE_JIT if which == "ref" else baseline.ENGINE_THOMPSON
        secs, crc, _ = baseline.run_lines(which, NFA_REGEX, None, host[:ns].numpy(), ns, PITCH, PITCH, eng,
                                          nthreads=cores)
        assert (crc == rc[:ns].cpu().numpy()).all(), "NFA tier: GPU verdicts differ from the CPU reference"
        out["cpu_baseline"] = {"value": ns * PITCH / secs / 1e9, "unit": "GB/s", "cores": cores,
                               "kind": "reference" if which == "ref" else "port",
                               "sample": f"Thompson {'JIT' if which == 'ref' else 'interpreter port'} over the first "
                                         f"{ns} lines, verdicts compared"}
    return out


def c5_shard(h, total, first, count):
    """bytes [first, first+count) of "abccc" x N + "aaabbccb" (bench/gen-data.pl:9 scaled to `total`)"""
    torch = h.torch
    out = torch.empty(count, dtype=torch.uint8, device="cuda")
    unit = torch.tensor(list(b"abccc"), dtype=torch.uint8, device="cuda")
    step = 5 << 24
    rolled = torch.roll(unit, -(first % 5)).repeat(step // 5)
    for i in range(0, count, step):
        m = min(step, count - i)
        out[i:i + m] = rolled[:m]               # step is a multiple of 5: the phase is kept
    if first + count == total:
        out[count - 8:] = torch.tensor(list(b"aaabbccb"), dtype=torch.uint8, device="cuda")
    return out


def bench_c5(h, steps, warmup):
    """one stream, 64 KB reference chunks, sharded over the ranks: halo + record exchange over NCCL"""
    from oracle import cpu_baseline as baseline
    from sregex_b200 import capi, corpus, cuda
    from sregex_b200 import dist as sdist
    torch = h.torch
    total = h.args.c5_bytes
    body = (total - 8) // 5 * 5
    total = body + 8                    # whole "abccc" units, then the tail that matches
    first, count = sdist.shard_range(total // 4096, h.rank, h.world)
    first, count = first * 4096, (count * 4096 if h.rank + 1 < h.world else total - first * 4096)
    shard = c5_shard(h, total, first, count)
    prog = cuda.CudaProgram(corpus.BENCH_REGEX)
    res = [None]

    def step():
        if h.world == 1:
            rc, _, mc = prog.thompson_stream(shard, count, CHUNK, True)
            res[0] = (rc, mc)
        else:
            rc, off = sdist.stream_match_sharded(prog, shard, count, first, eof=True)
            res[0] = (rc, off)

    total_ms, call_ms, clocks, launches = h.timed(step, steps, warmup)
    rc, where = res[0]
    # the match ends on the last byte: the EOF step sees it (last chunk; no in-stream offset)
    assert rc == capi.SRE_OK and where == ((total - 1) // CHUNK if h.world == 1 else -1), (rc, where)
    out = {
        "metric": METRICS["c5"], "value": total * steps / (total_ms * 1e-3) / 1e9, "unit": "GB/s",
        "steps": steps, "ms_per_step": total_ms / steps, "scaling": "strong",
        "config": {"workload": "C5: sre_vm_thompson_exec over one stream in 64 KB chunks with SRE_AGAIN carry "
                               "(bench/gen-data.pl text scaled up, regex of bench/Makefile:62, match at the very "
                               "end), chunk-parallel scan",
                   "stream_bytes": total, "bytes_this_rank": count, "chunk_bytes": CHUNK,
                   "sharding": (f"contiguous parts over {h.world} ranks; per step one all_gather of 256-byte halos, "
                                f"one all_gather of 32-byte records, one all_reduce(min) of the match offset, "
                                f"inside the timed region") if h.world > 1 else "one GPU"},
        # 1 B per input byte + one 32-byte record per 4 KB piece
        "roofline": h.roofline(count + count // 4096 * 32, call_ms,
                               "sre_cuda_thompson_exec_stream: k_stream_pieces + compose / fix / descend / locate",
                               note="whole call on this rank's part (at N > 1 incl. the NCCL exchange)"),
        "gpu_launches": launches, "clocks": clocks,
    }
    try:
        host = h.pinned_copy(shard)
        e2e_steps = 2
        if h.world == 1:
            r = [None]

            def hstep():
                r[0] = prog.thompson_stream_host(host, count, CHUNK, True, slice_bytes=1 << 30)
            dt = h.timed_host(hstep, e2e_steps)
            assert r[0][0] == capi.SRE_OK and r[0][2] == (total - 1) // CHUNK
            call = "sre_cuda_thompson_exec_stream_host (1 GiB slices copied while the previous one is scanned)"
        else:
            stage = torch.empty_like(shard)

            def hstep():
                stage.copy_(host, non_blocking=True)
                res[0] = sdist.stream_match_sharded(prog, stage, count, first, eof=True)
            dt = h.timed_host(hstep, e2e_steps)
            assert res[0][0] == capi.SRE_OK
            call = "H2D of the rank's part + sre_cuda_thompson_stream_reduce / _resolve with the NCCL exchange"
        out["e2e"] = {"value": total * e2e_steps / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": count,
                      "d2h_bytes_per_step": 64, "steps": e2e_steps, "call": call}
        if h.rank == 0 and h.world == 1:
            which = ref_kind()
            sbytes = min(count, 64 << 20)
            eng = baseline.ENGINE_JIT if which == "ref" else baseline.ENGINE_THOMPSON
            secs, last_rc, calls = baseline.run_stream(which, corpus.BENCH_REGEX, None, host[:sbytes].numpy(),
                                                       CHUNK, eng)
            out["cpu_baseline"] = {"value": sbytes / secs / 1e9, "unit": "GB/s", "cores": 1,
                                   "kind": "reference" if which == "ref" else "port",
                                   "sample": f"{'Thompson JIT' if which == 'ref' else 'oracle port'}: one ctx fed "
                                             f"the first {sbytes >> 20} MiB in 64 KB chunks with SRE_AGAIN carry "
                                             f"(one stream = one core)", "last_rc": last_rc}
    except Exception as e:
        out["e2e"] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="all", choices=["all", "c2", "c3", "c4", "c5", "text", "nfa"])
    ap.add_argument("--variant", type=int, default=int(os.environ.get("SRE_VARIANT", "0")))
    ap.add_argument("--lines", type=int, default=NLINES)
    ap.add_argument("--c4-lines", type=int, default=C4_TOTAL_LINES)
    ap.add_argument("--c5-bytes", type=int, default=C5_TOTAL_BYTES)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--engine", default="auto", choices=["auto", "tiled", "skip", "generic", "nfa"])
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    h = Harness(args)
    steps, warmup = args.steps, max(args.warmup, 3)
    few = max(3, min(steps, 20))        # the heavier configs: fewer steps, same rules
    extra = {}
    out = None
    if args.config in ("all", "c2", "c3", "text", "nfa"):
        out, dev, host = bench_c2(h, steps if args.config in ("all", "c2") else few, warmup)
        if args.config == "c3" or (args.config == "all" and not args.no_extras):
            try:
                extra["c3"] = bench_c3(h, few, warmup, dev, host)
            except Exception as e:
                extra["c3"] = {"error": repr(e)}
        if args.config == "text" or (args.config == "all" and not args.no_extras):
            try:
                extra["text"] = bench_text(h, few, warmup, dev)
            except Exception as e:
                extra["text"] = {"error": repr(e)}
        if args.config == "nfa" or (args.config == "all" and not args.no_extras):
            try:
                extra["nfa"] = bench_nfa(h, max(3, min(steps, 10)), warmup, dev, host)
            except Exception as e:
                extra["nfa"] = {"error": repr(e)}
        del dev, host
        h.torch.cuda.empty_cache()
    if args.config == "c4" or (args.config == "all" and not args.no_extras):
        try:
            extra["c4"] = bench_c4(h, max(3, min(steps, 10)), warmup)
        except Exception as e:
            extra["c4"] = {"error": repr(e)}
        h.torch.cuda.empty_cache()
    if args.config == "c5" or (args.config == "all" and not args.no_extras):
        try:
            extra["c5"] = bench_c5(h, max(3, min(steps, 10)), warmup)
        except Exception as e:
            extra["c5"] = {"error": repr(e)}
    h.sampler.stop_flag = True
    h.sampler.join()
    if args.config in ("c3", "c4", "c5", "text", "nfa") and "error" not in extra[args.config]:
        # one config as the headline (profiling runs)
        head = extra.pop(args.config)
        base = {"n_gpus": h.world, "warmup": warmup, "higher_is_better": True, "vs_baseline": None, "dtype": "u8",
                "data": "synthetic"}
        base.update(head)
        out = base
    if h.rank == 0:
        if extra:
            out["extra"] = extra
        print(json.dumps(out))
    if h.world > 1:
        h.dist.barrier()
        h.dist.destroy_process_group()


if __name__ == "__main__":
    main()
