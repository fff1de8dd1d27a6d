#!/usr/bin/env python3
"""bench.py -- headline benchmark of the sregex match-execution hot path on B200.

Workload (BASELINE.json configs[1]): Thompson boolean match of one regex over
1,048,576 independent 1 KB synthetic log lines per GPU.
A "step" is one pass of the hot path over that batch.

  python bench.py --gpus N --steps K --warmup W          (our arm)
  python bench.py --impl reference ...                    (reference CPU arm)

Prints ONE JSON line (rank 0).  `value` = input GB/s with the corpus resident
in HBM (CUDA events, max over ranks); `e2e` = the same metric through the
host-buffer C-ABI call (H2D of the corpus and D2H of the verdicts inside the
timed region); `roofline` = algorithmic bytes / kernel time vs the measured HBM
peak; `cpu_baseline` = the reference's own Thompson paths on the host cores over
a bounded sample.
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
REGEX_NAME = 'HTTP/1\\.[01]" 5\\d\\d '
METRIC = "input GB/s scanned (Thompson boolean, 1 regex, 1M x 1 KB lines per GPU)"


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
    """SM clock + throttle reasons during the timed region (NVML, 4 ms period).
    Started before the warm-up steps -- the first NVML queries take the driver
    for milliseconds and would stall kernel launches inside a short timed region
    -- and told with mark() where the timed region begins."""

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
        """the timed region starts now: earlier samples (but the last one) and
        earlier throttle reasons do not count"""
        self.first = max(0, len(self.samples) - 1)
        self.reasons = set()

    def result(self):
        s = sorted(self.samples[self.first:])
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_baseline_sample(lines_host, regex, cores, which):
    """reference Thompson JIT + interpreter over a bounded sample, all cores"""
    from oracle import cpu_baseline as baseline
    n = lines_host.shape[0]
    out = {}
    for name, eng in (("jit", baseline.ENGINE_JIT), ("interp", baseline.ENGINE_THOMPSON)):
        if which == "oracle" and name == "jit":
            continue
        secs, rc, _ = baseline.run_lines(which, regex, None, lines_host, n, PITCH, PITCH, eng,
                                         nthreads=cores)
        out[name] = n * PITCH / secs / 1e9
        out["hits"] = int((rc == 0).sum())
    return out


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    from oracle import cpu_baseline as baseline
    from sregex_b200 import corpus
    cores = host_cores()
    which = "ref" if baseline.available("ref") else "oracle"
    engine = baseline.ENGINE_JIT if which == "ref" else baseline.ENGINE_THOMPSON
    n = min(NLINES, 8192 * cores)
    lines = corpus.log_lines(n, PITCH).numpy()
    for _ in range(args.warmup):
        baseline.run_lines(which, corpus.C2_REGEX, None, lines, n, PITCH, PITCH, engine, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        baseline.run_lines(which, corpus.C2_REGEX, None, lines, n, PITCH, PITCH, engine, nthreads=cores)
    dt = time.perf_counter() - t0
    value = n * PITCH * args.steps / dt / 1e9
    sample = (f"{n} of the {NLINES} lines per step ({n * PITCH >> 20} MiB), "
              f"{'sre_vm_thompson_jit handler' if which == 'ref' else 'oracle port of sre_vm_thompson_exec'}"
              f", one private program per thread, fresh ctx+pool per line")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic",
        "config": {"workload": "C2: Thompson boolean, regex " + REGEX_NAME + ", 1 KB log lines",
                   "lines_per_step": n, "line_bytes": PITCH},
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": cores,
                         "kind": "reference" if which == "ref" else "port", "sample": sample},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--variant", type=int, default=int(os.environ.get("SRE_VARIANT", "0")))
    ap.add_argument("--lines", type=int, default=NLINES)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--engine", default="auto", choices=["auto", "tiled", "skip", "generic", "nfa"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from oracle import cpu_baseline as baseline
    from sregex_b200 import corpus, cuda

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.lines

    # corpus shard of this rank, resident in HBM (1 GiB > 126 MB of L2, so every
    # step streams it from HBM again)
    dev = torch.empty((n, PITCH), dtype=torch.uint8, device="cuda")
    blk = 1 << 17
    for i in range(0, n, blk):
        m = min(blk, n - i)
        dev[i:i + m] = corpus.log_lines(m, PITCH, device="cuda", first_line=rank * n + i)
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

    def step():
        prog.thompson_lines(dev, n, PITCH, PITCH, engine=engine, out=rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    # the closing match count (and, at N > 1, NCCL's first collective) pay a one-off
    # set-up cost (lazy kernel loading): not part of a step
    warm = (rc == 0).sum()
    if world > 1:
        dist.all_reduce(warm)
    barrier()

    while sampler.nv is not None and len(sampler.samples) < 3:
        time.sleep(0.002)       # the sampler's first queries are over before anything is timed
    cuda.launch_count(reset=True)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    t_start.record()
    for a, b in evs:
        a.record()
        step()
        b.record()
    hits = (rc == 0).sum()
    if world > 1:
        dist.all_reduce(hits)
    t_end.record()
    barrier()
    launches = cuda.launch_count()
    sampler.stop_flag = True
    sampler.join()

    total_ms = t_start.elapsed_time(t_end)
    kern_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps
    tmax = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    bytes_per_step = n * PITCH
    value = world * bytes_per_step * args.steps / (total_ms * 1e-3) / 1e9

    # roofline of the dominant kernel (k_dfa_lines): algorithmic bytes per launch
    # = 1 B read per input byte + 4 B verdict per line
    algo_bytes = bytes_per_step + 4 * n
    peak, peak_src = measured_hbm_peak()
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("bytes_per_launch")
        except Exception:
            traffic = None

    out = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {
            "workload": "C2: Thompson boolean, regex " + REGEX_NAME + ", 1,048,576 x 1 KB log lines per GPU",
            "lines_per_gpu": n, "line_bytes": PITCH, "sharding": f"lines x {world} ranks, no data-path collective",
            "engine": engine_name, "dfa_states": info.dfa_states,
            "nfa_states": info.nfa_states, "variant": args.variant,
            "l2_policy": "input (1 GiB per GPU) larger than L2 (126 MB); no flush needed",
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "kernel": {"dfa_skip": "k_dfa_lines_skipw", "dfa_tiled": "k_dfa_lines_tma_early"}.get(
                         engine_name, engine_name),
                     "kernel_ms": kern_ms, "algorithmic_bytes": algo_bytes},
        "gpu_launches": launches,
        "clocks": sampler.result(),
        "hits": int(hits.item()),
    }

    # ---- e2e through the host-buffer C-ABI call (every rank, its own shard) -------
    e2e_steps = max(2, min(5, args.steps))
    host = torch.empty((n, PITCH), dtype=torch.uint8, pin_memory=True)
    host.copy_(dev)
    host_rc = torch.empty(n, dtype=torch.int32, pin_memory=True)
    prog.thompson_lines_host(host, n, PITCH, PITCH, host_rc)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prog.thompson_lines_host(host, n, PITCH, PITCH, host_rc)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    assert torch.equal(host_rc, rc.cpu()), "host-buffer path disagrees with the device-resident path"
    out["e2e"] = {"value": world * bytes_per_step * e2e_steps / dt / 1e9, "unit": "GB/s",
                  "h2d_bytes_per_step": bytes_per_step, "d2h_bytes_per_step": 4 * n,
                  "steps": e2e_steps, "note": "pinned host buffers, all ranks concurrently, max over ranks"}

    if rank == 0:
        # ---- CPU baseline on this box's host cores (N=1 only) ----------------------
        if world == 1:
            cores = host_cores()
            which = "ref" if baseline.available("ref") else "oracle"
            ns = min(n, 8192 * cores)
            sample = host[:ns].numpy()
            cb = cpu_baseline_sample(sample, corpus.C2_REGEX, cores, which)
            best = cb.get("jit", cb.get("interp"))
            out["cpu_baseline"] = {
                "value": best, "unit": "GB/s", "cores": cores,
                "kind": "reference" if which == "ref" else "port",
                "sample": f"first {ns} lines ({ns * PITCH >> 20} MiB) of the workload, one OS thread per core, "
                          f"private program per thread, fresh ctx+pool per line",
                "thompson_jit_gbs": cb.get("jit"), "thompson_interp_gbs": cb.get("interp"),
            }
            assert cb["hits"] == int((host_rc[:ns] == 0).sum()), "CPU baseline disagrees with GPU verdicts"

        # ---- extras: the other BASELINE configs at bench size (not the headline) ----
        if not args.no_extras and world == 1:
            extra = {}

            def timed_gbs(fn, nbytes, reps=2):
                fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    fn()
                b.record()
                torch.cuda.synchronize()
                return nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9

            try:
                # C3: Pike with 4 capture groups over every line (gate + start hint + Pike kernels)
                p3 = cuda.CudaProgram(corpus.C3_REGEX)
                prc = torch.empty(n, dtype=torch.int32, device="cuda")
                pov = torch.empty((n, p3.nslots), dtype=torch.int64, device="cuda")
                extra["c3_pike_4groups_gbs"] = timed_gbs(
                    lambda: p3.pike_lines(dev, n, PITCH, PITCH, out_rc=prc, out_ovec=pov), n * PITCH)
                extra["c3_lines"] = n
                extra["c3_matched_lines"] = int((prc == 0).sum())
                # ... beside the reference's Pike VM on the host cores, same rows compared
                cores = host_cores()
                which = "ref" if baseline.available("ref") else "oracle"
                ns = min(n, 2048 * cores)
                secs, crc, cov = baseline.run_lines(which, corpus.C3_REGEX, None, host[:ns].numpy(), ns, PITCH,
                                                    PITCH, baseline.ENGINE_PIKE, nthreads=cores,
                                                    ovec_slots=p3.nslots)
                extra["c3_cpu_pike_gbs"] = ns * PITCH / secs / 1e9
                extra["c3_cpu_sample_lines"] = ns
                assert (prc[:ns].cpu().numpy() == crc).all() and (pov[:ns].cpu().numpy() == cov).all(), \
                    "C3: GPU Pike rows differ from the CPU reference"
                # C4: 64-pattern set: which pattern matched (Thompson gate, then Pike on the hits)
                pm = cuda.CudaProgram(corpus.multi_pattern_set(64))
                m = n
                extra["c4_multi64_gate_gbs"] = timed_gbs(lambda: pm.thompson_lines(dev, n, PITCH, PITCH), n * PITCH)
                mrc = torch.empty(m, dtype=torch.int32, device="cuda")
                mov = torch.empty((m, pm.nslots), dtype=torch.int64, device="cuda")
                extra["c4_multi64_id_gbs"] = timed_gbs(
                    lambda: pm.pike_lines(dev, m, PITCH, PITCH, out_rc=mrc, out_ovec=mov), m * PITCH)
                extra["c4_lines"] = m
                extra["c4_matched_fraction"] = float((mrc >= 0).float().mean())
                ns4 = min(n, 128 * cores)
                secs, crc, cov = baseline.run_lines(which, corpus.multi_pattern_set(64), None, host[:ns4].numpy(),
                                                    ns4, PITCH, PITCH, baseline.ENGINE_PIKE, nthreads=cores,
                                                    ovec_slots=pm.nslots)
                extra["c4_cpu_pike_gbs"] = ns4 * PITCH / secs / 1e9
                extra["c4_cpu_sample_lines"] = ns4
                assert (mrc[:ns4].cpu().numpy() == crc).all() and (mov[:ns4].cpu().numpy() == cov).all(), \
                    "C4: GPU matched ids / ovectors differ from the CPU reference"
                extra["c4_dfa_states"] = pm.info.dfa_states
                # C5: one stream, chunk-parallel transfer-function scan (the whole resident corpus
                # as a single stream, 64 KB reference chunks)
                p1 = cuda.CudaProgram(corpus.BENCH_REGEX)
                flat = dev.view(-1)
                extra["c5_stream_scan_gbs"] = timed_gbs(
                    lambda: p1.thompson_stream(flat, flat.numel(), 65536, True), flat.numel(), reps=3)
                extra["c5_stream_bytes"] = flat.numel()
            except Exception as e:      # extras never break the headline line
                extra["error"] = repr(e)
            out["extra"] = extra
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
