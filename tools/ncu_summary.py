"""Summarises .ncu-rep files (ncu --set full captures brought back in gpurun_out/) into the text
that is committed under profiles/: duration, DRAM traffic, pipe utilisation, occupancy, stalls.
  python tools/ncu_summary.py gpurun_out/r02_*.ncu-rep > profiles/r02_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/shared data pipe % (all wavefronts)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory pipe %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]
STALLS = "smsp__average_warps_issue_stalled_"

for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print(f"## {path}: no data\n")
        continue
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        col = {n: (r[i], units[i]) for i, n in enumerate(head)}
        print(f"## {path.split('/')[-1]}: {col['Kernel Name'][0][:110]}")
        for key, label in WANT:
            if key in col:
                print(f"   {label:32s} {col[key][0]:>16s} {col[key][1]}")
        st = sorted(((float(v[0]), n[len(STALLS):].replace('_per_issue_active.ratio', '')) for n, v in col.items()
                     if n.startswith(STALLS) and n.endswith("_per_issue_active.ratio") and v[0]), reverse=True)
        print("   top stalls (warps per issue):    " + ", ".join(f"{n} {x:.2f}" for x, n in st[:5]))
        print()
