#!/bin/bash
# GPU check after a change to the Pike tiers: the whole GPU suite (the determinised Pike VM is the
# default tier for programs with look-ahead assertions; SRE_PDFA_LOOKAHEAD=0 turns that off), then
# the timing of the tiers on regexes with assertions.  Run on the GPU box from the repo root; logs
# under gpurun_out/.
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/la_pytest.log 2>&1
echo "pytest rc=$?"
tail -3 $O/la_pytest.log
N=$((1 << 18)) timeout 120 python tools/pike_lookahead_bench.py > $O/la_bench.log 2>&1
echo "bench rc=$?"
tail -12 $O/la_bench.log
