#!/bin/bash
# GPU check of the determinised Pike VM for programs with look-ahead assertions (k_pike_lineage
# with SRE_PDFA_LOOKAHEAD=1): every GPU test that reaches sre_cuda_pike_exec_lines, then the timing
# of the two tiers.  Run on the GPU box from the repo root; logs under gpurun_out/.
O=gpurun_out
mkdir -p $O
export SRE_PDFA_LOOKAHEAD=1
timeout 420 python -m pytest tests -m gpu -q -p no:cacheprovider \
    -k "pike or tier or fuzz_gpu or assertions or multi_regex or cli or misfire or global_scan or lookahead or two_threads" \
    > $O/la_pytest.log 2>&1
echo "pytest (look-ahead on) rc=$?"
tail -3 $O/la_pytest.log
N=$((1 << 18)) timeout 120 python tools/pike_lookahead_bench.py > $O/la_bench.log 2>&1
echo "bench rc=$?"
cat $O/la_bench.log | tail -8
