#!/bin/bash
# One ncu --set full capture per dominant kernel (each after its bench config has run clean without
# ncu), plus the launch list of the default bench.  Run on the GPU box from the repo root:
#   bash tools/profile_round.sh r02      -> gpurun_out/r02_ncu_summary.txt, r02_launches.csv
R=${1:-r02}
O=gpurun_out
cap() {   # name, kernel regex, skip, bench args...
    local name=$1 rx=$2 skip=$3; shift 3
    python bench.py "$@" --steps 2 --warmup 3 > $O/${R}_plain_$name.log 2>&1 || { echo "$name: plain run failed"; return; }
    ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/${R}_$name \
        python bench.py "$@" --steps 2 --warmup 3 > $O/${R}_ncu_$name.log 2>&1
    echo "$name: $(ls -la $O/${R}_$name.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
    # the summary is made here and the report dropped: gpurun copies back at most 64 MiB
    python tools/ncu_summary.py $O/${R}_$name.ncu-rep >> $O/${R}_ncu_summary.txt 2>/dev/null
    rm -f $O/${R}_$name.ncu-rep
}
rm -f $O/${R}_ncu_summary.txt
cap c2_skipw     'k_dfa_lines_skipw'      3 --config c2 --no-extras
cap c2_tiled     'k_dfa_lines_tma_early'  3 --config c2 --no-extras --engine tiled
cap c3_hint      'k_dfa_lines_hint'       2 --config c3
cap c3_lineage   'k_pike_lineage'         2 --config c3
cap c4_big       'k_dfa_lines_big'        2 --config c4 --c4-lines 1048576
cap c5_pieces    'k_stream_pieces'        1 --config c5 --c5-bytes 4294967296
cap text_verdicts 'k_text_verdicts$'      2 --config text
cap text_finish  'k_text_finish'          2 --config text
cap nfa_packed   'k_nfa_packed'           1 --config nfa
# launch list of the default bench (kernel durations; cold-cache, serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|::k_' -c 600 --csv \
    --log-file $O/${R}_launches.csv python bench.py --steps 2 --warmup 1 > $O/${R}_ncu_launches.log 2>&1
echo "launch list: $(grep -c gpu__time_duration $O/${R}_launches.csv) launches"
