#!/bin/bash
# C5-grid level-time sweep (tools/level_overhead.py): strip launches side by side or one by one, chunking rule, bulk rows per CTA
out=${1:-gpurun_out/level_sweep.log}
: > $out
python tools/level_overhead.py >> $out 2>&1
FDW_RPC_RULE=0 python tools/level_overhead.py >> $out 2>&1
FDW_MULTIRECT=0 python tools/level_overhead.py >> $out 2>&1
FDW_MULTIRECT=0 FDW_RPC_RULE=0 python tools/level_overhead.py >> $out 2>&1
for rpc in 5 6; do
  FDW_ROWS_PER_CTA=$rpc python tools/level_overhead.py >> $out 2>&1
done
