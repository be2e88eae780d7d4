#!/bin/bash
# 2 GPUs: the headline slab step (16384^2 per GPU) under the slab-loop knobs; prints ms per level
out=${1:-gpurun_out/headline2_sweep.log}
: > $out
run() {
  label=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 6 --warmup 3 --no-configs --no-cpu-baseline --no-traffic 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('%-40s %8.1f Gpts/s  %7.4f ms/level  e2e %7.1f' % ('$label', d['value'], d['ms_per_step'] / 250, d['e2e']['value']))
" >> $out
}
run "default (graph, 2 levels)" A=1
run "FDW_GRAPH_LEVELS=8" FDW_GRAPH_LEVELS=8
run "FDW_GRAPH=0 (direct)" FDW_GRAPH=0
run "default (graph, 2 levels) again" A=1
run "FDW_GRAPH_LEVELS=4" FDW_GRAPH_LEVELS=4
run "FDW_FUSE_FLAGS=0" FDW_FUSE_FLAGS=0
cat $out
