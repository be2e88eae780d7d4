#!/bin/bash
# 2 GPUs: C5 domain-divided level times (bench_configs.c5_domain_divided) under the library's knobs.
# usage: tools/sweep_slab2.sh <out.log>   (run under gpurun --gpus 2)
out=${1:-gpurun_out/slab2_sweep.log}
: > $out
run() { # label, env...
  label=$1; shift
  env "$@" FDW_CONFIGS=c5 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_configs.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{') and 'mod_main' in l:
        d = json.loads(l)
        m, r = d['mod_main'], d['rtm_main']
        print('%-44s mod wall %6.1f dev %6.1f | rtm wall %6.1f fwd dev %6.1f bwd dev %6.1f us/level  pslab launches %s' % ('$label', m['us_per_level'], m['level_loop_us_per_level_device_rank0'], r['us_per_level'], r['forward_level_loop_us_per_level_device_rank0'], r['backward_level_loop_us_per_level_device_rank0'], d['persistent_slab_launches_rank0']))
" >> $out
}
export FDW_C5_NT=${NT:-1500}
echo "# 2048 x 2048 model over 2 GPUs (1064-row slabs = the per-GPU slab of the stated model over 8 GPUs), persistent slab kernel" >> $out
export FDW_C5_NX=2048
run "128-thread CTAs (default)" A=1
run "FDW_PSLAB_THREADS=96" FDW_PSLAB_THREADS=96
run "FDW_PSLAB_THREADS=64" FDW_PSLAB_THREADS=64
run "128-thread CTAs again" A=1
run "FDW_PSLAB_THREADS=96 again" FDW_PSLAB_THREADS=96
cat $out
