#!/bin/bash
# ncu --set full captures of the epilogue kernels with the shipped geometry, one strip + one bulk
# launch of the last recorded level of each phase; 8192 x 8192 (+40) grid.  Run under gpurun; reports -> gpurun_out/<tag>_*.ncu-rep
tag=${1:-prof}
export FDW_LEVEL_GRAPH=0   # plain launches: -s / -c count them in issue order
python tools/prof_cpu_family.py > gpurun_out/${tag}_plain_run.log 2>&1 || exit 1
N="ncu --set full --clock-control none --import-source on -k regex:k_step"
WHICH=plain  NT=3 $N -s 3 -c 2  -o gpurun_out/${tag}_c_plain  python tools/prof_cpu_family.py > gpurun_out/${tag}_ncu.log 2>&1
WHICH=model  NT=3 $N -s 3 -c 2  -o gpurun_out/${tag}_c_record python tools/prof_cpu_family.py >> gpurun_out/${tag}_ncu.log 2>&1
WHICH=rtm    NT=3 $N -s 3 -c 2  -o gpurun_out/${tag}_c_hstore python tools/prof_cpu_family.py >> gpurun_out/${tag}_ncu.log 2>&1
WHICH=rtm    NT=3 $N -s 8 -c 2  -o gpurun_out/${tag}_c_inject_img python tools/prof_cpu_family.py >> gpurun_out/${tag}_ncu.log 2>&1
WHICH=gpufam NT=4 $N -s 12 -c 3 -o gpurun_out/${tag}_g_back  python tools/prof_cpu_family.py >> gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}_*.ncu-rep
