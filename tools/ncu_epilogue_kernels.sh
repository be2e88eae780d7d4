#!/bin/bash
# ncu --set full captures of the epilogue kernels with the shipped geometry, one strip + one bulk
# launch of the last recorded level of each phase; 8192 x 8192 (+40) grid.  Run under gpurun.  The reports embed the
# library's whole 50 MB fatbin, so they are reduced ON the box to the raw-metric CSV (gpurun_out/<tag>_<name>.raw.csv)
# and deleted (gpurun merges at most 64 MiB back).
tag=${1:-prof}
export FDW_LEVEL_GRAPH=0   # plain launches: -s / -c count them in issue order
python tools/prof_cpu_family.py > gpurun_out/${tag}_plain_run.log 2>&1 || exit 1
N="ncu --set full --clock-control none -k regex:k_step"
cap() { # name, WHICH, NT, skip, count
  WHICH=$2 NT=$3 $N -s $4 -c $5 -o /tmp/${tag}_$1 python tools/prof_cpu_family.py >> gpurun_out/${tag}_ncu.log 2>&1
  ncu -i /tmp/${tag}_$1.ncu-rep --page raw --csv > gpurun_out/${tag}_$1.raw.csv 2>> gpurun_out/${tag}_ncu.log
  rm -f /tmp/${tag}_$1.ncu-rep
}
cap c_plain      plain  3 3 2
cap c_record     model  3 3 2
cap c_hstore     rtm    3 3 2
cap c_inject_img rtm    3 8 2
cap g_back       gpufam 4 12 3
grep -v "^==PROF==" gpurun_out/${tag}_ncu.log > gpurun_out/${tag}_ncu.txt; rm -f gpurun_out/${tag}_ncu.log
ls -la gpurun_out/${tag}_*
