#!/bin/bash
# One gpurun call that produces the artefacts summarised under profiles/ for a version tag:
#   gpurun --timeout 1500 -- 'bash scripts/final_measure.sh r01_v7'
# bench lines are taken WITHOUT a profiler; the ncu passes come after and their printed numbers are never bench values.
# The .ncu-rep files exceed gpurun's 64 MiB return limit: their raw pages (and the SASS page of the dominant kernel) are exported
# here and the reports deleted; scripts/profile_summary.py reads the exported CSVs.
TAG=${1:-r01_vX}
O=gpurun_out
mkdir -p $O
set -x
timeout 600 python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err || exit 1
timeout 300 python bench.py --limits 0 --no-cpu-baseline > $O/${TAG}_bench_nolimits.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_ncu_launches.csv \
  python bench.py --batch 2048 --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_launches.log 2>&1
# dominant kernel: k_pcg_tm (the bulk passes; k_pcg3 runs the passes with fewer active instances than SMs)
PK=${PCG_KERNEL:-k_pcg_tm}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$PK -s ${PCG_SKIP:-20} -c 1 -f -o /tmp/${TAG}_pcg \
  python bench.py --batch 2048 --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_pcg.log 2>&1
ncu -i /tmp/${TAG}_pcg.ncu-rep --page raw --csv > $O/${TAG}_pcg_raw.csv
ncu -i /tmp/${TAG}_pcg.ncu-rep --page source --csv --print-source sass | gzip > $O/${TAG}_pcg_sass.csv.gz
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_schur_rows|k_pinv|k_linesearch|k_kkt_diag|k_fd' -s 40 -c 5 -f -o /tmp/${TAG}_others \
  python bench.py --batch 2048 --steps 1 --warmup 1 --no-cpu-baseline > $O/${TAG}_ncu_others.log 2>&1
ncu -i /tmp/${TAG}_others.ncu-rep --page raw --csv > $O/${TAG}_others_raw.csv
ncu -i /tmp/${TAG}_others.ncu-rep --page source --csv --print-source sass -k regex:k_linesearch | gzip > $O/${TAG}_linesearch_sass.csv.gz
ls -la /tmp/*.ncu-rep $O | tail -20
tail -c 600 $O/${TAG}_bench_default.json
