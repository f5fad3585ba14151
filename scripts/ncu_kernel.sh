#!/bin/bash
# ncu --set full capture (with source) of one launch of kernel $KERNEL (regex) inside a bench.py step; pages exported to gpurun_out/
#   KERNEL=k_linesearch SKIP=30 TAG=r02_ls bash scripts/ncu_kernel.sh
O=gpurun_out; mkdir -p $O
K=${KERNEL:-k_linesearch}; TAG=${TAG:-r02_$K}
CMD="python bench.py --batch ${BATCH:-2048} --steps 1 --warmup 1 --no-cpu-baseline"
timeout 300 $CMD > $O/${TAG}_plain.log 2>&1 || { tail -5 $O/${TAG}_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -s ${SKIP:-30} -c 1 -f -o /tmp/$TAG $CMD > $O/${TAG}_ncu.log 2>&1
ncu -i /tmp/$TAG.ncu-rep --page raw --csv > $O/${TAG}_raw.csv
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source cuda | gzip > $O/${TAG}_src.csv.gz
ncu -i /tmp/$TAG.ncu-rep --page source --csv --print-source sass | gzip > $O/${TAG}_sass.csv.gz
ls -la $O/${TAG}_*
