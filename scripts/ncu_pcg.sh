#!/bin/bash
# ncu --set full capture of one launch of the PCG kernel for each variant in $VARIANTS; raw pages exported to gpurun_out/
O=gpurun_out; mkdir -p $O
for V in ${VARIANTS:-3 6}; do
  export B2T_PCG_VARIANT=$V
  timeout 300 python scripts/pcg_probe.py ${BATCH:-296} > $O/probe_v$V.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pcg -s 1 -c 1 -f -o /tmp/pcg_v$V python scripts/pcg_probe.py ${BATCH:-296} > $O/ncu_v$V.log 2>&1
  ncu -i /tmp/pcg_v$V.ncu-rep --page raw --csv > $O/${TAG:-r02}_pcg_v${V}_raw.csv
  ncu -i /tmp/pcg_v$V.ncu-rep --page source --csv --print-source sass | gzip > $O/${TAG:-r02}_pcg_v${V}_sass.csv.gz
  tail -2 $O/probe_v$V.log
done
ls -la $O | tail
