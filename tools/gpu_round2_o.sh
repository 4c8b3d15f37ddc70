set -x
O=gpurun_out/r2za
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"k_values_gather" -c 2 -o $O/gather python tools/time_general.py 64 > $O/ncu_run.log 2>&1
ncu -i $O/gather.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/gather.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
tail -2 $O/ncu_run.log
