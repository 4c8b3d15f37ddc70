O=gpurun_out/r2zg
mkdir -p $O
SMFEM_TILE=v3 ncu --set full --clock-control none --import-source on -k regex:"k_values_tile3" -s 2 -c 1 -o $O/tile3 python tools/profile_target.py 100 1 > $O/ncu_run.log 2>&1
ncu -i $O/tile3.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/tile3.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
tail -2 $O/ncu_run.log
