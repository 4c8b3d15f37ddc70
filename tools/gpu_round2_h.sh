set -x
O=gpurun_out/r2t
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:"k_matfree_color2" -s 8 -c 2 -o $O/matfree python tools/profile_matfree.py 100 > $O/ncu_run.log 2>&1
ncu -i $O/matfree.ncu-rep --page raw --csv > $O/raw.csv 2>/dev/null
ncu -i $O/matfree.ncu-rep --page source --csv > $O/src.csv 2>/dev/null
tail -3 $O/ncu_run.log
