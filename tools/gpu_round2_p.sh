O=gpurun_out/r2zb
mkdir -p $O
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_values_gather|k_gather_elem|k_extract_diag" -c 12 --csv --log-file $O/launches.csv python tools/time_general.py 64 > $O/run.log 2>&1
grep -E "k_values_gather|k_gather_elem" $O/launches.csv | cut -d, -f5,13- | head -24
