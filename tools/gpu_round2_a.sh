set -x
O=gpurun_out/r2m
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 2>$O/bench_n1.err | tail -1 > $O/bench_n1.json; tail -c 600 $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 2 2>$O/bench_ref.err | tail -1 > $O/bench_ref.json
ncu --set full --clock-control none --import-source on -k regex:"k_values_tile2" -c 4 -o $O/tile2 python tools/profile_target.py 100 1 > $O/ncu_run.log 2>&1
ncu -i $O/tile2.ncu-rep --page source --csv > $O/src2.csv 2>/dev/null
python tools/time_tile2_out.py 100 > $O/out_v2.txt 2>&1
nproc > $O/host.txt; lscpu | head -20 >> $O/host.txt; free -g >> $O/host.txt; nvidia-smi topo -m >> $O/host.txt 2>&1
ls -la $O
