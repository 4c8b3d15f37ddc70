# The commands behind profiles/r2_final_* (run through gpurun; one GPU unless noted).
set -x
O=${O:-gpurun_out/final}
mkdir -p $O
# 1 GPU: tests, bench (both arms), smoke, ncu launch list, ncu --set full capture of the three hot kernels
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1
python bench.py --steps 20 --warmup 5 2>$O/bench_n1.err | tail -1 > $O/bench_n1.json
python bench.py --impl reference --steps 5 --warmup 2 2>$O/bench_ref.err | tail -1 > $O/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-solve > $O/launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_values_tile2|k_spmv_group|k_matfree_color2" -c 14 -o $O/full python tools/profile_target.py 100 1 > $O/full_run.log 2>&1
# N GPUs (gpurun --gpus N): multi-GPU parity tests and the bench line
#   python -m pytest tests/test_gpu_multi.py -m gpu -q -s
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus N --steps 10 --warmup 3
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29523 tools/sweep_assembly_multi.py 50 100 150 200 250 300
# experiments: tools/time_tile2.py (layer-march variants + ablation), tools/time_matfree.py, tools/time_general.py,
#   tools/solve_matfree_big.py, tools/sanitize_target.py under compute-sanitizer (memcheck, racecheck)
