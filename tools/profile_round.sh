set -x
mkdir -p gpurun_out/r1k
python bench.py --steps 20 --warmup 5 2>gpurun_out/r1k/bench_n1.err | tail -1 > gpurun_out/r1k/bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r1k/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r1k/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r1k/launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_values_tile|k_spmv_group" -c 6 -o gpurun_out/r1k/full python tools/profile_target.py 100 1 > gpurun_out/r1k/full_run.log 2>&1
ncu -i gpurun_out/r1k/full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum > gpurun_out/r1k/ncu_raw.csv 2>&1
python tools/sweep_assembly.py > gpurun_out/r1k/sweep.jsonl 2>gpurun_out/r1k/sweep.err
ls -la gpurun_out/r1k
