set -x
O=gpurun_out/r2w
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 2>$O/bench_n1.err | tail -1 > $O/bench_n1.json; tail -c 400 $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 2 2>$O/bench_ref.err | tail -1 > $O/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-solve > $O/launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_values_tile2|k_spmv_group|k_matfree_color2" -c 14 -o $O/full python tools/profile_target.py 100 1 > $O/full_run.log 2>&1
ncu -i $O/full.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,lts__t_sectors_srcunit_tex_op_write.sum > $O/ncu_raw.csv 2>&1
ls -la $O
