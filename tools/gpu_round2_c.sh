set -x
O=gpurun_out/r2o
mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1; nproc >> $O/topo.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k "4-64 or 4-81 or 8-64 or 4-12 or 8-16" > $O/pytest_multi_n4_n8.txt 2>&1; tail -5 $O/pytest_multi_n4_n8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 2>$O/bench_n8.err | tail -1 > $O/bench_n8.json; tail -c 800 $O/bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 2>$O/bench_n4.err | tail -1 > $O/bench_n4.json
timeout 300 $TR --nproc-per-node 2 --master-port 29523 tools/sweep_assembly_multi.py 50 100 150 200 250 300 > $O/sweep_c5_2gpu.jsonl 2>$O/sweep2.err
timeout 300 $TR --nproc-per-node 4 --master-port 29524 tools/sweep_assembly_multi.py 50 100 150 200 250 300 > $O/sweep_c5_4gpu.jsonl 2>$O/sweep4.err
timeout 300 $TR --nproc-per-node 8 --master-port 29525 tools/sweep_assembly_multi.py 50 100 150 200 250 300 400 > $O/sweep_c5_8gpu.jsonl 2>$O/sweep8.err
timeout 300 $TR --nproc-per-node 8 --master-port 29526 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_ref_n8.json
ls -la $O
