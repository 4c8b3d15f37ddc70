set -x
O=gpurun_out/r2y
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k "4-64 or 4-81 or 8-64 or 4-12 or 8-16 or 4-16 or 8-40" > $O/pytest_multi_n4_n8.txt 2>&1; tail -4 $O/pytest_multi_n4_n8.txt | cut -c1-600
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 2>$O/bench_n8.err | tail -1 > $O/bench_n8.json; tail -c 400 $O/bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 2>$O/bench_n4.err | tail -1 > $O/bench_n4.json
ls -la $O
