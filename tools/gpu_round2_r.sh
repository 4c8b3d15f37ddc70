set -x
O=gpurun_out/r2ze
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s -k "4-64 or 8-64" > $O/pytest_multi_n4_n8_b.txt 2>&1; tail -3 $O/pytest_multi_n4_n8_b.txt | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 2>$O/bench_n8.err | tail -1 > $O/bench_n8.json; tail -c 300 $O/bench_n8.err
timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 2>$O/bench_n4.err | tail -1 > $O/bench_n4.json
ls -la $O
