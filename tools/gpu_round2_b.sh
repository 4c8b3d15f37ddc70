set -x
O=gpurun_out/r2n
mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>$O/bench_n2.err | tail -1 > $O/bench_n2.json
tail -c 1500 $O/bench_n2.err
head -c 3000 $O/bench_n2.json
