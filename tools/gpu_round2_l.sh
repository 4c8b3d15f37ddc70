set -x
O=gpurun_out/r2zd
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > $O/pytest_multi_n2.txt 2>&1; tail -6 $O/pytest_multi_n2.txt | cut -c1-900
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2>$O/bench_n2.err | tail -1 > $O/bench_n2.json
tail -c 600 $O/bench_n2.err
python -c "
import json; d=json.load(open('$O/bench_n2.json')); print(d['value'], d['pcg']['wait_us_per_iter'], d['pcg']['ms_per_iter'], d['pcg']['matrix_free'], d['pcg']['multigrid']['ms_total'], d['pcg']['multigrid_matrix_free'])"
