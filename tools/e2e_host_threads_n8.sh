O=gpurun_out/r2zi
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for th in 0 1 2; do
SMFEM_HOST_THREADS=$th timeout 200 $TR --nproc-per-node 8 --master-port 2953$th bench.py --gpus 8 --steps 5 --warmup 3 --no-solve 2>$O/e2e_t$th.err | tail -1 > $O/e2e_t$th.json
python -c "
import json; d=json.load(open('$O/e2e_t$th.json')); e=d['e2e']; print('threads',$th,'e2e ms',e['ms_per_step'],e['ms_per_step_all'],'h2d',e['h2d_bytes_per_step'])"
done
