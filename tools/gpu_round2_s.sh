set -x
O=gpurun_out/r2zf
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
python bench.py --steps 20 --warmup 5 2>$O/bench_n1.err | tail -1 > $O/bench_n1.json; tail -c 300 $O/bench_n1.err
python bench.py --impl reference --steps 5 --warmup 2 2>$O/bench_ref.err | tail -1 > $O/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
