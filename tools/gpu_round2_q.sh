set -x
O=gpurun_out/r2zc
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matrix_free" > $O/pytest_matfree.txt 2>&1; tail -12 $O/pytest_matfree.txt
timeout 600 python tools/solve_matfree_big.py 100 200 300 400 > $O/solve_big.txt 2>&1; cat $O/solve_big.txt
