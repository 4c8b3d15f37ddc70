set -x
O=gpurun_out/r2s
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matrix_free" > $O/pytest_matfree.txt 2>&1; tail -15 $O/pytest_matfree.txt
timeout 300 python tools/time_matfree.py 100 > $O/time_matfree.txt 2>&1; cat $O/time_matfree.txt
