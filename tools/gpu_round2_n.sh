set -x
O=gpurun_out/r2z
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "general" > $O/pytest_general.txt 2>&1; tail -15 $O/pytest_general.txt
timeout 300 python tools/time_general.py 64 > $O/time_general.txt 2>&1; cat $O/time_general.txt
