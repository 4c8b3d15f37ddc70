set -x
O=gpurun_out/r2v
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_same_bits" > $O/pytest_tile.txt 2>&1; tail -5 $O/pytest_tile.txt
timeout 300 python tools/time_tile2.py 100 > $O/time_tile2.txt 2>&1; grep -E "variant|tile v" $O/time_tile2.txt
