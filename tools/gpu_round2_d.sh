set -x
O=gpurun_out/r2p
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "layer_march or side_stream or tile" > $O/pytest_tile.txt 2>&1; tail -5 $O/pytest_tile.txt
timeout 300 python tools/time_tile2.py 100 > $O/time_tile2.txt 2>&1; cat $O/time_tile2.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s -k "single_process" > $O/pytest_single_process.txt 2>&1; tail -12 $O/pytest_single_process.txt | cut -c1-600
