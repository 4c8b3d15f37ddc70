set -x
O=gpurun_out/r2r
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "extract_borders or project_nodes or golden" > $O/pytest_borders.txt 2>&1; tail -15 $O/pytest_borders.txt
