set -x
O=gpurun_out/r2u
mkdir -p $O
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > $O/sanitizer_memcheck.log 2>&1; tail -6 $O/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_target.py > $O/sanitizer_racecheck.log 2>&1; tail -6 $O/sanitizer_racecheck.log
