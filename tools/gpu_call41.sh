set -x
O=gpurun_out/r03r
mkdir -p $O
timeout 300 python tools/sanitize_smoke.py > $O/plain.log 2>&1; tail -3 $O/plain.log
timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py > $O/memcheck.log 2>&1; tail -8 $O/memcheck.log
