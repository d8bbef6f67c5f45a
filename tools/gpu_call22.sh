set -x
O=gpurun_out/r02w
mkdir -p $O
V=variants
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_next.py -m gpu -q -x -k "interp" > $O/pytest_interp.log 2>&1; tail -3 $O/pytest_interp.log
for k in interp interp_ring; do
  timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_interp_nodiamond.so > $O/ab_$k.log 2>&1; cat $O/ab_$k.log
done
