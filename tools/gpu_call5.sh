set -x
O=gpurun_out/r02e
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
V=variants
timeout 300 python tools/ab_multi.py 28 p2a 10 3 default $V/lib_p2a_map.so > $O/ab_p2a_28.log 2>&1
timeout 300 python tools/ab_multi.py 28 p2a_ring 10 3 default $V/lib_p2a_map.so > $O/ab_p2a_ring_28.log 2>&1
ls -la $O
