set -x
O=gpurun_out/r03h
mkdir -p $O
V=variants
timeout 600 python -m pytest tests -m gpu -q -x -k "vector" > $O/pytest_v2p.log 2>&1; tail -3 $O/pytest_v2p.log
timeout 300 python tools/stress_v2p_fast.py > $O/stress_v2p.log 2>&1; tail -2 $O/stress_v2p.log
timeout 300 python tools/ab_multi.py 28 v2p 5 3 default $V/lib_v2p_prev.so > $O/ab_v2p.log 2>&1; cat $O/ab_v2p.log
