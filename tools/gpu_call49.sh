set -x
O=gpurun_out/r03z
mkdir -p $O
V=variants
timeout 900 python -m pytest tests -m gpu -q -x -k "angle_to_pixel or a2p or vector or interp or config or smoke or api" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python tools/stress_a2p_fast.py > $O/stress_a2p.log 2>&1; tail -2 $O/stress_a2p.log
timeout 300 python tools/stress_v2p_fast.py > $O/stress_v2p.log 2>&1; tail -1 $O/stress_v2p.log
timeout 300 python tools/ab_multi.py 30 a2p 5 3 default $V/lib_prev.so > $O/ab_a2p.log 2>&1; cat $O/ab_a2p.log
timeout 300 python tools/ab_multi.py 28 a2p 10 3 default $V/lib_prev.so > $O/ab_a2p28.log 2>&1; cat $O/ab_a2p28.log
timeout 300 python tools/ab_multi.py 28 v2p 5 3 default $V/lib_prev.so > $O/ab_v2p.log 2>&1; cat $O/ab_v2p.log
