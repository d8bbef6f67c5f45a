set -x
O=gpurun_out/r03u
mkdir -p $O
V=variants
timeout 900 python -m pytest tests -m gpu -q -x -k "angle_to_pixel or a2p or vector or config or smoke or api" > $O/pytest.log 2>&1; tail -3 $O/pytest.log
timeout 300 python tools/stress_a2p_fast.py > $O/stress_a2p.log 2>&1; tail -2 $O/stress_a2p.log
timeout 300 python tools/stress_v2p_fast.py > $O/stress_v2p.log 2>&1; tail -2 $O/stress_v2p.log
timeout 300 python tools/ab_multi.py 30 a2p 5 3 default $V/lib_a2p_nofold.so > $O/ab_a2p.log 2>&1; cat $O/ab_a2p.log
timeout 300 python tools/ab_multi.py 28 v2p 5 3 default $V/lib_v2p_nofold.so > $O/ab_v2p.log 2>&1; cat $O/ab_v2p.log
timeout 300 python tools/sustained_a2p.py 30 250 a2p > $O/sustained_fold.txt 2>&1; head -18 $O/sustained_fold.txt
HPGEOM_B200_LIB=$PWD/variants/lib_a2p_nofold.so timeout 300 python tools/sustained_a2p.py 30 250 a2p > $O/sustained_nofold.txt 2>&1; head -18 $O/sustained_nofold.txt
