set -x
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
V=variants
timeout 300 python tools/ab_multi.py 30 a2p 10 4 default $V/lib_a2p_v1.so $V/lib_a2p_v2q.so > $O/ab_a2p_30.log 2>&1
timeout 300 python tools/ab_multi.py 28 a2p 20 4 default $V/lib_a2p_v1.so $V/lib_a2p_v2q.so > $O/ab_a2p_28.log 2>&1
timeout 300 python tools/ab_multi.py 28 a2p_ring 20 3 default $V/lib_a2p_v1.so $V/lib_a2p_v2q.so > $O/ab_a2p_ring_28.log 2>&1
timeout 300 python tools/ab_multi.py 28 p2v 10 3 default $V/lib_p2v_old.so $V/lib_p2v_map_sd.so $V/lib_p2v_st_minb3.so > $O/ab_p2v_28.log 2>&1
timeout 300 python tools/ab_multi.py 28 v2p 10 3 default $V/lib_p2v_old.so > $O/ab_v2p_28.log 2>&1
ls -la $O
