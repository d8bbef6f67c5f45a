set -x
O=gpurun_out/r02o
mkdir -p $O
V=variants
timeout 900 python -m pytest tests -m gpu -q -x -k "angle or vector or config2 or smoke or parity" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
export HPGB_AB_NSIDE=1048576
for k in p2a p2a_ring p2v; do
  timeout 300 python tools/ab_multi.py 27 $k 5 3 default $V/lib_p2_notab.so $V/lib_p2v_plain.so $V/lib_p2t_minb4.so $V/lib_p2t_h1.so $V/lib_p2t_st4.so > $O/ab_$k.log 2>&1; cat $O/ab_$k.log
done
export HPGB_AB_NSIDE=4096
for k in p2a p2v; do
  timeout 300 python tools/ab_multi.py 27 $k 5 3 default $V/lib_p2_notab.so > $O/ab4096_$k.log 2>&1; cat $O/ab4096_$k.log
done
