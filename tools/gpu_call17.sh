set -x
O=gpurun_out/r02r
mkdir -p $O
V=variants
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for k in p2a p2a_ring p2v; do
  timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_p2_notab.so $V/lib_p2v_plain.so > $O/ab_$k.log 2>&1; grep -v "result ==" $O/ab_$k.log
done
for ns in 1048576 4096; do
export HPGB_AB_NSIDE=$ns
for k in p2a p2a_ring p2v; do
  timeout 300 python tools/ab_multi.py 27 $k 5 3 default $V/lib_p2_notab.so $V/lib_p2v_plain.so > $O/ab_${ns}_$k.log 2>&1; grep -v "result ==" $O/ab_${ns}_$k.log
done
done
unset HPGB_AB_NSIDE
timeout 300 python tools/quickbench.py 27 interp,bounds > $O/qb.txt 2>&1; cat $O/qb.txt
