set -x
O=gpurun_out/r02l
mkdir -p $O
V=variants
timeout 900 python -m pytest tests -m gpu -q -x -k "neighbors or upgrade or config4 or smoke" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for k in neigh neigh_ring; do timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_nb_direct256.so > $O/ab_$k.log 2>&1; cat $O/ab_$k.log; done
for k in upg upg_ring; do timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_upg_no256.so > $O/ab_$k.log 2>&1; cat $O/ab_$k.log; done
