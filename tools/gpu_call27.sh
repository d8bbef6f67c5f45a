set -x
O=gpurun_out/r03b
mkdir -p $O
V=variants
timeout 600 python -m pytest tests -m gpu -q -x -k "neigh" > $O/pytest_neigh.log 2>&1; tail -3 $O/pytest_neigh.log
for k in neigh neigh_ring; do
  timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_nb_reg.so $V/lib_nb_d3.so > $O/ab_$k.log 2>&1; cat $O/ab_$k.log
done
