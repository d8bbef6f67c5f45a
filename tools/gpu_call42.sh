set -x
O=gpurun_out/r03s
mkdir -p $O
V=variants
for k in interp interp_ring; do
  timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_interp_b768.so $V/lib_interp_b704.so $V/lib_interp_b512.so > $O/ab_$k.log 2>&1; cat $O/ab_$k.log
done
