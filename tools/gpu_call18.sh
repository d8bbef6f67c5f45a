set -x
O=gpurun_out/r02s
mkdir -p $O
V=variants
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
for k in p2a p2a_ring; do
  timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_p2a_v1.so $V/lib_p2a_magic.so $V/lib_p2a_v1magic.so > $O/ab_$k.log 2>&1; grep -v "result ==" $O/ab_$k.log
done
timeout 300 python tools/ab_multi.py 28 p2v 5 3 default $V/lib_p2a_magic.so $V/lib_p2a_plain.so $V/lib_p2a_plainmagic.so > $O/ab_p2v.log 2>&1; grep -v "result ==" $O/ab_p2v.log
timeout 600 python tools/quickbench.py 28 a2p,a2p_ring,n2r,r2n,p2a,p2v,v2p,neigh,interp,upgrade,bounds,lonlat > $O/quickbench.txt 2>&1; cat $O/quickbench.txt
