set -x
O=gpurun_out/r02m
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 600 python tools/quickbench.py 28 a2p,a2p_ring,n2r,r2n,p2a,p2v,v2p,neigh,interp,upgrade,bounds,lonlat > $O/quickbench.txt 2>&1; cat $O/quickbench.txt
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench1.json 2> $O/bench1.err; cat $O/bench1.json
