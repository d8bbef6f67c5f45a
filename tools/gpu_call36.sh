set -x
O=gpurun_out/r03m
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench1.json 2> $O/bench1.err; cat $O/bench1.json
timeout 900 python bench.py > $O/bench_default.json 2> $O/bench_default.err; cut -c1-300 $O/bench_default.json
timeout 600 python tools/quickbench.py 28 a2p,a2p_ring,n2r,r2n,p2a,p2v,v2p,neigh,interp,upgrade,bounds,lonlat > $O/quickbench.txt 2>&1; cat $O/quickbench.txt
prof() {  # name kernel log2 regex
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$4" -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  rm -f $O/prof_$1.ncu-rep
}
timeout 300 python tools/ab_multi.py 28 p2a 5 3 default variants/lib_p2a_selparts.so > $O/ab_p2a.log 2>&1; cat $O/ab_p2a.log
timeout 300 python tools/ab_multi.py 28 p2a_ring 5 3 default variants/lib_p2a_selparts.so > $O/ab_p2a_ring.log 2>&1; cat $O/ab_p2a_ring.log
prof v2p v2p 27 'k_stream.*Vec2Pix'
cat $O/prof_v2p.summary.txt
