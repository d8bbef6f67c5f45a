set -x
O=gpurun_out/r03f
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench1.json 2> $O/bench1.err; cat $O/bench1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cut -c1-400 $O/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_py.csv python bench.py --steps 2 --warmup 1 > $O/ncu_launches.log 2>&1
prof() {  # name kernel log2 regex [skip]
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:$4" -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  rm -f $O/prof_$1.ncu-rep
}
timeout 200 python tools/ncu_headline.py 30 separate 1 > $O/plain_head.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:k_stream -c 2 -f -o $O/prof_head python tools/ncu_headline.py 30 separate 1 > $O/ncu_head.log 2>&1
python tools/ncu_summary.py $O/prof_head.ncu-rep > $O/prof_head.summary.txt 2>&1; rm -f $O/prof_head.ncu-rep
prof r2n r2n 27 'k_stream.*OpReorder'
prof p2a p2a 27 'k_stream.*Pix2Ang'
prof p2a_ring p2a_ring 27 'k_stream.*Pix2Ang'
prof p2v p2v 27 'k_stream.*Pix2Vec'
prof v2p v2p 27 'k_stream.*Vec2Pix'
prof neigh neigh 27 k_neighbors
prof neigh_ring neigh_ring 27 k_neighbors
prof interp interp 27 k_interp
prof interp_ring interp_ring 27 k_interp
prof upg upg 27 k_upgrade
prof upg_ring upg_ring 27 k_upgrade
timeout 600 python tools/quickbench.py 28 a2p,a2p_ring,n2r,r2n,p2a,p2v,v2p,neigh,interp,upgrade,bounds,lonlat > $O/quickbench.txt 2>&1; cat $O/quickbench.txt
grep -c . $O/prof_*.summary.txt
