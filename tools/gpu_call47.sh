set -x
O=gpurun_out/r03x
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_s10.json 2> $O/bench_s10.err
sleep 3
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_s20.json 2> $O/bench_s20.err
python -c "
import json
for f in ('$O/bench_s10.json','$O/bench_s20.json'):
    d=json.load(open(f)); print(f, d['value'], d['roofline']['frac'], d['clocks']['reasons'], 'sustained', d['sustained']['value'], d['sustained']['roofline_frac'], 'e2e', d['e2e']['value'], 'v2p', d['also']['vector_to_pixel']['roofline_frac'])"
sleep 3
timeout 300 python tools/sustained_a2p.py 30 250 a2p > $O/sustained_a2p.txt 2>&1; head -18 $O/sustained_a2p.txt
timeout 200 python tools/ncu_headline.py 30 separate 1 > $O/plain_head.log 2>&1 && timeout 900 ncu --set full --clock-control none -k regex:k_stream -c 2 -f -o $O/prof_head python tools/ncu_headline.py 30 separate 1 > $O/ncu_head.log 2>&1
python tools/ncu_summary.py $O/prof_head.ncu-rep > $O/prof_head.summary.txt 2>&1; rm -f $O/prof_head.ncu-rep; head -20 $O/prof_head.summary.txt
timeout 600 python tools/quickbench.py 28 a2p,a2p_ring,n2r,r2n,p2a,p2v,v2p,neigh,interp,upgrade,bounds,lonlat > $O/quickbench.txt 2>&1; cat $O/quickbench.txt
