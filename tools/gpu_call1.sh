set -x
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > $O/smi.txt
timeout 300 python tools/quickbench.py 28 a2p,a2p_ring,n2r,r2n,p2a,p2v,v2p,neigh,interp,upgrade,bounds,lonlat > $O/quick28.log 2>&1
for v in a2p_sweep a2p_sweep_magic a2p_sweep_queue a2p_sweep_mq; do timeout 300 tools/ubench/$v 30 10 3 > $O/$v.log 2>&1; done
for v in a2p_sweep_nomem a2p_sweep_nomem_mq; do timeout 200 tools/ubench/$v 30 10 1 > $O/$v.log 2>&1; done
timeout 200 tools/ubench/a2p_sweep 28 20 1 > $O/a2p_sweep_28.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1
prof() {  # name log2 layout import-source
  timeout 200 python tools/ncu_headline.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 900 ncu --set full --clock-control none $4 -k regex:k_stream -c 2 -f -o $O/prof_$1 python tools/ncu_headline.py $2 $3 1 > $O/ncu_$1.log 2>&1
  ncu -i $O/prof_$1.ncu-rep --page raw --csv > $O/prof_$1.raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
}
prof sep30 30 separate "--import-source on"
ncu -i $O/prof_sep30.ncu-rep --page source --csv > $O/prof_sep30.source.csv 2>/dev/null
prof carved30 30 carved ""
prof sep28 28 separate ""
rm -f $O/prof_carved30.ncu-rep $O/prof_sep28.ncu-rep
ls -la $O; du -sh gpurun_out
