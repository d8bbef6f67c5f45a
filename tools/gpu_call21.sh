set -x
O=gpurun_out/r02v
mkdir -p $O
prof() {  # name kernel log2 regex
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1.source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
prof interp interp 27 k_interp
prof p2a p2a 27 k_stream
prof p2v p2v 27 k_stream
prof neigh_ring neigh_ring 27 k_neighbors
prof upg_ring upg_ring 27 k_upgrade
ls -la $O
