set -x
O=gpurun_out/r02n
mkdir -p $O
prof() {  # name kernel log2 regex
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1.source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
prof p2a p2a 27 k_stream
prof p2a_ring p2a_ring 27 k_stream
prof p2v p2v 27 k_stream
prof a2p a2p 27 k_stream
prof v2p v2p 27 'k_stream.*Vec2Pix'
prof upg_ring upg_ring 27 k_upgrade
prof neigh_ring neigh_ring 27 k_neighbors
head -50 $O/prof_p2a.summary.txt
