set -x
O=gpurun_out/r03c
mkdir -p $O
V=variants
for k in interp interp_ring; do
  timeout 300 python tools/ab_multi.py 28 $k 5 3 default $V/lib_interp_nokd.so > $O/ab_$k.log 2>&1; cat $O/ab_$k.log
done
prof() {  # name kernel log2 regex
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1.source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
prof neigh neigh 27 k_neighbors
cat $O/prof_neigh.summary.txt
