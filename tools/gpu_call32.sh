set -x
O=gpurun_out/r03i
mkdir -p $O
prof() {  # name kernel log2 regex
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$4" -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1.source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
prof v2p v2p 27 'k_stream.*Vec2Pix'
prof a2p_ring a2p_ring 27 'k_stream.*Ang2Pix'
prof a2p a2p 27 'k_stream.*Ang2Pix'
