set -x
O=gpurun_out/r02i
mkdir -p $O
V=variants
timeout 300 python tools/ab_multi.py 28 r2n 10 3 default $V/lib_r2n_m3.so $V/lib_r2n_m5.so $V/lib_r2n_m5s2.so $V/lib_r2n_m4s2.so $V/lib_r2n_m4s4.so > $O/ab_r2n_28b.log 2>&1
cat $O/ab_r2n_28b.log
prof() {  # name kernel log2 regex
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1.source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
prof r2n r2n 27 k_stream
cat $O/prof_r2n.summary.txt
