set -x
O=gpurun_out/r02p
mkdir -p $O
prof() {  # name kernel log2 regex count
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1.source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
export HPGB_AB_NSIDE=1048576
prof p2a_tab20 p2a 27 k_stream
prof p2a_build20 p2a 27 k_ringtab
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 20 --csv --log-file $O/launches20.csv python tools/ncu_one.py p2a 27 1 > /dev/null 2>&1
export HPGB_AB_NSIDE=4096
prof p2a_tab4096 p2a 27 k_stream
prof p2v_tab4096 p2v 27 k_stream
cat $O/prof_p2a_tab20.summary.txt $O/prof_p2a_build20.summary.txt
