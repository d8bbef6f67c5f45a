set -x
O=gpurun_out/r02h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_reorder.py tests/test_gpu_parity_report.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
V=variants
timeout 300 python tools/ab_multi.py 28 r2n 10 3 default $V/lib_r2n_v1.so > $O/ab_r2n_28.log 2>&1
timeout 300 python tools/ab_multi.py 30 r2n 5 3 default $V/lib_r2n_v1.so > $O/ab_r2n_30.log 2>&1
prof() {  # name kernel log2 regex
  timeout 200 python tools/ncu_one.py $2 $3 1 > $O/plain_$1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$4 -c 1 -f -o $O/prof_$1 python tools/ncu_one.py $2 $3 1 > $O/ncu_$1.log 2>&1
  python tools/ncu_summary.py $O/prof_$1.ncu-rep > $O/prof_$1.summary.txt 2>&1
  ncu -i $O/prof_$1.ncu-rep --page source --csv > $O/prof_$1.source.csv 2>/dev/null
  rm -f $O/prof_$1.ncu-rep
}
prof r2n r2n 27 k_stream
prof p2a p2a 27 k_stream
prof p2v p2v 27 k_stream
cp gpurun_out/parity_report.json $O/ 2>/dev/null
ls -la $O
