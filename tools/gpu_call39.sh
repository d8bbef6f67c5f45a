set -x
O=gpurun_out/r03p
mkdir -p $O
for v in default a2p_hint2us a2p_hint20us; do
  if [ $v = default ]; then unset HPGEOM_B200_LIB; else export HPGEOM_B200_LIB=$PWD/variants/lib_$v.so; fi
  timeout 300 python tools/sustained_a2p.py 30 250 a2p > $O/sustained_$v.txt 2>&1; head -18 $O/sustained_$v.txt
  sleep 2
done
unset HPGEOM_B200_LIB
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_s20.json 2> $O/bench_s20.err; python -c "
import json; d=json.load(open('$O/bench_s20.json')); print(d['value'], d['roofline']['frac'], d['clocks']); print(d['sustained'])"
