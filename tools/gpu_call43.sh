set -x
O=gpurun_out/r03t
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log; tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; python -c "
import json; d=json.load(open('$O/bench.json')); print(d['value'], d['roofline']['frac'], d['clocks'], d['sustained']['value'], d['e2e']['value'], d['gpu_launches'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-200
