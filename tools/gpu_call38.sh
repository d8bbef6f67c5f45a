set -x
O=gpurun_out/r03o
mkdir -p $O
timeout 300 python tools/sustained_a2p.py 30 300 n2r > $O/sustained_n2r.txt 2>&1; head -40 $O/sustained_n2r.txt
nvidia-smi --query-gpu=power.limit,power.max_limit,power.default_limit,clocks.max.sm --format=csv > $O/power_limits.txt; cat $O/power_limits.txt
