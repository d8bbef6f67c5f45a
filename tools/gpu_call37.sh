set -x
O=gpurun_out/r03n
mkdir -p $O
timeout 300 python tools/sustained_a2p.py 30 300 a2p > $O/sustained_a2p.txt 2>&1; cat $O/sustained_a2p.txt
timeout 300 python tools/sustained_a2p.py 30 300 copy > $O/sustained_copy.txt 2>&1; head -24 $O/sustained_copy.txt
