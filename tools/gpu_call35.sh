set -x
O=gpurun_out/r03l
mkdir -p $O
for sp in 0 500 2000; do for c in 1 8; do echo "HPGB_HOST_SPIN_US=$sp HPGB_HOST_MIN_CHUNKS=$c" >> $O/e2e_numpy.txt; HPGB_HOST_SPIN_US=$sp HPGB_HOST_MIN_CHUNKS=$c timeout 300 python tools/e2e_numpy.py 2>&1 | cut -c1-75 >> $O/e2e_numpy.txt; done; done
cat $O/e2e_numpy.txt
python tools/pool_gap.py > $O/pool_gap.txt 2>&1; HPGB_HOST_SPIN_US=0 python tools/pool_gap.py >> $O/pool_gap.txt 2>&1; cat $O/pool_gap.txt
