set -x
O=gpurun_out/r03j
mkdir -p $O
nproc > $O/host.txt; lscpu | grep -E "Model name|Socket|Thread|Core|L3|NUMA node\(s\)" >> $O/host.txt
for t in 8 16 24 32; do HPGB_HOST_THREADS=$t python tools/host_copy_rate.py 1024 5 >> $O/host_copy.txt 2>&1; done
cat $O/host_copy.txt
for t in 8 16 24 32; do echo "HPGB_HOST_THREADS=$t" >> $O/e2e_numpy.txt; HPGB_HOST_THREADS=$t timeout 300 python tools/e2e_numpy.py >> $O/e2e_numpy.txt 2>&1; done
cat $O/e2e_numpy.txt
