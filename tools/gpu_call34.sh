set -x
O=gpurun_out/r03k
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_multi.py tests/test_gpu_next.py -m gpu -q -x > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for c in 1 4 8 16; do echo "HPGB_HOST_MIN_CHUNKS=$c" >> $O/e2e_numpy.txt; HPGB_HOST_MIN_CHUNKS=$c timeout 300 python tools/e2e_numpy.py >> $O/e2e_numpy.txt 2>&1; done
cat $O/e2e_numpy.txt
