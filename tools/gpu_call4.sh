set -x
O=gpurun_out/r02d
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/pcie_peaks.py --gpus 2 > $O/pcie_peaks_2gpu.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --gather > $O/bench_2gpu.log 2>&1
ls -la $O
