set -x
N=${1:-8}
O=gpurun_out/r02u
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log; tail -3 $O/pytest_multi.log
timeout 900 python tools/pcie_peaks.py --gpus $N > $O/pcie_peaks_${N}gpu.log 2>&1; tail -20 $O/pcie_peaks_${N}gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 --gather > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err; tail -c 3000 $O/bench_${N}gpu.json
if [ "$N" = "8" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 10 --warmup 3 --gather > $O/bench_4gpu.json 2> $O/bench_4gpu.err
fi
