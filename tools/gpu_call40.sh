set -x
N=${1:-8}
O=gpurun_out/r03q
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?" >> $O/pytest_multi.log; tail -3 $O/pytest_multi.log
for n in 8 4 2; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --gather > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err; tail -1 $O/bench_${n}gpu.json | cut -c1-300
  fi
done
