set -x
O=gpurun_out/r02b
mkdir -p $O
timeout 120 tools/ubench/pipes2 > $O/pipes2.log 2>&1
for v in a2p_sweep a2p_sweep_magic a2p_sweep_queue a2p_sweep_mq; do timeout 300 tools/ubench/$v 30 10 2 > $O/$v.log 2>&1; done
for v in a2p_sweep_nomem a2p_sweep_nomem_mq; do timeout 200 tools/ubench/$v 30 10 1 > $O/$v.log 2>&1; done
timeout 200 tools/ubench/a2p_sweep_mq 28 20 1 > $O/a2p_sweep_mq_28x20.log 2>&1
timeout 200 tools/ubench/a2p_sweep_mq 28 200 1 > $O/a2p_sweep_mq_28x200.log 2>&1
timeout 200 tools/ubench/a2p_sweep_mq 30 40 1 > $O/a2p_sweep_mq_30x40.log 2>&1
ls -la $O
