#!/bin/bash
# 8-GPU evidence (gpurun --gpus 8): the bench contract line at N=8, config 4 (fused op fwd+bwd under DDP, and the
# whole network's training step under DDP) at N=1 and N=8 on the same box, and K2b sharded over the reference views
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "bench n8 exit $?"
tail -c 400 gpurun_out/bench_n8.err
: > gpurun_out/ddp_n8.log
timeout 300 python scripts/train_ddp_demo.py >> gpurun_out/ddp_n8.log 2>gpurun_out/ddp_n8.err; echo "op n1 exit $?"
timeout 300 $TR --nproc-per-node 8 --master-port 29612 scripts/train_ddp_demo.py >> gpurun_out/ddp_n8.log 2>>gpurun_out/ddp_n8.err; echo "op n8 exit $?"
timeout 300 python scripts/train_ddp_demo.py --full >> gpurun_out/ddp_n8.log 2>>gpurun_out/ddp_n8.err; echo "full n1 exit $?"
timeout 300 $TR --nproc-per-node 8 --master-port 29613 scripts/train_ddp_demo.py --full >> gpurun_out/ddp_n8.log 2>>gpurun_out/ddp_n8.err; echo "full n8 exit $?"
cat gpurun_out/ddp_n8.log; tail -3 gpurun_out/ddp_n8.err
: > gpurun_out/filter_n8.log
timeout 300 python scripts/filter_sharded.py >> gpurun_out/filter_n8.log 2>gpurun_out/filter_n8.err; echo "filter n1 exit $?"
timeout 300 $TR --nproc-per-node 2 --master-port 29614 scripts/filter_sharded.py >> gpurun_out/filter_n8.log 2>>gpurun_out/filter_n8.err; echo "filter n2 exit $?"
timeout 300 $TR --nproc-per-node 8 --master-port 29615 scripts/filter_sharded.py >> gpurun_out/filter_n8.log 2>>gpurun_out/filter_n8.err; echo "filter n8 exit $?"
cat gpurun_out/filter_n8.log; tail -3 gpurun_out/filter_n8.err
