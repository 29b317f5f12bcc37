#!/bin/bash
# Multi-GPU pass on ONE box without the H2D probe: sharded inference bench + training step.  usage: scripts/gpu_multi_short.sh <N> <tag>
N=${1:-2}; tag=${2:-multi}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
(timeout 600 bash -c "$(declare -f run); N=$N; run 29511 bench.py --gpus $N --steps 10 --warmup 3" > gpurun_out/${tag}_bench_n${N}.json 2> gpurun_out/${tag}_bench_n${N}.err; echo "rc=$?" >> gpurun_out/${tag}_bench_n${N}.err)
(timeout 300 bash -c "$(declare -f run); N=$N; run 29512 bench.py --gpus $N --train --steps 20 --warmup 3" > gpurun_out/${tag}_train_n${N}.json 2> gpurun_out/${tag}_train_n${N}.err; echo "rc=$?" >> gpurun_out/${tag}_train_n${N}.err)
tail -c 200 gpurun_out/${tag}_bench_n${N}.err; tail -c 300 gpurun_out/${tag}_train_n${N}.err
