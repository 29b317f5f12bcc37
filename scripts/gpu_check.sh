#!/bin/bash
# One GPU-box pass: GPU tests, smoke(), the default bench line, then the ncu launch list of a short bench run.
# usage: scripts/gpu_check.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-chk}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt
(timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log)
(timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${tag}_smoke.log)
(timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "rc=$?" >> gpurun_out/${tag}_bench.err)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-eager > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log; tail -2 gpurun_out/${tag}_smoke.log; tail -c 600 gpurun_out/${tag}_bench.err
