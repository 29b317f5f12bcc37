#!/bin/bash
# ncu --set full captures of the two dominant kernels + the persistent LSTM's per-item timeline (experiment build).
# usage: scripts/gpu_prof.sh <tag> [bankconv|persist|timeline ...]
tag=${1:-prof}; shift
what=${@:-bankconv persist timeline}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    bankconv)
      timeout 120 python scripts/bankconv_time.py 83000 > gpurun_out/${tag}_bankconv_time.log 2>&1 &&
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_bankconv -s 3 -c 1 \
        -o gpurun_out/${tag}_bankconv -f python scripts/bankconv_time.py 83000 > gpurun_out/${tag}_bankconv_ncu.log 2>&1 ;;
    persist)
      timeout 120 python scripts/lstm_step_time.py 82944 > gpurun_out/${tag}_lstm_time.log 2>&1 &&
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:lstm_persist -s 3 -c 1 \
        -o gpurun_out/${tag}_persist -f python scripts/lstm_step_time.py 82944 > gpurun_out/${tag}_persist_ncu.log 2>&1 ;;
    timeline)
      timeout 120 python scripts/persist_timeline.py > gpurun_out/${tag}_persist_timeline.log 2>&1 ;;
    attention)
      timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_pb -s 2 -c 1 \
        -o gpurun_out/${tag}_attention -f python bench.py --steps 1 --warmup 3 --no-cpu --no-eager > gpurun_out/${tag}_attention_ncu.log 2>&1 ;;
  esac
done
tail -2 gpurun_out/${tag}_*time*.log
