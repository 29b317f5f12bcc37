#!/bin/bash
# A/B of the persistent LSTM recurrence variants (tmrnet_b200/variants/*.so, built by hand for the measurement)
tag=${1:-ab}
mkdir -p gpurun_out
for v in tmrnet_b200/variants/*.so; do
  timeout 200 python scripts/lstm_step_time.py 82944 --lib $v --check >> gpurun_out/${tag}_lstm_ab.log 2>&1
done
timeout 200 python scripts/lstm_step_time.py 82944 --check >> gpurun_out/${tag}_lstm_ab.log 2>&1
cat gpurun_out/${tag}_lstm_ab.log | grep -v Warn
