#!/bin/bash
# History-feature kernel: parity tests, CUDA-event timing against the HBM roofline, and (NCU=1) the ncu capture.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_format.py -x -q > gpurun_out/pytest_hist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_hist.log
timeout 200 python scripts/history_probe.py > gpurun_out/history_probe.log 2>&1 || { tail -5 gpurun_out/history_probe.log; exit 1; }
if [ "$NCU" = "1" ]; then
  timeout 300 ncu --set full --clock-control none -k regex:history_features -c 12 -f -o /tmp/prof_history \
    python scripts/history_probe.py --iters 1 > gpurun_out/ncu_history.log 2>&1
  ncu -i /tmp/prof_history.ncu-rep --page raw --csv > gpurun_out/prof_history.raw.csv 2>/dev/null
fi
tail -3 gpurun_out/pytest_hist.log; tail -1 gpurun_out/history_probe.log | cut -c1-400
