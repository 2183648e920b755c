set -x
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_configs.py -m gpu -q > gpurun_out/pytest_n.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_n.log
for c in cfg4_gru256_1m cfg3_lstm256_50k; do timeout 200 python bench.py --config $c --no-cpu --no-sub --steps 10 > gpurun_out/bench_n_$c.log 2> gpurun_out/bench_n_$c.err; done
LIGHT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy"
ncu --clock-control none $LIGHT -k "regex:ce_tc_backward_ts" -c 2 -f -o /tmp/prof4 python bench.py --config cfg4_gru256_1m --steps 1 --warmup 3 --no-cpu --no-sub > gpurun_out/ncu_f_prof4.log 2>&1
ncu -i /tmp/prof4.ncu-rep --page raw --csv > gpurun_out/prof4_dw_cfg4.raw.csv 2>/dev/null
