#!/bin/bash
# Round-2 evidence run (one GPU, one gpurun call): GPU test suite, plain bench (own arm and the reference arm), then the
# ncu launch lists and --set full captures of the hot kernels.  Every ncu pass comes after the plain run has exited 0.
# The .ncu-rep files are turned into raw CSV pages on the box and deleted (gpurun brings back at most 64 MiB).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log
python bench.py > gpurun_out/bench_r2_full.log 2> gpurun_out/bench_r2_full.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.log 2> gpurun_out/bench_r2_reference.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_r2.log 2>&1
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 700 --csv --log-file gpurun_out/launches_r2_cfg4.csv \
  python bench.py --config cfg4_gru256_1m --steps 2 --warmup 3 --no-cpu --no-sub > gpurun_out/ncu_l_cfg4.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file gpurun_out/launches_r2b_cfg2.csv \
  python bench.py --config cfg2_reddit_gru128 --steps 2 --warmup 3 --no-cpu --no-sub > gpurun_out/ncu_l_cfg2b.log 2>&1
full() {  # name, kernel regex, count, config, ncu selection
  $NCU $5 -k "regex:$2" -c $3 -f -o /tmp/$1 \
    python bench.py --config $4 --steps 1 --warmup 3 --no-cpu --no-sub > gpurun_out/ncu_f_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  ls -la /tmp/$1.ncu-rep
  rm -f /tmp/$1.ncu-rep
}
# cfg4's kernels run 80-90 ms: the SASS-patching passes of --set full exceed the kernels' own 2 s mbarrier watchdog on the
# second launch, so the dW kernel gets the hardware-counter sections only
full prof3_ce_cfg4 "ce_tc_backward_ts" 1 cfg4_gru256_1m "--set full"
LIGHT="--section SpeedOfLight --section MemoryWorkloadAnalysis --section ComputeWorkloadAnalysis --section LaunchStats --section Occupancy"
full prof3_dw_cfg4 "ce_tc_backward_ts" 2 cfg4_gru256_1m "$LIGHT"
full prof3_ce_cfg3 ce_tc_backward_ts 2 cfg3_lstm256_50k "--set full"
full prof3_cfg3_rnn rnn_tc 2 cfg3_lstm256_50k "--set full"
full prof3_cfg2 "ce_tc_backward_ts|rnn_" 4 cfg2_reddit_gru128 "--set full"
full prof3_cfg1 "lstm_" 2 cfg1_msnbc_lstm100 "--set full"
du -sh gpurun_out
echo done
