#!/bin/bash
# End-of-round validation on one GPU: GPU suite, smoke, scan wave probe, default bench (own arm).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_v.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_v.log 2>&1
timeout 120 python scripts/scan_waves.py > gpurun_out/scan_waves.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_v.log 2> gpurun_out/bench_v.err; echo "bench rc=$?"
tail -3 gpurun_out/pytest_v.log; cat gpurun_out/smoke_v.log | tail -2; cat gpurun_out/scan_waves.log | tail -3
