#!/bin/bash
# End-of-round validation on two GPUs: the NCCL tests (DP, row-sparse DP, vocabulary-parallel, model surface) and the
# N = 2 bench line with parity_check.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dp.py -x -q > gpurun_out/pytest_v_dp2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v_dp2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --no-cpu > gpurun_out/bench_v_n2.log 2> gpurun_out/bench_v_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/pytest_v_dp2.log
