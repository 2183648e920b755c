#!/bin/bash
# Is the slow cfg2 logits kernel a property of the multi-GPU box (one process alone) or of concurrent processes?
mkdir -p gpurun_out
nvidia-smi -q -d POWER,CLOCK,PERFORMANCE -i 0 > gpurun_out/box_smi_q.txt 2>&1
nvidia-smi --query-gpu=index,name,power.limit,power.default_limit,clocks.max.sm,compute_mode,mig.mode.current --format=csv > gpurun_out/box_smi.csv 2>&1
nvidia-smi topo -m > gpurun_out/box_topo.txt 2>&1
nproc > gpurun_out/box_nproc.txt
CUDA_VISIBLE_DEVICES=0 python bench.py --config cfg2_reddit_gru128 --no-cpu --no-sub --steps 20 > gpurun_out/box_single.log 2> gpurun_out/box_single.err
for i in 0 1 2 3; do
  CUDA_VISIBLE_DEVICES=$i python bench.py --config cfg2_reddit_gru128 --no-cpu --no-sub --steps 200 > gpurun_out/box_conc_$i.log 2> gpurun_out/box_conc_$i.err &
done
wait
echo done
