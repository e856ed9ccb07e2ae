#!/bin/bash
# One GPU round under gpurun: parity tests, bench line, ncu launch list, ncu --set full of the fused kernel.
# usage: scripts/gpu_round.sh <tag> [notests] [noncu]
TAG=${1:-r}
mkdir -p gpurun_out
if [[ "$*" != *notests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.txt
  cat gpurun_out/pytest_gpu_$TAG.txt
fi
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
if [[ "$*" != *noncu* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_project_cast_hist -s 3 -c 1 \
      -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_full_$TAG.log
fi
if [[ "$*" == *u8ncu* ]]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_hist_u8_cols -s 4 -c 1 \
      -o gpurun_out/prof_u8_$TAG -f python scripts/kbench.py 1000000 8 > gpurun_out/ncu_u8_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_u8_$TAG.log
fi
if [[ "$*" == *tmancu* ]]; then
  LOEXEC_TMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_project_cast_hist_tma -s 3 -c 1 \
      -o gpurun_out/prof_tma_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_tma_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_tma_$TAG.log
fi
