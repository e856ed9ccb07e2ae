#!/bin/bash
# One 1-GPU round under gpurun: parity tests, bench lines of the three workloads, ncu launch list + --set full capture.
# usage: scripts/gpu_round.sh <tag> [notests] [noncu] [sweep] [u8ncu]
TAG=${1:-r}
mkdir -p gpurun_out
if [[ "$*" != *notests* ]]; then
  timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu_$TAG.txt
  cat gpurun_out/pytest_gpu_$TAG.txt
fi
for W in s100 m s10; do
  timeout 600 python bench.py --workload $W --steps 20 --warmup 3 > gpurun_out/bench_${W}_$TAG.json 2> gpurun_out/bench_${W}_$TAG.err
  echo "rc=$?"; cat gpurun_out/bench_${W}_$TAG.json; tail -3 gpurun_out/bench_${W}_$TAG.err
done
if [[ "$*" == *sweep* ]]; then
  timeout 600 python scripts/tile_sweep.py 2>&1 | tail -50
fi
if [[ "$*" == *ab* ]]; then timeout 600 python scripts/ab_libs.py $PWD/learningorchestra_b200/lib/libloexec.so $PWD/learningorchestra_b200/lib/libloexec_r1.so 2>&1 | tail -4; timeout 600 python scripts/ab_libs.py $PWD/learningorchestra_b200/lib/libloexec.so $PWD/learningorchestra_b200/lib/libloexec_r1.so 12500000 2>&1 | tail -2; fi
if [[ "$*" != *noncu* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
      --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_project_cast_hist -s 3 -c 1 \
      -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_$TAG.log 2>&1
  tail -3 gpurun_out/ncu_full_$TAG.log
fi
if [[ "$*" == *u8ncu* ]]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_hist_u8_cols -s 4 -c 1 \
      -o gpurun_out/prof_u8_$TAG -f python bench.py --workload m --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_u8_$TAG.log 2>&1
  tail -2 gpurun_out/ncu_u8_$TAG.log
fi
