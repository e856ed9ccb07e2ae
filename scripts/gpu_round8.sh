#!/bin/bash
# The 8-GPU round (gpurun --gpus 8): local 8-device group tests, the driver's bench command at N = 8, its variants,
# N = 4 / 2 / 1 on the same box for the scaling table, config M, and the CPU arm.
TAG=${1:-r}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
if [[ "$*" != *notests* ]]; then
timeout 600 python -m pytest "tests/test_gpu_group.py::test_local_group_over_all_visible_devices" \
    "tests/test_gpu_executors.py::test_rest_binned_histogram_shards_over_every_visible_gpu" \
    "tests/test_gpu_group.py::test_rank_group_nccl_merge_equals_oracle" -x -q 2>&1 | tail -6 | tee gpurun_out/pytest_group_$TAG.txt
fi
run() {  # name, ngpus, extra args...
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 "$@" > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err
  fi
  echo "rc=$? $name"; cut -c1-2600 gpurun_out/bench_${name}_$TAG.json; grep -v "^W0\|^\*\*\*\|OMP_NUM\|NCCL version" gpurun_out/bench_${name}_$TAG.err | tail -3
}
run s100_n8 8 --workload s100
run s100_n8_nooverlap 8 --workload s100 --no-overlap --no-e2e
run s100_n8_nccl 8 --workload s100 --merge nccl --no-e2e
run m_n8 8 --workload m
run m_n8_nccl 8 --workload m --merge nccl --no-e2e
run s100_n4 4 --workload s100 --no-e2e
run s100_n2 2 --workload s100 --no-e2e
run s100_n1 1 --workload s100 --no-e2e --no-cpu --executor-rows 0
run m_n1 1 --workload m --no-e2e --no-cpu
if [[ "$*" != *noref* ]]; then
timeout 300 python bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
echo "rc=$? reference arm"; cut -c1-900 gpurun_out/bench_ref_$TAG.json
fi
