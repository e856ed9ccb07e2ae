#!/bin/bash
# 8-GPU box, config M only (1 M x 784 uint8): the byte histogram at N = 8 (peer merge, NCCL), 4, 2, 1, plus the group tests.
TAG=${1:-m}
mkdir -p gpurun_out
timeout 600 python -m pytest "tests/test_gpu_group.py" "tests/test_gpu_executors.py::test_rest_binned_histogram_shards_over_every_visible_gpu" -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_group_$TAG.txt
run() {  # name, ngpus, extra args...
  local name=$1 n=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 "$@" > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $n --steps 20 --warmup 3 "$@" > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err
  fi
  echo "rc=$? $name"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_${name}_$TAG.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','us_per_step','ms_per_step','n_gpus','parity')}, d['roofline'].get('frac'), (d.get('e2e') or {}).get('value'))"
}
run m_n8 8 --workload m
run m_n8_nccl 8 --workload m --merge nccl --no-e2e --no-cpu
run m_n4 4 --workload m --no-e2e --no-cpu
run m_n2 2 --workload m --no-e2e --no-cpu
run m_n1 1 --workload m --no-e2e --no-cpu
run s100_n8 8 --workload s100 --no-e2e --no-cpu
