#!/bin/bash
# One multi-GPU round under gpurun --gpus N: group tests across real peers, bench at N for s100 / m (peer and nccl merge).
TAG=${1:-r}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
if [[ "$*" != *notests* ]]; then
  timeout 900 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_group_$TAG.txt
fi
run() {  # name, extra args...
  local name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err
  echo "rc=$? $name"; cat gpurun_out/bench_${name}_$TAG.json; grep -v "^W0\|^\*\*\*\|OMP_NUM" gpurun_out/bench_${name}_$TAG.err | tail -4
}
run s100_n$N --workload s100
run s100_n${N}_nccl --workload s100 --merge nccl --no-e2e
run m_n$N --workload m
run m_n${N}_nccl --workload m --merge nccl --no-e2e
if [[ "$*" == *u8sweep* ]]; then timeout 300 python scripts/u8_sweep.py 2>&1 | tail -14; fi
if [[ "$*" == *ab* ]]; then
  timeout 600 python scripts/ab_libs.py $PWD/learningorchestra_b200/lib/libloexec.so $PWD/learningorchestra_b200/lib/libloexec_static.so $PWD/learningorchestra_b200/lib/libloexec_r1.so 2>&1 | tail -8
fi
