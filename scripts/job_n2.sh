mkdir -p gpurun_out
for W in s100 m; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload $W --steps 20 --warmup 3 > gpurun_out/bench_${W}_n2_final.json 2> gpurun_out/bench_${W}_n2_final.err
echo "rc=$? $W"; python -c "
import json; d=json.loads(open('gpurun_out/bench_${W}_n2_final.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['roofline']['frac'], d['parity'], (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('counts_match_golden'))"
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -2
