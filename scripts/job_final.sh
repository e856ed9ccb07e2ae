#!/bin/bash
# Final 1-GPU round: full parity suite, smoke, the three bench workloads + the reference arm, launch lists, the (f) kernels,
# the shipped byte-histogram kernel under ncu, sanitizer.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for W in s100 m s10; do
  timeout 900 python bench.py --workload $W --steps 20 --warmup 3 > gpurun_out/bench_${W}_final.json 2> gpurun_out/bench_${W}_final.err
  echo "rc=$? $W"; python -c "
import json; d=json.loads(open('gpurun_out/bench_${W}_final.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['parity']['ok'], (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'), d.get('e2e_executor'))"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "rc=$? reference arm"; cut -c1-600 gpurun_out/bench_ref_final.json
timeout 300 python scripts/aux_bench.py > gpurun_out/aux_bench.log 2>&1; python -c "
import json; d=json.load(open('gpurun_out/aux_bench.json'))
for k,v in d.items(): print(k, {a: (round(b,3) if isinstance(b,float) else b) for a,b in v.items() if 'pinned' in a or 'kernel' in a})"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_m_final.csv python bench.py --workload m --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_s100_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_hist_u8_cols_lanes -s 4 -c 1 -o gpurun_out/prof_u8_final -f python bench.py --workload m --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_u8_final.log 2>&1; tail -1 gpurun_out/ncu_u8_final.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_project_cast_hist_bins -s 4 -c 1 -o gpurun_out/prof_bins_final -f python scripts/bins_bench.py > gpurun_out/ncu_bins_final.log 2>&1; tail -1 gpurun_out/ncu_bins_final.log
timeout 900 bash scripts/sanitize.sh 2>&1 | tail -8
