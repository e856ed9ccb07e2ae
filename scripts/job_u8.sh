timeout 600 python -m pytest tests -m gpu -x -q -k "host or executors or group" 2>&1 | tail -3
timeout 300 python bench.py --workload m --steps 20 > gpurun_out/bench_m_n1.json 2> gpurun_out/bench_m_n1.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_m_n1.json').read().strip().splitlines()[-1]); print('m', d['value'], d['us_per_step'], d['roofline']['frac'], d['e2e'], d['cpu_baseline']['value'])"
timeout 600 python bench.py --workload s100 --steps 10 --no-cpu --executor-rows 0 > gpurun_out/bench_s100_e2e.json 2> gpurun_out/bench_s100_e2e.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_s100_e2e.json').read().strip().splitlines()[-1]); print('s100', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'])"
