timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "u8" 2>&1 | tail -3
timeout 300 python scripts/u8_sweep.py 7 11 2>&1 | tail -7
for R in 1000000 125000; do timeout 200 python bench.py --workload m --rows $R --steps 20 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rows', d['config']['rows'], 'us/step', d['us_per_step'], 'iso_ms', d['roofline']['kernel_ms_isolated'], 'frac', d['roofline']['frac'], d['parity'])"; done
