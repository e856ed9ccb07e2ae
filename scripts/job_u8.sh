timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for R in 1 2 4 8; do echo "chunk rounds $R"; LOEXEC_U8_CHUNK_ROUNDS=$R timeout 300 python scripts/u8_sweep.py 11 2>&1 | tail -3; done
timeout 300 python scripts/u8_sweep.py 11 2>&1 | tail -3
for R in 1000000 125000; do timeout 200 python bench.py --workload m --rows $R --steps 20 --no-e2e --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rows', d['config']['rows'], 'us/step', d['us_per_step'], 'iso_ms', d['roofline']['kernel_ms_isolated'], 'frac', d['roofline']['frac'], d['parity'])"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_hist_u8_cols -s 4 -c 1 -o gpurun_out/prof_u8_f2 -f python bench.py --workload m --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_u8_f2.log 2>&1; tail -2 gpurun_out/ncu_u8_f2.log
