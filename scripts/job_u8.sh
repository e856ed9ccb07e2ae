timeout 300 python -m pytest tests -m gpu -x -q -k "host or memory_arrangement or smoke" 2>&1 | tail -2
timeout 200 python bench.py --workload m --steps 10 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['us_per_step'], d['roofline']['frac'], d['parity']['ok'], d['e2e']['value'], d['e2e']['h2d_GBs_slowest_rank'])"
