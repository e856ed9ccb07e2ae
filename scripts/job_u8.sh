timeout 900 python -m pytest tests -m gpu -x -q -k "parse or value_count or executors or number or columnar or group_by or hash" 2>&1 | tail -3
timeout 600 python scripts/aux_bench.py 2>&1 | tail -75
