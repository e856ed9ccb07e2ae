timeout 900 python -m pytest tests -m gpu -x -q -k "nbins or group" 2>&1 | tail -3
timeout 300 python scripts/bins_bench.py 2>&1 | tail -9
