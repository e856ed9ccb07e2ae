timeout 600 python -m pytest tests -m gpu -x -q -k "every_kernel_variant" 2>&1 | tail -2
timeout 600 python scripts/u8_sweep.py 11 13 14 11 14 2>&1 | tail -15
