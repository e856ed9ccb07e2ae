timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/u8_sweep.py 7 11 12 2>&1 | tail -9; cp gpurun_out/u8_sweep.json gpurun_out/u8_sweep_final.json
timeout 300 python bench.py --workload m --steps 20 > gpurun_out/bench_m_n1.json 2> gpurun_out/bench_m_n1.err; tail -c 1500 gpurun_out/bench_m_n1.json
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/atoms_probe scripts/probes/atoms_probe.cu && /tmp/atoms_probe > gpurun_out/atoms_probe.jsonl; cat gpurun_out/atoms_probe.jsonl
timeout 600 bash scripts/sanitize.sh 2>&1 | tail -8
