mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python tools/bench_configs.py 2lit
python tools/soak_parity.py 16 500 120 1 slot
