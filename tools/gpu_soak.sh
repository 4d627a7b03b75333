mkdir -p gpurun_out
python tools/soak_parity.py 64 500 300 0 > gpurun_out/soak_r01.jsonl
python tools/soak_parity.py 16 500 300 1 slot >> gpurun_out/soak_r01.jsonl
python tools/soak_parity.py 2048 15 200 0 >> gpurun_out/soak_r01.jsonl
python tools/soak_parity.py 4 8192 60 0 slot >> gpurun_out/soak_r01.jsonl
cat gpurun_out/soak_r01.jsonl
