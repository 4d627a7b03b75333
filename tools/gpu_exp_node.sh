mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 0 1; do MKF_POSE_CACHE=$c timeout 300 python tools/bench_node.py dropin 200 2>/dev/null | tail -1 | cut -c1-400; done
timeout 300 python tools/bench_call_latency.py 2>/dev/null | tail -2 | cut -c1-600
