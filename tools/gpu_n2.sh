mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_r01_n2.json 2> gpurun_out/bench_r01_n2.err
tail -c 1500 gpurun_out/bench_r01_n2.json; tail -5 gpurun_out/bench_r01_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 | cut -c1-300
