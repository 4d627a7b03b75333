mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 200 --warmup 5 > gpurun_out/bench_r01_n$N.json 2> gpurun_out/bench_r01_n$N.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_r01_n$N.json').read()); print('N',d['n_gpus'],'value',d['value'],'ms/step',d['ms_per_step'],'e2e',d['e2e']['value'],'rows',d['gathered_rows'],'frac',d['roofline']['frac'])"
tail -3 gpurun_out/bench_r01_n$N.err
