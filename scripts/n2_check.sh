timeout 600 python -m pytest tests/test_gpu_sharded_workloads.py -q 2>&1 | grep -v Warning | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 600 gpurun_out/bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'step ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
print(json.dumps(d.get('scan'))); print(json.dumps(d.get('sweep')))
"
