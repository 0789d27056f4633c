TAG=${1:-r4z}
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/${TAG}_bench_2gpu.json 2> gpurun_out/${TAG}_bench_2gpu.err; echo "bench 2 exit $?"
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 ) > gpurun_out/${TAG}_bench_reference_2gpu.json 2> gpurun_out/${TAG}_bench_reference_2gpu.err; echo "ref 2 exit $?"
python - <<PY
import json
for f in ('${TAG}_bench_2gpu.json', '${TAG}_bench_reference_2gpu.json'):
    try:
        ls=[l for l in open('gpurun_out/'+f).read().strip().splitlines() if l.startswith('{')]
        d=json.loads(ls[-1])
        print(f, 'json lines', len(ls), 'n_gpus', d['n_gpus'], 'value', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'workload', d['config']['workload'][:60])
        if d.get('strong_scaling'): print('   strong', json.dumps(d['strong_scaling'])[:500])
    except Exception as e:
        print(f, 'parse failed', e)
PY
