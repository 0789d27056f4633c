TAG=${1:-r3e}
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/${TAG}_bench_8gpu.json 2> gpurun_out/${TAG}_bench_8gpu.err; echo "bench 8 exit $?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/${TAG}_bench_8gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
    print('N=8 value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'u8', d['e2e_uint8'] and round(d['e2e_uint8']['value'],2), 'scaling', d['scaling'])
    print('strong', json.dumps(d.get('strong_scaling'))[:900])
    print('clocks', d.get('clocks'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_8gpu.err').read()[-3000:])
PY
( time timeout 300 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider -m gpu ) > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"; tail -4 gpurun_out/${TAG}_pytest_multi.log
