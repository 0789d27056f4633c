TAG=${1:-r2l}
NG=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
( time timeout 900 python -m pytest tests/test_gpu_multi.py -q --tb=short -p no:cacheprovider -s ) > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi exit $?"
tail -6 gpurun_out/${TAG}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus $NG --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${NG}gpu.json 2> gpurun_out/${TAG}_bench_${NG}gpu.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_${NG}gpu.json').read().strip().splitlines()[-1])
    print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'e2e_u8', d['e2e_uint8'] and round(d['e2e_uint8']['value'],2), 'latency', d.get('latency_ms_unpipelined'))
    print('strong', json.dumps(d.get('strong_scaling'), indent=1))
    print('adaptive', d.get('streaming_adaptive'))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_${NG}gpu.err').read()[-3000:])
PY
