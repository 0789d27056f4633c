mkdir -p gpurun_out
( timeout 100 python -m pytest tests/test_gpu_model.py -m gpu -q --tb=short -p no:cacheprovider -x -k "without_2d_proposals" ) > gpurun_out/r1k_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r1k_pytest.log
timeout 100 python bench.py --no-cpu-baseline > gpurun_out/r1k_bench.json 2> gpurun_out/r1k_bench.err; echo "bench exit $?"; head -c 600 gpurun_out/r1k_bench.json; echo; python -c "
import json; d=json.loads(open('gpurun_out/r1k_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['roofline']['traffic'], d['roofline']['frac'], d['roofline_deform_agg']['traffic'], d['roofline_deform_agg']['frac'])"; tail -3 gpurun_out/r1k_bench.err
