TAG=${1:-r3l}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "linear or mha_module" ) 2>&1 | tail -15
timeout 100 python tools/time_decoder_ops.py 2>&1 | grep "cg\|linear"
timeout 100 python tools/time_decoder_ops.py --umma 2>&1 | grep "cg\|linear"
