( timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --tb=short -p no:cacheprovider -k "mha" ) 2>&1 | tail -25
timeout 100 python tools/time_decoder_ops.py 2>&1 | grep "mha"
