for CG in 0 1 2; do timeout 100 python tools/time_decoder_ops.py --cg $CG 2>&1 | grep "cg \|linear 1\|linear 900\|mha\|layernorm\|softmax"; done
timeout 100 python tools/time_decoder_ops.py --cg 0 --bn 32 2>&1 | grep "cg \|linear 1\|linear 900"
timeout 100 python tools/time_decoder_ops.py --cg 0 --bn 128 2>&1 | grep "cg \|linear 1\|linear 900"
