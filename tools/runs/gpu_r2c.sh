# round-2 GPU pass C: where does the conv kernel's time go?  per-role timelines, fp16x3 vs fp16mx, and pipeline-role knock-outs
TAG=${1:-r2c}
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_mx.py -q --tb=short -p no:cacheprovider ) > gpurun_out/${TAG}_pytest_mx.log 2>&1; echo "pytest mx exit $?"
tail -15 gpurun_out/${TAG}_pytest_mx.log
{
for S in s2 s3 s4 s4b s5 c2 c3 c4 fpn stem2; do
  for P in fp16x3 fp16mx; do timeout 60 python tools/conv_timeline.py --shape $S --precision $P; done
done
echo "=== knock-outs (exp: 1 no epilogue work, 2 no TMA loads, 4 no MMAs)"
for S in s2 s4 c3; do
  for P in fp16x3 fp16mx; do for E in 1 2 4 3 5 6; do timeout 60 python tools/conv_timeline.py --shape $S --precision $P --exp $E | head -1; done; done
done
} > gpurun_out/${TAG}_conv_timeline.txt 2>&1
cat gpurun_out/${TAG}_conv_timeline.txt
