TAG=${1:-r4f}
mkdir -p gpurun_out
( echo "== epi8"; timeout 200 python tools/adaptive_bench.py
  echo "== epi4"; FAR3D_LIB_PATH=$PWD/far3d_b200/lib/libfar3d_sm100_epi4.so timeout 200 python tools/adaptive_bench.py
  echo "== epi8 reserve 32768"; timeout 200 python tools/adaptive_bench.py --reserve 32768 ) > gpurun_out/${TAG}_adaptive.txt 2>&1
cat gpurun_out/${TAG}_adaptive.txt
