# round-2 GPU pass E: ncu --set full of the conv kernel, layer classes s2 / s4b / c3 / fpn, fp16x3 vs fp16mx (raw pages as CSV)
TAG=${1:-r2e}
mkdir -p gpurun_out
for S in s2 s4b c3 fpn; do
  for P in fp16x3 fp16mx; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_persistent -s 1 -c 1 -f -o /tmp/${TAG}_${S}_${P} \
      python tools/prof_kernels.py conv --shape $S --precision $P --iters 1 > gpurun_out/${TAG}_${S}_${P}.log 2>&1
    echo "$S $P exit $?"; tail -1 gpurun_out/${TAG}_${S}_${P}.log
    ncu -i /tmp/${TAG}_${S}_${P}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${S}_${P}_raw.csv 2>/dev/null
  done
done
cp /tmp/${TAG}_s2_fp16mx.ncu-rep gpurun_out/
ls -la gpurun_out/ | tail -20
