TAG=${1:-r2w}
mkdir -p gpurun_out
FAR3D_NVCC_EXTRA=-DFAR3D_DA_PHASES python -c "
from far3d_b200 import build; build.build(force=True)" > /dev/null 2>&1; echo "rebuild exit $?"
{ timeout 100 python tools/agg_phases.py --nq 1047; timeout 100 python tools/agg_phases.py --nq 1047 --static-grid; timeout 100 python tools/agg_phases.py --nq 900 --static-grid; } > gpurun_out/${TAG}_agg_phases.txt 2>&1
cat gpurun_out/${TAG}_agg_phases.txt
