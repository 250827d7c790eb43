set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python __graft_entry__.py --smoke 2>&1 | tail -5
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python - <<'PY'
from zfit_b200 import _znll as Z
for ms in (50, 500, 3000):
    print('fp64 peak TFLOP/s over', ms, 'ms:', Z.fp64_peak(0, ms))
PY
