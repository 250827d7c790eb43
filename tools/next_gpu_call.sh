#!/usr/bin/env bash
# What the next GPU session should run FIRST (round 1 ended with the GPU budget spent; everything below is built and
# CPU-checked but not yet measured on a B200).  Two steps: build the variant library here (nvcc cross-compiles), then one
# gpurun call.  Every leg writes into gpurun_out/ and is independent of the others' success.
#
#   here :  tools/next_gpu_call.sh build
#   GPU  :  gpurun --timeout 1500 -- 'bash tools/next_gpu_call.sh run'
#
# Questions it answers:
#   1. full `-m gpu` suite with the current default kernels (incl. the Hermite / Laguerre NVRTC cases, the example-flow
#      tests, the on-disk image cache on its first real load, the cached-pointer step path of _znll.Plan);
#   2. the same suite with ZNLL_RING=1 (cp.async ring for three- and four-stream trees) -> may it become the default?
#   3. bench lines of all five configs, default and ring (cfg 4 / 5-weighted are the ones the ring touches);
#   4. cfg 2 with and without the bind-time sort (loss option sort_events): what a step costs on unsorted events, i.e.
#      after how many evaluations the 14 ms sort has paid for itself (a fit is ~18 passes) -> lazy-sort policy;
#   5. step overhead above the kernel after the host-side work of the end of round 1 (tools/step_overhead.py).
set -u
cd "$(dirname "$0")/.."
mode="${1:-run}"
RING_LIB="zfit_b200/libznll_ring.so"

if [ "$mode" = build ]; then
  python -m zfit_b200.build || exit 1
  ZNLL_EXTRA_CFLAGS="-DZNLL_RING=1" ZNLL_LIB_OUT="$RING_LIB" ZNLL_OBJ_DIR="zfit_b200/_build_ring" python -m zfit_b200.build || exit 1
  ls -la zfit_b200/libznll.so "$RING_LIB"
  exit 0
fi

mkdir -p gpurun_out
# 1. default suite
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/n_pytest_default.log 2>&1; echo "pytest default rc=$?"
# 3a. default bench lines
for c in 2 3 4 5 1; do
  timeout 300 python bench.py --config $c --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/n_bench_cfg$c.json 2> gpurun_out/n_bench_cfg$c.err
done
# 5. overhead above the kernel
timeout 200 python tools/step_overhead.py 2 > gpurun_out/n_step_overhead_cfg2.txt 2>&1
timeout 200 python tools/step_overhead.py 1 > gpurun_out/n_step_overhead_cfg1.txt 2>&1
# 4. sorted vs unsorted events (cfg 2)
timeout 300 python - > gpurun_out/n_sort_ab_cfg2.txt 2>&1 <<'EOF'
import time, numpy as np, torch
import zfit_b200 as zfit
from zfit_b200 import workloads as W
w = W.build(2, device="cuda:0")
params = list(w["loss"].get_params()); th = [p.value() for p in params]
for sort in (True, False, "lazy"):
    loss = type(w["loss"])(model=w["model"], data=w["data"], options={"sort_events": sort})
    torch.cuda.synchronize(); t0 = time.perf_counter()
    loss.value_gradient(params, full=False)
    torch.cuda.synchronize(); t_first = time.perf_counter() - t0
    K = 200
    t0 = time.perf_counter()
    for k in range(K):
        for p, v in zip(params, th): p.assign(v * (1 + 1e-9 * k))
        loss.value_gradient(params, full=False)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / K
    print(f"sort_events={sort}: first call (bind) {t_first*1e3:.2f} ms, step {dt*1e3:.4f} ms")
EOF
# 2. + 3b. ring variant
if [ -f "$RING_LIB" ]; then
  ZNLL_LIB="$PWD/$RING_LIB" timeout 900 python -m pytest tests -m gpu -q > gpurun_out/n_pytest_ring.log 2>&1; echo "pytest ring rc=$?"
  for c in 4 5 3; do
    ZNLL_LIB="$PWD/$RING_LIB" timeout 300 python bench.py --config $c --steps 300 --warmup 10 --no-cpu-baseline > gpurun_out/n_ring_cfg$c.json 2> gpurun_out/n_ring_cfg$c.err
  done
fi
tail -3 gpurun_out/n_pytest_default.log gpurun_out/n_pytest_ring.log 2>/dev/null
cat gpurun_out/n_sort_ab_cfg2.txt gpurun_out/n_step_overhead_cfg2.txt 2>/dev/null | head -20
