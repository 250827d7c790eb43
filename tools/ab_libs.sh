#!/usr/bin/env bash
# A/B of kernel variants on the GPU box: every library given as NAME=PATH runs the same bench legs, interleaved and
# repeated, so that box-to-box and run-to-run noise (a few per cent on these kernels) can be told from an effect.
#   gpurun -- 'bash tools/ab_libs.sh "2 3 4 5" 3 base=zfit_b200/libznll.so bps3=zfit_b200/libznll_bps3.so'
set -u
cd "$(dirname "$0")/.."
cfgs="$1"; reps="$2"; shift 2
mkdir -p gpurun_out
for r in $(seq 1 "$reps"); do
  for spec in "$@"; do
    name="${spec%%=*}"; lib="${spec#*=}"
    for c in $cfgs; do
      ZNLL_LIB="$PWD/$lib" timeout 200 python bench.py --config $c --steps 200 --warmup 10 --no-cpu-baseline --no-fit --no-bind \
        > gpurun_out/ab_${name}_cfg${c}_r$r.json 2> gpurun_out/ab_${name}_cfg${c}_r$r.err
      python - "$name" "$c" "$r" gpurun_out/ab_${name}_cfg${c}_r$r.json <<'PY'
import json, sys
name, c, r, f = sys.argv[1:]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f"{name:>8} cfg{c} rep{r}: kernel {d['roofline']['kernel_ms']:.4f} ms  step {d['ms_per_step']:.4f} ms  clocks {d['clocks']['sm_mhz']} {d['clocks']['reasons']}")
except Exception as e:
    print(f"{name:>8} cfg{c} rep{r}: FAILED {e}")
PY
    done
  done
done
