#!/usr/bin/env bash
# One `ncu --set full` capture of the step kernel of each config (run on the GPU box, AFTER the same command exited 0
# without ncu), plus the launch list of the headline command.  Reports land in gpurun_out/; condense them here with
# tools/ncu_summary.py into profiles/.
#   gpurun --timeout 900 -- 'bash tools/ncu_capture.sh r2 2 3 4 5'
set -u
cd "$(dirname "$0")/.."
tag="$1"; shift
mkdir -p gpurun_out
for c in "$@"; do
  cmd="python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --no-fit --no-bind"
  timeout 200 $cmd > gpurun_out/${tag}_plain_cfg$c.json 2> gpurun_out/${tag}_plain_cfg$c.err || { echo "cfg $c: plain run failed"; continue; }
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:nll_kernel -s 4 -c 1 -f -o gpurun_out/${tag}_prof_cfg$c $cmd > gpurun_out/${tag}_ncu_cfg$c.log 2>&1
  echo "cfg $c ncu rc=$?"
  # condense on the box (the reports are 20 - 40 MB each with the source page; gpurun brings back 64 MB at most)
  ev=100000000; [ "$c" = 5 ] && ev=125000000; [ "$c" = 3 ] && ev=50000000
  python tools/ncu_summary.py gpurun_out/${tag}_prof_cfg$c.ncu-rep $ev gpurun_out/${tag}_cfg${c}_nll_kernel_ncu_full.json \
    "ncu --set full --clock-control none --import-source on -k regex:nll_kernel -s 4 -c 1 $cmd" | tail -1
  [ "$c" = "${KEEP_REP:-2}" ] || rm -f gpurun_out/${tag}_prof_cfg$c.ncu-rep
done
cmd="python bench.py --config 2 --steps 5 --warmup 3 --no-cpu-baseline --no-fit --no-bind"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_cfg2.csv $cmd > gpurun_out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
