set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
CMD="python bench.py --config 2 --events 20000000 --steps 3 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_l.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:nll_kernel -s 2 -c 2 -o gpurun_out/prof_cfg2_r1 $CMD > gpurun_out/ncu_f.log 2>&1
tail -5 gpurun_out/ncu_f.log
ls -la gpurun_out
