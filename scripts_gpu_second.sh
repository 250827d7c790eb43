set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -30
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; tail -3 gpurun_out/bench_cfg2.err; cat gpurun_out/bench_cfg2.json
for c in 1 3 4 5; do python bench.py --config $c --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err; tail -2 gpurun_out/bench_cfg$c.err; cat gpurun_out/bench_cfg$c.json; done
