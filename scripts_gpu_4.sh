python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for c in 2 3 4 5 1; do python bench.py --config $c --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b$c.json 2> gpurun_out/b$c.err || tail -5 gpurun_out/b$c.err; python - <<PY
import json
d=json.loads(open('gpurun_out/b$c.json').read()); r=d['roofline']
print('cfg', d['config']['workload'][:5], 'evals/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'kernel_ms', round(r['kernel_ms'],4), 'bound', r['bound'], 'frac', round(r['frac'],3), 'TF', round(r['algorithmic_flops_per_event']*d['config']['events_per_gpu']/r['kernel_ms']/1e9,2), 'GB/s', round(r['hbm_achieved_gbs']), d['clocks'])
PY
done
