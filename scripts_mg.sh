timeout 300 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -5
for n in 1 2; do
if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err;
else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err; fi
grep -iE "error|timeout" gpurun_out/scale_$n.err | head -3 | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/scale_$n.json').read().strip().splitlines()[-1])
print($n, 'evals/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],4), {k:v for k,v in d.get('fit',{}).items() if k!='fitted'})
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --config 5 --steps 50 --warmup 5 > gpurun_out/scale5_2.json 2> gpurun_out/scale5_2.err; grep -iE "error|timeout" gpurun_out/scale5_2.err | head -3 | cut -c1-300; cat gpurun_out/scale5_2.json | cut -c1-400
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
