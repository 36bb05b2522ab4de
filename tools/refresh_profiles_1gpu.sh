# Developer tool: re-measure the single-GPU lines kept under profiles/ (outputs under gpurun_out/).
tag=${1:-v8}
python bench.py --steps 20 --warmup 3 --calls 4 > gpurun_out/${tag}_tick20.json 2> gpurun_out/${tag}_tick20.err
python bench.py --steps 10 --warmup 3 --envs 16384 --iters 30 --tol 1e-8 > gpurun_out/${tag}_converge.json 2> gpurun_out/${tag}_converge.err
python bench.py --steps 20 --warmup 3 --envs 8192 > gpurun_out/${tag}_8192.json 2> gpurun_out/${tag}_8192.err
python bench.py --steps 50 --warmup 3 --envs 37888 > gpurun_out/${tag}_37888.json 2> gpurun_out/${tag}_37888.err
python bench.py --closed-loop --graphs 1 > gpurun_out/${tag}_closed_loop_dec10_graph.json 2> gpurun_out/${tag}_closed_loop.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_r1_${tag}.csv python bench.py --steps 20 --warmup 3 > gpurun_out/launches_${tag}.log 2>&1
for f in tick20 converge 8192 37888 closed_loop_dec10_graph; do tail -c 400 gpurun_out/${tag}_$f.json | head -c 0; python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_$f.json').read().strip().splitlines()[-1]); print('$f', d.get('value'), d.get('ms_per_step'))
except Exception as e: print('$f', 'ERR', e)
PY
done
