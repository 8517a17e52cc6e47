python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale_parity.py -m gpu -x -q 2>&1 | tail -6
python bench.py --no-train-step --no-cpu-baseline --legs "" --steps 10 > gpurun_out/q.json 2> gpurun_out/q.err; tail -2 gpurun_out/q.err
python -c "
import json
d=json.load(open('gpurun_out/q.json')); print(d['value'], d['e2e']['value'], d['per_view_api']['value']); print({k:round(v['ms_avg'],4) for k,v in d['stages'].items()})
"
