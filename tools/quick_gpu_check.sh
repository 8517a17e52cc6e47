# usage: bash tools/quick_gpu_check.sh [pytest -k expression]   (env is passed through)
python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -6
python bench.py --no-train-step --no-cpu-baseline --legs "" --steps 10 > gpurun_out/q.json 2> gpurun_out/q.err; tail -2 gpurun_out/q.err
python tools/show_bench.py gpurun_out/q.json
