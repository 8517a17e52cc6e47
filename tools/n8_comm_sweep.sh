# 8-GPU sweep of the gradient-exchange modes (run under gpurun --gpus 8)
N=${N:-8}
run() {  # name, extra env, args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --no-train-step --no-cpu-baseline --legs "" "$@" > gpurun_out/n${N}_$name.json 2> gpurun_out/n${N}_$name.err
  python -c "
import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d.get('exchange_check'))
" gpurun_out/n${N}_$name.json || tail -5 gpurun_out/n${N}_$name.err
}
run nvls --comm nvls
run nvls_sh4 --comm nvls_sh --comm-chunks 4
run nvls_sh2 --comm nvls_sh --comm-chunks 2
run nvls_sh8 --comm nvls_sh --comm-chunks 8
TGR_NVLS_CTAS=148 run nvls_sh4_c148 --comm nvls_sh --comm-chunks 4
TGR_NVLS_CTAS=16 run nvls_sh4_c16 --comm nvls_sh --comm-chunks 4
