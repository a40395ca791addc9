mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
run() { # name nproc extra-args...
  name=$1; np=$2; shift 2
  timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $np --cpu-sample 0 "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  echo "$name rc=$?"; grep -E "exchange verified|Error|error" gpurun_out/bench_$name.err | tail -3
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$name.json').read());print('$name',round(d['value']),d['ms_per_step'],round(d['e2e']['value']),d['kernel_ms'],d.get('strong_scaling'),d.get('exchange'))"
}
run n8 8 --steps 10 --warmup 3 --verify-exchange
run n4 4 --steps 10 --warmup 3
run c3_n8 8 --config C3 --steps 10 --warmup 3
run c3_n1 1 --config C3 --steps 10 --warmup 3
