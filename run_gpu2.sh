mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2>gpurun_out/bench_n2.err; echo "n2 rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1]);print(d['value'], d['e2e']['value'], d.get('strong_scaling'), d['kernel_ms'])"
tail -3 gpurun_out/bench_n2.err | cut -c1-300
