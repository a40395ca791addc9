mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_n2.err > gpurun_out/bench_n2.json
echo "stdout lines: $(wc -l < gpurun_out/bench_n2.json)"; python -c "
import json;d=json.loads(open('gpurun_out/bench_n2.json').read());print(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['kernel_ms'])"
grep -c "NCCL version" gpurun_out/bench_n2.err
