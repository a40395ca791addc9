mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --cpu-sample 0 --verify-exchange > gpurun_out/bench_n2.json 2>gpurun_out/bench_n2.err; echo "rc=$?"; tail -5 gpurun_out/bench_n2.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_n2.json').read());print('N2',d['value'],d['ms_per_step'],d['e2e']['value'],d['kernel_ms'],d.get('strong_scaling'),d.get('exchange'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --cpu-sample 0 --chunk 512 > gpurun_out/bench_n2_c512.json 2>gpurun_out/bench_n2_c512.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_n2_c512.json').read());print('N2 chunk512',d['value'],d['ms_per_step'],d.get('strong_scaling'))"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 --impl reference > gpurun_out/bench_n2_ref.json 2>gpurun_out/bench_n2_ref.err; tail -c 600 gpurun_out/bench_n2_ref.json
