nvidia-smi -L | head -4
timeout 200 python -m pytest tests/test_comm_gpu.py -q -m gpu > gpurun_out/r02s_comm_tests_4gpu.txt 2>&1; tail -4 gpurun_out/r02s_comm_tests_4gpu.txt
timeout 150 python -m pytest tests/test_zzz_devices_gpu.py -q -m gpu -k "library_side" >> gpurun_out/r02s_comm_tests_4gpu.txt 2>&1; tail -3 gpurun_out/r02s_comm_tests_4gpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02s_bench_weak_4of4.json 2> gpurun_out/r02s_bench_weak_4of4.err; cut -c1-260 gpurun_out/r02s_bench_weak_4of4.json; python -c "
import json; d=json.loads(open('gpurun_out/r02s_bench_weak_4of4.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['collective'], d['cpu_baseline']['value'] if d.get('cpu_baseline') else None)"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --scaling strong --no-cpu-baseline > gpurun_out/r02s_bench_strong_4of4.json 2> gpurun_out/r02s_bench_strong_4of4.err; cut -c1-200 gpurun_out/r02s_bench_strong_4of4.json
timeout 100 python scripts/devices_bench.py --points 1000000 --devices "4 4+nccl" > gpurun_out/r02s_devices_bench_4gpu.jsonl 2>/dev/null; cat gpurun_out/r02s_devices_bench_4gpu.jsonl
