#!/bin/bash
# Round-2 multi-GPU call (gpurun --gpus N): S4PCS_DEVICES on REAL devices, weak + strong scaling of bench.py at N ranks
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_zzz_devices_gpu.py tests/test_zz_lanes_gpu.py -x -q -m gpu > gpurun_out/r02f_devices_tests_${N}gpu.txt 2>&1; tail -5 gpurun_out/r02f_devices_tests_${N}gpu.txt
SPECS="1 2"; [ "$N" -ge 4 ] && SPECS="1 2 4"; [ "$N" -ge 8 ] && SPECS="1 2 4 8"
timeout 400 python scripts/devices_bench.py --points 1000000 --devices "$SPECS" > gpurun_out/r02f_devices_bench_${N}gpu.jsonl 2>&1; cat gpurun_out/r02f_devices_bench_${N}gpu.jsonl
for sc in strong weak; do
  for n in ${RANKS:-1 $N}; do
    if [ "$n" = "1" ]; then
      timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --scaling $sc --no-cpu-baseline > gpurun_out/r02f_bench_${sc}_1of${N}.json 2> gpurun_out/r02f_bench_${sc}_1of${N}.err
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --scaling $sc > gpurun_out/r02f_bench_${sc}_${n}of${N}.json 2> gpurun_out/r02f_bench_${sc}_${n}of${N}.err
    fi
    python - "$sc" "$n" "gpurun_out/r02f_bench_${sc}_${n}of${N}.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[3]).read().strip().splitlines()[-1]); print(sys.argv[1], 'N='+sys.argv[2], {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling','winner_key')}, d['e2e']['value'])
except Exception as e: print(sys.argv[1], sys.argv[2], 'no bench line', e)
P
  done
done
grep -c "NCCL INFO" gpurun_out/r02f_bench_strong_${N}of${N}.err; grep -m3 "NCCL INFO.*nranks\|NVLS\|comm 0x" gpurun_out/r02f_bench_strong_${N}of${N}.err | cut -c1-200
ls -la gpurun_out | grep r02f_
