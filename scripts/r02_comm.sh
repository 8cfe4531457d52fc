#!/bin/bash
# Round-2 call for row e inside the library (csrc/comm.cu): gpurun --gpus N -- bash scripts/r02_comm.sh
#   NCCL-backed reduction tests on real devices, the C++ layer with S4PCS_NCCL=1, bench.py under torchrun with the
#   collective inside libs4g (default) and with round 1's torch glue (--collective torch) for A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_comm_gpu.py tests/test_zzz_devices_gpu.py -x -q -m gpu > gpurun_out/r02q_comm_tests_${N}gpu.txt 2>&1; tail -6 gpurun_out/r02q_comm_tests_${N}gpu.txt
SPECS="1 2 2+nccl"; [ "$N" -ge 4 ] && SPECS="1 2 2+nccl 4 4+nccl"; [ "$N" -ge 8 ] && SPECS="1 2 2+nccl 4 4+nccl 8 8+nccl"
timeout 500 python scripts/devices_bench.py --points 1000000 --devices "$SPECS" > gpurun_out/r02q_devices_bench_${N}gpu.jsonl 2> gpurun_out/r02q_devices_bench_${N}gpu.err; cat gpurun_out/r02q_devices_bench_${N}gpu.jsonl
run() {  # scaling ranks collective
  local out=gpurun_out/r02q_bench_$1_$2of${N}_$3
  if [ "$2" = "1" ]; then
    timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --scaling $1 --collective $3 --no-cpu-baseline > $out.json 2> $out.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $2 --steps 10 --warmup 3 --scaling $1 --collective $3 > $out.json 2> $out.err
  fi
  python - "$1" "$2" "$3" "$out.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[4]).read().strip().splitlines()[-1]); print(sys.argv[1], 'N='+sys.argv[2], sys.argv[3], {k:d.get(k) for k in ('value','ms_per_step','gpu_launches','winner_key')}, round(d['e2e']['value']), d.get('collective',{}).get('comm_ranks'), d.get('collective',{}).get('enqueued_by_rank0'))
except Exception as e: print(sys.argv[1], sys.argv[2], sys.argv[3], 'no bench line', e)
P
}
for n in ${RANKS:-1 $N}; do
  run weak $n native; run weak $n torch; run strong $n native; run strong $n torch
done
grep -c "NCCL INFO" gpurun_out/r02q_bench_strong_${N}of${N}_native.err; grep -m6 "NCCL INFO.*nranks\|Init COMPLETE\|NVLS" gpurun_out/r02q_bench_strong_${N}of${N}_native.err | cut -c1-220
tail -3 gpurun_out/r02q_bench_weak_${N}of${N}_native.err | cut -c1-300
