set -x
N=${1:-2}
python -m pytest tests -m gpu -q -x 2>&1 | tail -8
for sc in weak strong; do
  for c in 2 3; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --config $c --gpus $N --steps 3 --warmup 3 --scaling $sc --no-cpu-baseline > gpurun_out/r2_bench_c${c}_n${N}_${sc}.json 2> gpurun_out/r2_bench_c${c}_n${N}_${sc}.err; echo "config $c n=$N $sc rc=$?"; tail -c 300 gpurun_out/r2_bench_c${c}_n${N}_${sc}.err
    python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_c${c}_n${N}_${sc}.json"))
    print("$sc config $c n=$N", b["value"], b["e2e"]["value"], b["e2e"]["pageable"]["value"], b["ms_per_step"], b["config"]["items_per_gpu"])
except Exception as ex: print("ERR", ex)
P
  done
done
# one process driving all GPUs through eg_ctx_create_multi (no torch, no torchrun): python wrapper over the same C ABI
python tools/multi_ctx_bench.py --gpus $N > gpurun_out/r2_multi_ctx_n${N}.json 2> gpurun_out/r2_multi_ctx_n${N}.err; tail -c 300 gpurun_out/r2_multi_ctx_n${N}.err; cat gpurun_out/r2_multi_ctx_n${N}.json
