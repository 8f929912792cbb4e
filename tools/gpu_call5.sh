set -x
N=${1:-8}
python -m pytest tests/test_gpu_multi.py -q -x -k "$N" 2>&1 | tail -5
for sc in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --config 2 --gpus $N --steps 3 --warmup 3 --scaling $sc --no-cpu-baseline > gpurun_out/r2_bench_c2_n${N}_${sc}.json 2> gpurun_out/r2_bench_c2_n${N}_${sc}.err; echo "config 2 n=$N $sc rc=$?"; tail -c 300 gpurun_out/r2_bench_c2_n${N}_${sc}.err
  python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_c2_n${N}_${sc}.json"))
    print("$sc config 2 n=$N", b["value"], b["e2e"]["value"], b["e2e"]["pageable"]["value"], b["ms_per_step"], b["config"]["items_per_gpu"])
except Exception as ex: print("ERR", ex)
P
done
python tools/multi_ctx_bench.py --gpus $N > gpurun_out/r2_multi_ctx_n${N}.json 2> gpurun_out/r2_multi_ctx_n${N}.err; tail -c 300 gpurun_out/r2_multi_ctx_n${N}.err; cat gpurun_out/r2_multi_ctx_n${N}.json
python tools/multi_ctx_bench.py --gpus $N --total 1048576 > gpurun_out/r2_multi_ctx_n${N}_strong.json 2>> gpurun_out/r2_multi_ctx_n${N}.err; cat gpurun_out/r2_multi_ctx_n${N}_strong.json
