set -x
N=${1:-2}
for sc in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --config 2 --gpus $N --steps 5 --warmup 3 --scaling $sc --no-cpu-baseline > gpurun_out/r2_bench_c2_n${N}_${sc}_s2.json 2> gpurun_out/r2_bench_c2_n${N}_${sc}_s2.err; echo "config 2 n=$N $sc rc=$?"; tail -c 300 gpurun_out/r2_bench_c2_n${N}_${sc}_s2.err
  python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_c2_n${N}_${sc}_s2.json"))
    print("$sc config 2 n=$N", b["value"], b["e2e"]["value"], b["e2e"]["pageable"]["value"], b["ms_per_step"], b["config"]["items_per_gpu"], b.get("per_rank"))
except Exception as ex: print("ERR", ex)
P
done
