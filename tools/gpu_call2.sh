python -m pytest tests -m gpu -q 2>&1 | tail -15
for c in 2 1 4 3 5; do
  timeout 500 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/r2_bench_c${c}_s1.json 2> gpurun_out/r2_bench_c${c}_s1.err; echo "config $c rc=$?"; tail -c 300 gpurun_out/r2_bench_c${c}_s1.err
  python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_c${c}_s1.json"))
    print(b["value"], b["e2e"]["value"], b["e2e"]["pageable"], b["roofline"]["kernel"], b["roofline"]["frac"], b["roofline"]["share_of_step"], b.get("saturated"), (b["cpu_baseline"] or {}).get("value"))
except Exception as ex: print("ERR", ex)
P
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n2_s1.json 2> gpurun_out/r2_bench_n2_s1.err; echo n2 rc=$?; tail -c 400 gpurun_out/r2_bench_n2_s1.err; cut -c1-700 gpurun_out/r2_bench_n2_s1.json
