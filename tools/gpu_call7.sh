set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
for c in 2 1 4 3 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_config${c}_s2.json 2> gpurun_out/r2_bench_config${c}_s2.err; echo "config $c rc=$?"; tail -c 200 gpurun_out/r2_bench_config${c}_s2.err
  python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_config${c}_s2.json"))
    print(b["value"], b["e2e"]["value"], b["e2e"]["pageable"]["value"], b["roofline"]["kernel"], b["roofline"]["frac"], b["roofline"]["share_of_step"], b.get("saturated",{}).get("value"), (b["cpu_baseline"] or {}).get("value"))
except Exception as ex: print("ERR", ex)
P
done
python tools/prover_bench.py --items 262144 --out gpurun_out/r2_provers_s2.json 2>&1 | tail -5
