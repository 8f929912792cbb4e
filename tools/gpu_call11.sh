# full validation of the current tree: GPU suite, smoke, the five bench lines, the reference arm
exec > gpurun_out/r2_validate_s5.txt 2>&1
date
( time python -m pytest tests -m gpu -q 2>&1 | tail -6 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in 2 1 4 3 5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_config${c}_s5.json 2> gpurun_out/r2_bench_config${c}_s5.err; echo "config $c rc=$?"; tail -c 200 gpurun_out/r2_bench_config${c}_s5.err
  python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_config${c}_s5.json"))
    print(b["value"], b["ms_per_step"], b["e2e"]["value"], b["e2e"]["pageable"]["value"], b["roofline"]["kernel"], b["roofline"]["frac"], b["roofline"]["share_of_step"], b.get("saturated",{}).get("value"), (b["cpu_baseline"] or {}).get("value"), b["gpu_launches"], b["clocks"])
except Exception as ex: print("ERR", ex)
P
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm_s5.json 2>gpurun_out/r2_bench_reference_arm_s5.err; cut -c1-400 gpurun_out/r2_bench_reference_arm_s5.json
date
