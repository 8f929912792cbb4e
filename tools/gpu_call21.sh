# final validation of round 2: GPU suite, smoke, the driver's bench command, config 5 (key tables), the reference arm
exec > gpurun_out/r2_validate_final.txt 2>&1
date
( time python -m pytest tests -m gpu -q 2>&1 | tail -4 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r2_bench_final.err
python - <<P
import json
b=json.load(open("gpurun_out/r2_bench_final.json"))
print(b["value"], b["ms_per_step"], b["e2e"], b["clocks"], b["roofline"]["frac"], b["roofline"]["executed"], b["roofline"]["traffic"], b["gpu_launches"], b["cpu_baseline"])
P
timeout 600 python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/r2_bench_config5_final.json 2> gpurun_out/r2_bench_config5_final.err; echo "config 5 rc=$?"
python - <<P
import json
b=json.load(open("gpurun_out/r2_bench_config5_final.json"))
print(b["value"], b["ms_per_step"], b["e2e"], b["roofline"]["kernel"], b["roofline"]["frac"], b["roofline"]["share_of_step"], b["cpu_baseline"])
P
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm_final.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_reference_arm_final.json
date
