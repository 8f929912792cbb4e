# validation of the 24-bit fixed-base windows: GPU suite, smoke, five bench lines, provers, ncu launch list + full capture of k_ring
exec > gpurun_out/r2_validate_s6.txt 2>&1
date
( time python -m pytest tests -m gpu -q -x 2>&1 | tail -4 ) 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
for c in 2 1 4 3 5; do
  timeout 900 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/r2_bench_config${c}_s6.json 2> gpurun_out/r2_bench_config${c}_s6.err; echo "config $c rc=$?"; tail -c 200 gpurun_out/r2_bench_config${c}_s6.err
  python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_config${c}_s6.json"))
    print(b["value"], b["ms_per_step"], b["e2e"]["value"], b["e2e"]["pageable"]["value"], b["roofline"]["kernel"], b["roofline"]["frac"], b["roofline"]["share_of_step"], b.get("saturated",{}).get("value"), (b["cpu_baseline"] or {}).get("value"), b["gpu_launches"], b["clocks"])
except Exception as ex: print("ERR", ex)
P
done
python tools/prover_bench.py --items 262144 --out gpurun_out/r2_provers_s6.json 2>&1 | tail -6
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c2_s6.csv python bench.py --config 2 --items 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_c2_s6.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_c2_s6.csv "bench.py --config 2 --items 262144 --steps 1 --warmup 1 under ncu (launch list; times are cold-cache and serialised), 24-bit fixed-base windows" > gpurun_out/r2_launches_c2_s6_summary.txt; head -14 gpurun_out/r2_launches_c2_s6_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ring -c 1 -o gpurun_out/r2_k_ring_c2_s6 python bench.py --config 2 --items 257638 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_full_c2_s6.log 2>&1
ls -la gpurun_out/r2_k_ring_c2_s6.ncu-rep
date
