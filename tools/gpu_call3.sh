set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -15
python tools/prover_bench.py --items 262144 --out gpurun_out/r2_provers.json 2>&1 | tail -8
# racecheck on the tally reduction (shared memory + shuffles), small shapes only: the tool serialises heavily
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_large.py -q -k "tally_reduction_shapes and (129 or 18945)" > gpurun_out/r2_racecheck_tally.txt 2>&1; tail -6 gpurun_out/r2_racecheck_tally.txt
# launch list of one config-2 step at 262k and of config 4 / 3 / 5 at reduced sizes
for c in 2 4 3 5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c$c.csv python bench.py --config $c --items 262144 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_c$c.log 2>&1
  python tools/launch_summary.py gpurun_out/r2_launches_c$c.csv "bench.py --config $c --items 262144 --steps 1 --warmup 1 under ncu (launch list; times are cold-cache and serialised)" > gpurun_out/r2_launches_c${c}_summary.txt; head -14 gpurun_out/r2_launches_c${c}_summary.txt
done
# full captures: the long-ring k_ring shape (config 4) and k_msm (config 5)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ring -c 1 -o gpurun_out/r2_k_ring_long_c4 python bench.py --config 4 --items 47662 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_full_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm -c 1 -o gpurun_out/r2_k_msm_c5 python bench.py --config 5 --items 262144 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_full_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ring -c 1 -o gpurun_out/r2_k_ring_c2 python bench.py --config 2 --items 257638 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_full_c2.log 2>&1
ls -la gpurun_out/*.ncu-rep
