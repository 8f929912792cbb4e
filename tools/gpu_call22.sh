# ncu launch lists with the final library: config 5 (key tables) and config 1 (pair engine)
for c in 5 1; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_c${c}_final.csv python bench.py --config $c $( [ $c = 5 ] && echo "--items 262144" ) --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_ncu_c${c}_final.log 2>&1
  python tools/launch_summary.py gpurun_out/r2_launches_c${c}_final.csv "bench.py --config $c --steps 1 --warmup 1 under ncu (launch list; times are cold-cache and serialised), final library of round 2" > gpurun_out/r2_launches_c${c}_final_summary.txt
  head -16 gpurun_out/r2_launches_c${c}_final_summary.txt
done
