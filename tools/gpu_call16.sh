# N = 2: multi-GPU tests and bench lines with the final library (24-bit windows, key tables)
exec > gpurun_out/r2_n2_s7.txt 2>&1
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3
for c in 2 5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --config $c --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c${c}_n2_weak_s7.json 2> gpurun_out/r2_bench_c${c}_n2_weak_s7.err; echo "config $c n=2 rc=$?"; tail -c 300 gpurun_out/r2_bench_c${c}_n2_weak_s7.err
  python - <<P
import json
try:
    b=json.load(open("gpurun_out/r2_bench_c${c}_n2_weak_s7.json"))
    print("weak config $c n=2", b["value"], b["e2e"]["value"], b["e2e"]["pageable"]["value"], b["ms_per_step"], b["config"]["items_per_gpu"], b.get("per_rank"))
except Exception as ex: print("ERR", ex)
P
done
timeout 600 python tools/multi_ctx_bench.py --gpus 2 2>&1 | tail -4
