set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_cabi_consumer.py -m gpu -q -k "seeded or sumsq or decryption or wire or from_ciphertext or device_pointer or constant_time or create_multi or consumer or single_choice" > gpurun_out/r2_memcheck.txt 2>&1; tail -5 gpurun_out/r2_memcheck.txt
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>gpurun_out/r2_bench_reference_arm.err; cut -c1-300 gpurun_out/r2_bench_reference_arm.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1m_s2.json 2> gpurun_out/r2_bench_1m_s2.err; echo rc=$?; tail -c 300 gpurun_out/r2_bench_1m_s2.err
python - <<P
import json
b=json.load(open("gpurun_out/r2_bench_1m_s2.json"))
print(b["value"], b["e2e"], b["ms_per_step"], b["clocks"], b["roofline"]["frac"], b["roofline"]["traffic_source"][:60], b["gpu_launches"], b["cpu_baseline"])
P
