# memcheck over the paths added at the end of round 2: key tables, Ciphertext operators, 24-bit windows in the constant-time walk, pair engine
exec > gpurun_out/r2_memcheck_s8.txt 2>&1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "key_tables or ciphertext_ops or constant_time or mode_3 or group_helpers or multi_mul or seeded" 2>&1 | grep -v "^$" | tail -12
