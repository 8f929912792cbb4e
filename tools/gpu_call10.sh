# compute-sanitizer over the pair engine (k_ring_pair uses two-lane shuffles): memcheck, synccheck, racecheck; new entry point test
exec > gpurun_out/r2_sanitizer_pair.txt 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ciphertext_ops or multi_mul or mode_3" 2>&1 | tail -3
for tool in memcheck synccheck racecheck; do
  echo "=== compute-sanitizer --tool $tool: pytest -m gpu -k mode_3"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mode_3" 2>&1 | grep -v "^$" | tail -8
done
