# pair-engine validation + A/B of its CTA size on BASELINE config 1 (10 000 Boolean ciphertexts) and small range batches
exec > gpurun_out/r2_pair_ab.txt 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mode_3 or mode_1 or mode_auto or verify_bool or verify_range" 2>&1 | tail -3
for lib in elastic_elgamal_b200/libeg_b200.so build_ab/libeg_pair32.so build_ab/libeg_pair128.so; do
  for rep in 1 2; do
    EG_B200_LIB=$PWD/$lib python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('$lib', 'value=%.0f' % d['value'], 'ms_per_step=%.3f' % d['ms_per_step'], 'e2e=%.0f' % d['e2e']['value'], r['kernel'], 'kernel_ms=%.3f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'])
"
  done
done
python bench.py --config 1 --steps 10 --warmup 3 --no-cpu-baseline --ring-mode 1 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('mode1', 'value=%.0f' % d['value'], 'ms_per_step=%.3f' % d['ms_per_step'], r['kernel'], 'kernel_ms=%.3f' % r['avg_launch_ms'])
"
