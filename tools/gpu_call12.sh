# A/B of the fixed-base window width (EG_WIDE_BITS 16 / 20 / 24) on config 2 (1 M ballots) and config 4 (range proofs)
exec > gpurun_out/r2_ab_wide_bits.txt 2>&1
for rep in 1 2; do
for lib in elastic_elgamal_b200/libeg_b200.so build_ab/libeg_wide20.so build_ab/libeg_wide24.so; do
  for c in 2 4; do
    EG_B200_LIB=$PWD/$lib timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('$lib config $c', 'value=%.0f' % d['value'], 'ms_per_step=%.2f' % d['ms_per_step'], 'e2e=%.0f' % d['e2e']['value'], 'kernel_ms=%.3f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'])
" || tail -3 /tmp/err.txt
  done
done
done
