# A/B: ge_hot_dbl with only its squarings expanded in place (multiplications as calls) against the fully expanded doubling
exec > gpurun_out/r2_ab_sqinline.txt 2>&1
for rep in 1 2; do
for lib in elastic_elgamal_b200/libeg_b200.so build_ab/libeg_sqinline.so; do
  for c in 2 4 5; do
    EG_B200_LIB=$PWD/$lib timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; k2=r.get('second_kernel') or {}
        print('$lib config $c', 'value=%.0f' % d['value'], 'ms_per_step=%.2f' % d['ms_per_step'], 'kernel_ms=%.3f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'], 'k_commit_ms=%.2f' % (k2.get('ms', 0) / max(1, k2.get('launches', 1))))
" || tail -3 /tmp/err.txt
  done
done
done
