# A/B of two builds under build_ab/ over a list of (config, items) cases:  bash tools/gpu_ab2.sh "2:131072 2:0 4:100000" lib1.so lib2.so
cases=$1; shift
for rep in 1 2; do
for lib in "$@"; do
  for cs in $cases; do
    c=${cs%%:*}; n=${cs##*:}
    EG_B200_LIB=$PWD/build_ab/$lib python bench.py --config $c --items $n --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('$lib config $c items $n', 'value=%.0f' % d['value'], 'ms_per_step=%.3f' % d['ms_per_step'], 'e2e=%.0f' % d['e2e']['value'], 'kernel_ms=%.2f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'])
"
  done
done
done
