# A/B of the builds under build_ab/ on configs 2 and 4 (run on the GPU box):  bash tools/gpu_ab.sh lib1.so lib2.so ...
for rep in 1 2; do
for lib in "$@"; do
  for c in 2 4; do
    EG_B200_LIB=$PWD/build_ab/$lib python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('$lib config $c', 'value=%.0f' % d['value'], 'e2e=%.0f' % d['e2e']['value'], 'pageable=%.0f' % d['e2e']['pageable']['value'], 'kernel_ms=%.2f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'])
"
  done
done
done
