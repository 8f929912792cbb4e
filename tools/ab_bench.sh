#!/bin/bash
# A/B bench of alternative builds of the CUDA library: tools/ab_bench.sh lib1.so lib2.so ...   (run on the GPU box)
for lib in "$@"; do
  EG_B200_LIB=$PWD/$lib python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${AB_ARGS} 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; k2=r.get('second_kernel') or {}
        print('$lib', 'value=%.0f' % d['value'], 'e2e=%.0f' % d['e2e']['value'], 'kernel_ms=%.2f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'], 'frac=%.3f' % r['frac'],
              'k_commit_ms=%.2f' % (k2.get('ms', 0) / max(1, k2.get('launches', 1))))
    elif line: print('$lib', line[:200])
"
done
