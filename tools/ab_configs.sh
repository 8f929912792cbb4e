#!/bin/bash
# A/B of alternative builds on other configs: tools/ab_configs.sh "3,4" lib1.so lib2.so ...   (run on the GPU box)
cfg=$1; shift
for lib in "$@"; do
  EG_B200_LIB=$PWD/$lib python tools/bench_configs.py --configs $cfg --steps 2 --warmup 1 2>&1 | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print('$lib', 'config', d['config'], 'gpu_e2e=%.0f' % d['gpu_e2e'], ' '.join('%s=%.0f' % (k, v) for k, v in d.items() if k.startswith('gpu_encrypt')))
    elif line: print('$lib', line[:200])
"
done
