# k_msm with thread-contiguous window tables in global scratch (persistent grid) against the local-array form
exec > gpurun_out/r2_ab_msm_scratch.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -x -q -m gpu -k "shares or sumsq or qv or commitment or possession or keyset or multi_mul or ciphertext_ops or decryption or fuzz" 2>&1 | tail -3
for rep in 1 2; do
for lib in build_ab/libeg_prev.so elastic_elgamal_b200/libeg_b200.so; do
  for c in 5 3; do
    EG_B200_LIB=$PWD/$lib timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline 2>/tmp/err.txt | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('$lib config $c', 'value=%.0f' % d['value'], 'ms_per_step=%.2f' % d['ms_per_step'], 'e2e=%.0f' % d['e2e']['value'], r['kernel'], 'kernel_ms=%.3f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'])
" || tail -3 /tmp/err.txt
  done
done
done
