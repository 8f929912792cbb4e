# key tables for share verification (config 5) + the instruction-fetch question: k_ring of the same source with 16-bit windows
exec > gpurun_out/r2_keytables.txt 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_multi.py -x -q -m gpu -k "shares or decryption or dlog" 2>&1 | tail -3
for rep in 1 2; do
  python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('config 5', 'value=%.0f' % d['value'], 'ms_per_step=%.2f' % d['ms_per_step'], 'e2e=%.0f' % d['e2e']['value'], 'pageable=%.0f' % d['e2e']['pageable']['value'], r['kernel'], 'kernel_ms=%.3f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'], 'frac=%.3f' % r['frac'])
"
done
EG_B200_LIB=$PWD/build_ab/libeg_wide16.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ring -c 1 -o gpurun_out/r2_k_ring_c2_w16_s7 python bench.py --config 2 --items 257638 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_full_c2_w16.log 2>&1
ls -la gpurun_out/r2_k_ring_c2_w16_s7.ncu-rep
