# A/B: chunk of 34 waves instead of 17; fixed-base prefetch into L1 instead of L2; ncu of k_msm with key tables (config 5)
exec > gpurun_out/r2_ab_chunk_prefetch.txt 2>&1
run() {  # lib label args...
  lib=$1; label=$2; shift 2
  EG_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline "$@" 2>/tmp/err.txt | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']
        print('$label', 'value=%.0f' % d['value'], 'ms_per_step=%.2f' % d['ms_per_step'], 'e2e=%.0f' % d['e2e']['value'], 'pageable=%.0f' % d['e2e']['pageable']['value'], 'kernel_ms=%.3f' % r['avg_launch_ms'], 'share=%.3f' % r['share_of_step'])
" || tail -3 /tmp/err.txt
}
for rep in 1 2; do
  run elastic_elgamal_b200/libeg_b200.so "default chunk (17 waves) config 2" --config 2
  run elastic_elgamal_b200/libeg_b200.so "chunk 515276 (34 waves) config 2" --config 2 --chunk 515276
  run build_ab/libeg_pf2.so "prefetch.L1 config 2" --config 2
  run elastic_elgamal_b200/libeg_b200.so "default config 4" --config 4
  run build_ab/libeg_pf2.so "prefetch.L1 config 4" --config 4
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_msm -c 1 -o gpurun_out/r2_k_msm_c5_keytables python bench.py --config 5 --items 262144 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r2_ncu_full_c5_kt.log 2>&1
ls -la gpurun_out/r2_k_msm_c5_keytables.ncu-rep
