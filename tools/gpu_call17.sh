# N = 2: is the slow rank of r2_n2_s7 reproducible, and does it depend on the fixed-base window width?
exec > gpurun_out/r2_n2_s8.txt 2>&1
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,clocks.max.mem,temperature.gpu,power.draw,ecc.errors.corrected.volatile.total --format=csv
for lib in elastic_elgamal_b200/libeg_b200.so build_ab/libeg_wide16.so elastic_elgamal_b200/libeg_b200.so; do
  EG_B200_LIB=$PWD/$lib timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --config 2 --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        b=json.loads(line)
        print('$lib', b['value'], b['e2e']['value'], b['ms_per_step'], b.get('per_rank'))
"
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.mem,clocks.max.mem,temperature.gpu,power.draw --format=csv
