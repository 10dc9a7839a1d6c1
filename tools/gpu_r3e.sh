set -x
mkdir -p gpurun_out
{ nproc; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"; ls /sys/devices/system/node/; cat /sys/devices/system/node/node*/cpulist; python -c "import os;print(sorted(os.sched_getaffinity(0)))"; nvidia-smi topo -m | head -6; } > gpurun_out/r3e_host.log 2>&1
F="--steps 48 --warmup 6 --no-cpu-baseline --no-train --no-large --no-kernel-times"
python bench.py $F > gpurun_out/r3e_default.json 2>/dev/null
OMP_NUM_THREADS=1 python bench.py $F > gpurun_out/r3e_omp1.json 2>/dev/null
LM_BENCH_CORES=0-7 python bench.py $F > gpurun_out/r3e_c0_7.json 2>/dev/null
LM_BENCH_CORES=0-15 python bench.py $F > gpurun_out/r3e_c0_15.json 2>/dev/null
LM_BENCH_CORES=16-31 python bench.py $F > gpurun_out/r3e_c16_31.json 2>/dev/null
LM_BENCH_CORES=0-1 python bench.py $F > gpurun_out/r3e_c0_1.json 2>/dev/null
python bench.py $F --e2e-shards 1 > gpurun_out/r3e_s1.json 2>/dev/null
CUDA_VISIBLE_DEVICES=1 python bench.py $F > gpurun_out/r3e_gpu1.json 2>/dev/null
cat gpurun_out/r3e_host.log
for f in gpurun_out/r3e_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', 'e2e', round(d['e2e']['ms_per_step'],3), 'readback', round(d['e2e']['with_obs_readback']['ms_per_step'],3))"; done
