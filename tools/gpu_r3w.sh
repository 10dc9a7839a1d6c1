set -x
mkdir -p gpurun_out
F="--gpus 8 --steps 20 --warmup 3 --no-train --no-large --no-kernel-times --no-cpu-baseline"
LM_BENCH_WC=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29771 bench.py $F > gpurun_out/r3w_bench8_wc.json 2> gpurun_out/r3w_bench8_wc.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29772 bench.py $F > gpurun_out/r3w_bench8_base.json 2> gpurun_out/r3w_bench8_base.err
for f in gpurun_out/r3w_bench8_*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', 'e2e', round(d['e2e']['ms_per_step'],4), 'value us', round(d['ms_per_step']*1e3,2))"; done
