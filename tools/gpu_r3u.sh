# round 2 final tree: N = 1 and N = 8 on one 8-GPU box (driver-style --steps 20)
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r3u_bench1.json 2> gpurun_out/r3u_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29761 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r3u_bench8_peer.json 2> gpurun_out/r3u_bench8_peer.err
for f in gpurun_out/r3u_bench*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', 'value', '%.4g'%d['value'], 'us/step', round(d['ms_per_step']*1e3,2), 'coll', d['collectives_in_timed_region'], d['collective_us'] and round(d['collective_us'],1), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3), 'train', d['train'] and (round(d['train']['ms_per_iteration'],2), round(d['train']['ms_update'],2)))"; done
