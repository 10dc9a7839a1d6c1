# round 2, 8-GPU call: 2-rank pytest tier, weak-scaling bench at N = 8 and 4 with the one-shot NVLink collective (and NCCL at 8)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2q_topo.log 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r2q_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2q_bench8_peer.json 2> gpurun_out/r2q_bench8_peer.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 8 --steps 20 --warmup 3 --collective nccl --no-e2e > gpurun_out/r2q_bench8_nccl.json 2> gpurun_out/r2q_bench8_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r2q_bench4_peer.json 2> gpurun_out/r2q_bench4_peer.err
tail -12 gpurun_out/r2q_tests.log; tail -5 gpurun_out/r2q_bench8_peer.err
for f in gpurun_out/r2q_bench*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', 'value', '%.4g'%d['value'], 'us/step', round(d['ms_per_step']*1e3,2), 'coll', d['collectives_in_timed_region'], d['collective_us'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3), 'train', d['train'] and (round(d['train']['ms_per_iteration'],2), round(d['train']['ms_update'],2)))"; done
