# round 2, 2-GPU call: 2-rank pytest tier, bench with the one-shot NVLink collective and with NCCL
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2z_topo.log 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s 2>&1 | tail -70 > gpurun_out/r2z_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2z_bench2_peer.json 2> gpurun_out/r2z_bench2_peer.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 3 --collective nccl --no-e2e > gpurun_out/r2z_bench2_nccl.json 2> gpurun_out/r2z_bench2_nccl.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/r2z_bench1.json 2> gpurun_out/r2z_bench1.err
tail -30 gpurun_out/r2z_tests.log; tail -5 gpurun_out/r2z_bench2_peer.err
for f in gpurun_out/r2z_bench*.json; do python -c "
import json; d=json.load(open('$f')); print('$f', 'value', '%.4g'%d['value'], 'us/step', round(d['ms_per_step']*1e3,2), 'coll', d['collectives_in_timed_region'], d['collective_us'], 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3), 'train', d['train'] and (round(d['train']['ms_per_iteration'],2), round(d['train']['ms_update'],2)))"; done
