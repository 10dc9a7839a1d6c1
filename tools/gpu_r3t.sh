set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r3t_bench2.json 2> gpurun_out/r3t_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29752 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r3t_ref2.json 2> gpurun_out/r3t_ref2.err
tail -3 gpurun_out/r3t_bench2.err
python -c "
import json; d=json.load(open('gpurun_out/r3t_bench2.json')); print('value', '%.4g'%d['value'], 'us/step', round(d['ms_per_step']*1e3,2), 'coll', d['collectives_in_timed_region'], d['collective_us'], 'e2e', round(d['e2e']['ms_per_step'],3), 'train', round(d['train']['ms_per_iteration'],2), d['warmup'], d['steps'])"
head -c 300 gpurun_out/r3t_ref2.json
