set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e_topo.log 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -s 2>&1 | tail -70 > gpurun_out/r2e_tests.log
F="--gpus 2 --steps 96 --warmup 6 --no-cpu-baseline"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py $F > gpurun_out/r2e_bench2_peer.json 2> gpurun_out/r2e_bench2_peer.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py $F --collective nccl --no-e2e > gpurun_out/r2e_bench2_nccl.json 2> gpurun_out/r2e_bench2_nccl.err
tail -40 gpurun_out/r2e_tests.log; tail -5 gpurun_out/r2e_bench2_peer.err
