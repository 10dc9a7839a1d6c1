timeout 600 python -m pytest tests/test_mlp_train.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/t.log
python tools/mlp_update_run.py > gpurun_out/plain_upd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_mlp_backward|k_mlp_wgrad|k_mlp_forward|k_ppo_loss' -s 8 -c 4 -o gpurun_out/prof_r1_v6_update python tools/mlp_update_run.py > gpurun_out/ncu_upd.log 2>&1
tail -2 gpurun_out/ncu_upd.log
