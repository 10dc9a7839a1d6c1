set -x
mkdir -p gpurun_out
python tools/e2e_probe.py 1 2 4 8 > gpurun_out/r2h_e2e_probe.log 2>&1
python tools/mlp_timeline.py > gpurun_out/r2h_mlp_timeline.log 2>&1
python tools/pcie_bw.py > gpurun_out/r2h_pcie.log 2>&1
cat gpurun_out/r2h_e2e_probe.log gpurun_out/r2h_mlp_timeline.log gpurun_out/r2h_pcie.log
