for sh in 1 8; do
python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-large --no-kernel-times --e2e-shards $sh > gpurun_out/e2e_$sh.json 2> gpurun_out/e2e_$sh.err
done
python tools/pcie_bw.py > gpurun_out/pcie.log 2>&1
