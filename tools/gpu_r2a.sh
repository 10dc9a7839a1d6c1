# round 2, GPU call A: full GPU suite, smoke, MLP epilogue variants, default bench
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2a_tests.log
python __graft_entry__.py smoke > gpurun_out/r2a_smoke.log 2>&1
python tools/mlp_variants.py locomanipulationrl_b200/liblm_b200.so variants/liblm_f1.so variants/liblm_f2.so variants/liblm_f3.so > gpurun_out/r2a_mlp_variants.log 2>&1
timeout 900 python bench.py --steps 96 --warmup 6 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_mlp_variants.log; tail -3 gpurun_out/r2a_bench.err
