python tools/kernel_zoo.py > gpurun_out/zoo.log 2>&1
