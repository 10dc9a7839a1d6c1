set -x
mkdir -p gpurun_out
python tools/post_timeline.py 65536 > gpurun_out/r3c_post_timeline.log 2>&1
python - > gpurun_out/r3c_post_timeline_defer.log 2>&1 <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import torch
src = open("tools/post_timeline.py").read().replace("task = LocoManipulationTask(cfg, n, scene=scenes[0], device=dev, fuse_obs_clamp=True)", "task = LocoManipulationTask(cfg, n, scene=scenes[0], device=dev, fuse_obs_clamp=True); task.set_defer_finish(True)")
exec(compile(src, "post_timeline_defer", "exec"))
PY
cat gpurun_out/r3c_post_timeline.log gpurun_out/r3c_post_timeline_defer.log
