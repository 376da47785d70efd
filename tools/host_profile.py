#!/usr/bin/env python
"""cProfile of the HOST side of the headline step (enqueue only, no synchronisation inside the
loop): where do the ~8 ms of Python / ctypes / driver time per step go?"""
import cProfile
import io
import os
import pstats
import runpy
import sys

sys.argv = ["step_timeline.py", "1"]
here = os.path.dirname(os.path.abspath(__file__))
ns = runpy.run_path(os.path.join(here, "step_timeline.py"), run_name="timeline")
import torch  # noqa: E402

step = ns["step"]
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    step(i)
pr.disable()
torch.cuda.synchronize()
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45)
    print(s.getvalue())
