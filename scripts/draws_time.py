"""Time seed_draws at the C3 shape (diagnostics)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
be = m.B200Backend("funnel", 65536, 2048)
be.seed_draws(1)
be.profile_reset(True)
for s in range(10):
    be.seed_draws(2 + s)
p = be.profile()
print("seed_draws: %.3f ms per call (events), %d launches" % (p["draw_ms"] / 10, p["draw_launches"]))
be.close()
