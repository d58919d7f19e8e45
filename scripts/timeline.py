"""Controller timeline of one cold + one warm pass of the GENERIC two-sweep kernel (diagnostics; SM clock cycles).
The streaming kernel has its own per-CTA timeline: scripts/stream_timeline.py."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
d = int(os.environ.get("MUSE_D", 65536)); n = int(os.environ.get("MUSE_N", 2048))
be = m.B200Backend("funnel", d, n, group=int(os.environ.get("MUSE_GROUP", 0)), cluster=int(os.environ.get("MUSE_CLUSTER", 0)),
                   kernel=int(os.environ.get("MUSE_KERNEL", 1)))
print("geometry", be.geometry())
be.set_data(np.random.default_rng(0).standard_normal(d)); be.seed_draws(42)
th0, th1 = np.array([1.0]), np.array([0.45])
be.map_score(th0, th0, 1e-2, include_data=True, warm_start=0)
be.debug_timeline(n + 1)
for name, th, ws in (("cold", th0, 0), ("warm", th1, 1)):
    be.map_score(th, th, 1e-2, include_data=True, warm_start=ws)
    t = be.debug_timeline(n + 1, fetch=True)
    g = be.geometry()["groups"]
    d01 = t[:, 1] - t[:, 0]; d12 = t[:, 2] - t[:, 1]; d23 = t[:, 3] - t[:, 2]; d34 = t[:, 4] - t[:, 3]; d45 = t[:, 5] - t[:, 4]
    tot = t[:, 5] - t[:, 0]
    # gap between consecutive units of the same group
    gaps = []
    for i in range(g, n + 1):
        gaps.append(t[i, 0] - t[i - g, 5])
    md = lambda a, b: np.median(t[:, a] - t[:, b])
    if t[:, 6].any():
        print(name, "consumer: INIT cmd→start %.0f loop %.0f fence %.0f reduce %.0f | TRIAL loop %.0f fence %.0f reduce %.0f | ctl: INIT issue→consumer start %.0f, TRIAL consumer end→HZ return %.0f"
              % (md(6, 0), md(7, 6), md(8, 7), md(9, 8), md(11, 10), md(12, 11), md(13, 12), md(6, 0), md(3, 13)))
    print(name, "cycles median: INIT %.0f | scalar→HZ %.0f | HZ(incl TRIAL) %.0f | post %.0f | outputs %.0f | total %.0f | inter-unit gap %.0f"
          % (np.median(d01), np.median(d12), np.median(d23), np.median(d34), np.median(d45), np.median(tot), np.median(gaps) if gaps else -1))
be.close()
