"""Per-CTA timeline of the streaming kernel (diagnostics): start/end skew, per-role finish times, SM ids."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import museinference_jl_b200 as m
d = int(os.environ.get("MUSE_D", 65536)); n = int(os.environ.get("MUSE_N", 2048))
be = m.B200Backend("funnel", d, n)
be.set_data(np.random.default_rng(0).standard_normal(d)); be.seed_draws(42)
th0, th1 = np.array([1.0]), np.array([0.45])
be.map_score(th0, th0, 1e-2, include_data=True, warm_start=0)
be.debug_timeline(n + 1)
for name, th, ws in (("cold", th0, 0), ("warm", th1, 1), ("cold", th0, 0)):
    be.profile_reset(True)
    be.map_score(th, th, 1e-2, include_data=True, warm_start=ws)
    p = be.profile()
    t = be.debug_timeline(148, fetch=True)
    t0 = t[:, 0].min()
    st, en = t[:, 0] - t0, t[:, 1] - t0
    print(name, "event ms %.3f | CTA start us: min %.1f max %.1f | end us: min %.1f med %.1f max %.1f | dur us: min %.1f med %.1f max %.1f"
          % (p["solve_ms"], st.min() / 1e3, st.max() / 1e3, en.min() / 1e3, np.median(en) / 1e3, en.max() / 1e3,
             (en - st).min() / 1e3, np.median(en - st) / 1e3, (en - st).max() / 1e3))
    print("   producer done med %.1f | consumers done med %.1f | finisher done med %.1f | distinct SMs %d"
          % (np.median(t[:, 3] - t0) / 1e3, np.median(t[:, 5] - t0) / 1e3, np.median(t[:, 4] - t0) / 1e3, len(set(t[:, 2]))))
    order = np.argsort(en)
    print("   slowest CTAs (cta, smid, end us):", [(int(i), int(t[i, 2]), round(float(en[i]) / 1e3, 1)) for i in order[-5:]])
be.close()
