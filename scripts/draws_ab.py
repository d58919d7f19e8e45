"""A/B of the normal generators of seed_draws (MUSE_DRAWS_IMPL=0 libm transform, 1 table-driven transform, 2 the same arithmetic with
the instruction count trimmed — two streams per thread, 3 — one stream per thread): device output against the oracle generator element
by element, a digest of the output (1, 2 and 3 must agree bit for bit), and the event time per C3 seed.  No torch import."""
import hashlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    import museinference_jl_b200 as m
    import oracle as O
    d, n, seed, off = 1000, 7, 0xDEADBEEF12345, 40
    be = m.B200Backend("funnel", d, n, sim_offset=off)
    be.seed_draws(seed)
    xi, nu = be.get_draws(0, n + 1)
    err = 0.0
    for k in range(n):
        err = max(err, np.abs(xi[k] - O.philox_normals(seed, off + k, 0, d)).max(), np.abs(nu[k] - O.philox_normals(seed, off + k, 1, d)).max())
    err = max(err, np.abs(xi[n] - O.philox_normals(seed, O.philox.MASTER_INDEX, 0, d)).max())
    dig = hashlib.sha256(xi.tobytes() + nu.tobytes())
    be.close()
    be = m.B200Backend("funnel", 4097, 3)              # odd d: the last pair is half a pair
    be.seed_draws(7)
    xo, no = be.get_draws(0, 4)
    dig.update(xo.tobytes() + no.tobytes())
    be.close()
    be = m.B200Backend("funnel", 65536, 2048)
    be.seed_draws(1)
    be.profile_reset(True)
    for s in range(10):
        be.seed_draws(2 + s)
    p = be.profile()
    xi, nu = be.get_draws(100, 64)
    dig.update(xi.tobytes() + nu.tobytes())
    xi2, nu2 = be.get_draws(2040, 9)                   # the last rows and the master draw
    dig.update(xi2.tobytes() + nu2.tobytes())
    print("impl %s: max |device - oracle| = %.3e ; seed_draws %.4f ms per C3 seed (2049 x 65536 x 2 normals) ; sample mean %.5f std %.5f ; sha256 %s"
          % (os.environ.get("MUSE_DRAWS_IMPL"), err, p["draw_ms"] / 10, np.concatenate([xi, nu]).mean(), np.concatenate([xi, nu]).std(),
             dig.hexdigest()[:16]))
    be.close()
else:
    for impl in (sys.argv[1:] or ["0", "1", "2", "3"]):
        env = dict(os.environ, MUSE_DRAWS_IMPL=impl)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True, timeout=300)
        print(r.stdout.strip() or ("impl %s FAILED: " % impl + r.stderr[-2000:]))
