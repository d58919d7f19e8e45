"""The table-driven Box–Muller transform of the draws kernel (museinference.jl_b200/csrc/muse_normal_math.cuh), checked
on the CPU: the header is written with explicit fma() so that its host build (tests/csrc/normal_math_host.cpp, g++)
computes bit for bit what the device computes.  References: long-double libm, and the oracle's NumPy generator
(oracle/philox.py), which the GPU test test_device_philox_matches_oracle then compares with the device output."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from oracle.philox import philox4x32_10

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("nm") / "libnm.so")
    src = os.path.join(ROOT, "tests", "csrc", "normal_math_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src], check=True)
    return C.CDLL(out)


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _sample_v(rng):
    return np.concatenate([
        rng.integers(0, 2 ** 53, size=400_000, dtype=np.uint64),
        (2 ** 53 - 2 - rng.integers(0, 2 ** 34, size=50_000)).astype(np.uint64),      # u → 1: r = sqrt(−2 ln u) → 0
        (2 ** 53 - 2 - rng.integers(0, 2 ** 47, size=50_000)).astype(np.uint64),
        (2 ** 53 - 2 - rng.integers(0, 2 ** 51, size=50_000)).astype(np.uint64),      # around the table entries next to c = 1
        rng.integers(0, 2 ** 24, size=50_000, dtype=np.uint64),                       # u → 0: the far tail
        np.array([0, 1, 2, 3, 2 ** 53 - 1, 2 ** 53 - 2, 2 ** 52, 2 ** 52 - 1, 2 ** 52 + 1, 2 ** 45 - 1, 2 ** 45, 2 ** 44,
                  2 ** 53 - 2 ** 45, 2 ** 53 - 2 ** 44], dtype=np.uint64)])


def test_neg2log_against_long_double(nm):
    v = _sample_v(np.random.default_rng(11))
    out = np.empty(v.size)
    nm.muse_host_neg2log(_vp(v), C.c_int(v.size), _vp(out))
    y = (v.astype(np.float64) + 0.5).astype(np.longdouble)       # the generator's definition of u₁·2⁵³ (one FP64 rounding)
    mm, k = np.frexp(y)
    l2 = np.log(np.longdouble(2))
    ref = np.where(k == 53, -2 * np.log1p(mm - 1), -2 * (np.log(mm) + (k - 53) * l2))   # cancellation-free as u → 1
    ok = ref != 0
    rel = np.abs((out.astype(np.longdouble) - ref)[ok] / ref[ok]).astype(np.float64)
    assert rel.max() < 1e-14                                     # worst next to the table entry c = 1 (≈ 5e-15)
    assert (out[~ok] == 0).all() and (out >= 0).all()
    rad, rad_ref = np.sqrt(out), np.sqrt(ref).astype(np.float64)
    assert np.abs(rad - rad_ref).max() < 2e-15                   # ≤ 1 ulp at r ≈ 8, relatively accurate as r → 0
    assert (np.abs(rad - rad_ref)[ok] / rad_ref[ok]).max() < 5e-15


def test_sincos_against_long_double(nm):
    v = _sample_v(np.random.default_rng(12))
    cs, sn = np.empty(v.size), np.empty(v.size)
    nm.muse_host_sincos(_vp(v), C.c_int(v.size), _vp(cs), _vp(sn))
    pil = np.longdouble("3.14159265358979323846264338327950288")
    ang = 2 * pil * ((v.astype(np.longdouble) + np.longdouble(0.5)) / np.longdouble(2) ** 53)
    assert np.abs(cs - np.cos(ang).astype(np.float64)).max() < 3e-16
    assert np.abs(sn - np.sin(ang).astype(np.float64)).max() < 3e-16


@pytest.mark.parametrize("sim,stream", [(0, 0), (41, 1), (O.philox.MASTER_INDEX, 0)])
def test_normals_match_the_oracle_generator(nm, sim, stream):
    seed, d = 0xDEADBEEF12345, 100_000
    npairs = d // 2
    r = np.ascontiguousarray(np.stack(philox4x32_10(np.arange(npairs, dtype=np.uint64), sim, stream, 0, seed & 0xFFFFFFFF,
                                                    (seed >> 32) & 0xFFFFFFFF), axis=1).astype(np.uint32))
    out = np.empty(2 * npairs)
    nm.muse_host_box_muller(_vp(r), C.c_int(npairs), _vp(out))
    ref = O.philox_normals(seed, sim, stream, d)                 # NumPy libm log/cos/sin of the same uniforms
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-14)     # the GPU test allows 2e-13 between device and oracle
    assert abs(out.mean()) < 0.02 and abs(out.std() - 1) < 0.01


def test_trimmed_kernel_transform_is_bit_identical_to_the_header(tmp_path):
    """`box_muller_regs` of csrc/muse_draws.cu (the transform of philox_draws_tab2_kernel, the default draws kernel: funnel-shift
    assembly of the 53-bit integers, exponent-OR conversion of the angle remainder, constants in registers) pasted verbatim into a
    host build and compared BIT FOR BIT with the header's `box_muller_tab` (the round-1 kernel's transform, pinned above against
    long-double references) on 4 million random Philox blocks and the edge cases of both uniforms."""
    src_cu = open(os.path.join(ROOT, "museinference.jl_b200", "csrc", "muse_draws.cu")).read()
    a, b = src_cu.index("struct NMRegs {"), src_cu.index("template <bool BOTH>")
    block = src_cu[a:b]
    assert "box_muller_regs" in block and "asm" not in block
    tmpl = open(os.path.join(ROOT, "tests", "csrc", "draws_regs_host.cpp.in")).read()
    src = os.path.join(ROOT, "tests", "csrc", "_draws_regs_host_gen.cpp")
    try:
        with open(src, "w") as fh:
            fh.write(tmpl.replace("@BOX_MULLER_REGS@", block))
        out = str(tmp_path / "libdr.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src], check=True)
    finally:
        if os.path.exists(src):
            os.remove(src)
    lib = C.CDLL(out)
    rng = np.random.default_rng(2026)
    r = rng.integers(0, 2 ** 32, size=(4_000_000, 4), dtype=np.uint64).astype(np.uint32)
    # edge cases: v = ((r1 >> 5) << 26) + (r0 >> 6) at 0, 1, 2⁵³ − 1 (u → 1 after rounding), around 2⁵², around the √2 split of the
    # mantissa and the cell boundaries of the angle table — for both uniforms of a block
    edge = []
    for hi in (0, 1, 31, 32, 2 ** 32 - 1, 2 ** 32 - 32, 2 ** 31, 2 ** 31 - 1, 0x6A09E667, 0x6A09F000, 0x6A09EFFF, 0xB504F333, 2 ** 18, 2 ** 18 - 1):
        for lo in (0, 63, 64, 2 ** 32 - 1, 2 ** 32 - 64, 2 ** 31):
            edge.append((lo, hi, lo ^ 0x5A5A5A5A, hi))
            edge.append((lo ^ 0x12345678, hi ^ 0xFFFF, lo, hi))
    r = np.ascontiguousarray(np.concatenate([r, np.array(edge, dtype=np.uint64).astype(np.uint32)]))
    n = r.shape[0]
    got, ref = np.empty(2 * n), np.empty(2 * n)
    lib.muse_host_box_muller_regs(_vp(r), C.c_int(n), _vp(got))
    lib.muse_host_box_muller_tab(_vp(r), C.c_int(n), _vp(ref))
    assert np.isfinite(ref).all()
    np.testing.assert_array_equal(got.view(np.uint64), ref.view(np.uint64))
