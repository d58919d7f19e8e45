"""The C-ABI library builds for sm_100a without a GPU, loads, and exports every symbol include/muse_b200.h
declares; without a CUDA device it fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "muse_b200.h")) as fh:
        src = fh.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(muse_b200_\w+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound(lib_built):
    import museinference_jl_b200 as m
    from importlib import import_module
    capi = import_module(m.__name__ + "._capi")
    declared = _declared_symbols()
    assert len(declared) >= 20
    raw = C.CDLL(m.library_path())
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/muse_b200.h but not exported"
    assert sorted(capi.SIGNATURES) == declared, "ctypes binding and header disagree"
    assert raw.muse_b200_abi_version() == capi.ABI_VERSION


def test_cfg_struct_layout_matches_header():
    import museinference_jl_b200 as m
    from importlib import import_module
    capi = import_module(m.__name__ + "._capi")
    with open(os.path.join(ROOT, "include", "muse_b200.h")) as fh:
        src = fh.read()
    body = re.search(r"typedef struct muse_cfg \{(.*?)\} muse_cfg;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"(\w+)\s*;", body)
    assert names == [n for n, _ in capi.muse_cfg._fields_]


def test_library_is_built_for_sm_100a_only(lib_built):
    import subprocess
    import museinference_jl_b200 as m
    out = subprocess.run(["cuobjdump", "--list-elf", m.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cuda_device_fails_loudly(lib_built):
    """On a box without a GPU create() must return ENODEVICE (and never compute on the CPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import museinference_jl_b200 as m
    with pytest.raises(m.MuseBackendError) as ei:
        m.B200Backend("funnel", 64, 4)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_unregistered_family_raises():
    import museinference_jl_b200 as m
    with pytest.raises(m.MuseBackendError) as ei:
        m.SimpleMuseProblem([0.0, 1.0], "turing-model")
    assert ei.value.code == -5

    class Other(m.AbstractMuseProblem):
        pass
    with pytest.raises(m.MuseBackendError):
        m.muse(Other(), [0.0])


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "museinference.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    txt = fh.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "muse_oracle" not in txt, f


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints exactly one JSON line on stdout with the contract's keys (CPU only)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--d", "2048", "--nsims", "64"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sims/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_struct_layouts_and_constants_match_a_c_compilation_of_the_header(tmp_path):
    """Compile include/muse_b200.h with gcc (as C, so the header must be plain C) and compare sizeof / offsetof of every
    struct that crosses the boundary, and the numeric constants, with the ctypes binding."""
    import subprocess
    import museinference_jl_b200 as m
    from importlib import import_module
    capi = import_module(m.__name__ + "._capi")
    structs = {"muse_cfg": capi.muse_cfg, "muse_profile": capi.muse_profile, "muse_pass_profile": capi.muse_pass_profile,
               "muse_iterate_out": capi.muse_iterate_out, "muse_cov_out": capi.muse_cov_out}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "muse_b200.h"', 'int main(void) {']
    for name, st in structs.items():
        lines.append(f'  printf("sizeof {name} %zu\\n", sizeof({name}));')
        for fname, _ in st._fields_:
            lines.append(f'  printf("offsetof {name}.{fname} %zu\\n", offsetof({name}, {fname}));')
    consts = ["MUSE_B200_ABI_VERSION", "MUSE_FAMILY_FUNNEL", "MUSE_FAMILY_HIERGAUSS", "MUSE_FAMILY_CORRGAUSS", "MUSE_FAMILY_TWOLAYER", "MUSE_START_ZEROS",
              "MUSE_START_PREV", "MUSE_START_TRUTH", "MUSE_START_USER", "MUSE_STATUS_G_CONVERGED", "MUSE_STATUS_XF_CONVERGED",
              "MUSE_STATUS_MAXITER", "MUSE_STATUS_LS_FAILED", "MUSE_STATUS_NONFINITE", "MUSE_PASS_KINDS", "MUSE_EUNSUPPORTED",
              "MUSE_ENODEVICE", "MUSE_ESTATE"]
    for c in consts:
        lines.append(f'  printf("const {c} %d\\n", (int)({c}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for line in out.splitlines():
        kind, key, val = line.split()
        got[(kind, key)] = int(val)
    for name, st in structs.items():
        assert got[("sizeof", name)] == C.sizeof(st), name
        for fname, _ in st._fields_:
            assert got[("offsetof", f"{name}.{fname}")] == getattr(st, fname).offset, f"{name}.{fname}"
    assert got[("const", "MUSE_B200_ABI_VERSION")] == capi.ABI_VERSION
    assert (got[("const", "MUSE_FAMILY_FUNNEL")], got[("const", "MUSE_FAMILY_HIERGAUSS")], got[("const", "MUSE_FAMILY_CORRGAUSS")],
            got[("const", "MUSE_FAMILY_TWOLAYER")]) == (capi.FAMILY_FUNNEL, capi.FAMILY_HIERGAUSS, capi.FAMILY_CORRGAUSS, capi.FAMILY_TWOLAYER)
    # the oracle's family ids are the header's
    import oracle as O
    assert [f.family_id for f in (O.Funnel, O.HierGauss, O.CorrGauss, O.TwoLayer)] == \
        [got[("const", f"MUSE_FAMILY_{k}")] for k in ("FUNNEL", "HIERGAUSS", "CORRGAUSS", "TWOLAYER")]
    assert [got[("const", f"MUSE_START_{k}")] for k in ("ZEROS", "PREV", "TRUTH", "USER")] == \
        [capi.START_ZEROS, capi.START_PREV, capi.START_TRUTH, capi.START_USER]
    assert [got[("const", f"MUSE_STATUS_{k}")] for k in ("G_CONVERGED", "XF_CONVERGED", "MAXITER", "LS_FAILED", "NONFINITE")] == \
        [capi.STATUS_G_CONVERGED, capi.STATUS_XF_CONVERGED, capi.STATUS_MAXITER, capi.STATUS_LS_FAILED, capi.STATUS_NONFINITE]
    assert got[("const", "MUSE_PASS_KINDS")] == len(capi.PASS_KINDS)
    assert {v: k for k, v in capi.E_NAMES.items()}["EUNSUPPORTED"] == got[("const", "MUSE_EUNSUPPORTED")]
    assert {v: k for k, v in capi.E_NAMES.items()}["ENODEVICE"] == got[("const", "MUSE_ENODEVICE")]


def test_plain_c_example_links_against_the_library_and_fails_loudly_without_a_gpu(lib_built, tmp_path):
    """examples/solve_funnel.c: the ABI used from C99 with nothing but the header and the .so.  Here (no GPU) it must
    report ENODEVICE — exit status 3 — instead of computing anything on the CPU; on a B200 it prints OK."""
    import subprocess
    import torch
    import museinference_jl_b200 as m
    libdir = os.path.dirname(m.library_path())
    exe = tmp_path / "solve_funnel"
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "solve_funnel.c"), "-o", str(exe), "-L", libdir, "-lmuse_b200", "-lm",
                    f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe), "1024", "64"], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert out.returncode == 0 and "OK" in out.stdout, out.stdout + out.stderr
    else:
        assert out.returncode == 3 and "no CPU fallback" in out.stderr, out.stdout + out.stderr


def test_committed_bench_line_satisfies_the_contract():
    """The bench line of the round (profiles/r01_v5_bench.json, produced by `python bench.py` on a B200) carries every key the
    measurement contract names, with consistent values."""
    import json
    with open(os.path.join(ROOT, "profiles", "r01_v5_bench.json")) as fh:
        d = json.loads(fh.read())
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in d, key
    assert d["unit"] == "sims/s" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["n_gpus"] == 1 and d["warmup"] >= 3
    assert "workload" in d["config"] and "model" not in d["config"] and "inputs_larger_than_l2" in d["config"]["l2"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"] > 0
    assert 0.5 < r["frac"] < 1.2 and set(r["passes"]) >= {"cold", "warm", "fiducial", "fd"}
    e = d["e2e"]
    assert e["unit"] == "sims/s" and 0 < e["value"] < d["value"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "sims/s" and c["sample"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    units = d["config"]["units_per_step"]
    assert abs(d["value"] - units / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-9
