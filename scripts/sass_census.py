"""SASS opcode census of libmuse_b200.so, per kernel (needs no GPU: `cuobjdump -sass` on the built library).

    python scripts/sass_census.py [profiles/r02_sass_census.txt]

Counts, for every kernel in the library, the instructions that prove which hardware path it takes
(/opt/skills/guides/B200_PROFILING.md): UBLKCP (cp.async.bulk), UTMALDG / UTMASTG (TMA tensor loads / stores), SYNCS
(mbarrier), DMMA (FP64 tensor core), LDGSTS (cp.async), REDUX, LDG/STG widths, DFMA/DADD/DMUL, and UTC*MMA / LDTM (tcgen05 —
expected absent: tcgen05 has no f64 kind)."""
import collections
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "museinference.jl_b200", "libmuse_b200.so")
COLS = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DMMA", "LDGSTS", "REDUX", "LDG.E.128", "STG.E.128", "LDS.128", "DFMA", "DADD", "DMUL",
        "MUFU", "ATOM", "RED", "UTCMMA", "LDTM", "BAR", "MEMBAR", "ERRBAR"]


def src_sha256():
    h = hashlib.sha256()
    cs = os.path.join(ROOT, "museinference.jl_b200", "csrc")
    for f in sorted(os.listdir(cs)) + [os.path.join("..", "..", "include", "muse_b200.h")]:
        with open(os.path.join(cs, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else None
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            c = kernels[cur]
            c["_total"] += 1
            for col in COLS:
                if op == col or op.startswith(col + ".") or (col in ("UTCMMA",) and op.startswith("UTC") and "MMA" in op):
                    c[col] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = [d[:d.rfind("(")].replace("(int)", "").replace("void ", "").replace("muse::", "").replace("<unnamed>::", "") for d in demangle]
    used = [c for c in COLS if any(k[c] for k in kernels.values())] + ["UTMALDG", "UTCMMA", "LDTM"]
    used = list(dict.fromkeys(used))
    lines = [f"# SASS census of libmuse_b200.so — arch {', '.join(arch)}; sources sha256 {src_sha256()[:16]}",
             "# produced by scripts/sass_census.py (cuobjdump -sass); counts are static instructions per kernel",
             "kernel".ljust(46) + "".join(c.rjust(10) for c in ["instrs"] + used)]
    for (mangled, c), nm in zip(kernels.items(), names):
        lines.append(nm[:45].ljust(46) + str(c["_total"]).rjust(10) + "".join(str(c[u]).rjust(10) for u in used))
    text = "\n".join(lines) + "\n"
    if out:
        with open(out, "w") as fh:
            fh.write(text)
    print(text)


if __name__ == "__main__":
    main()
