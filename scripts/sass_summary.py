"""CPU: SASS evidence of the SHIPPED build -> profiles/sass_<tag>.md.

    python scripts/sass_summary.py [tag]        tag defaults to the short git hash of HEAD (+ "-dirty")

Rebuilds gdkvm_b200/libgdkvm_gdr.so from the tree, disassembles it with `cuobjdump -sass`, and writes per kernel: the
register / spill line of `ptxas -v`, opcode counts of the instructions that prove tcgen05 / TMEM / TMA (B200_PROFILING.md:
UTCHMMA, UTCBAR, LDTM, STTM, UTMALDG, UTMASTG) next to the legacy-path ones, and the first lines of each family.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WHAT = collections.OrderedDict([
    ("UTCHMMA", "tcgen05.mma kind::f16 (5th-gen tensor core, TMEM accumulators)"),
    ("UTCQMMA", "tcgen05.mma other kinds"), ("UTCBAR", "tcgen05.commit -> mbarrier"),
    ("LDTM", "tcgen05.ld (TMEM -> registers)"), ("STTM", "tcgen05.st (registers -> TMEM)"),
    ("UTMALDG", "cp.async.bulk.tensor load (TMA)"), ("UTMASTG", "cp.async.bulk.tensor store (TMA)"),
    ("UTMAPF", "tensormap prefetch"), ("SYNCS", "mbarrier ops"), ("HMMA", "mma.sync (legacy tensor path)"),
    ("LDSM", "ldmatrix"), ("STSM", "stmatrix"), ("BAR", "named barriers"), ("LDS", "ld.shared"), ("STS", "st.shared"),
    ("LDG", "ld.global"), ("STG", "st.global"), ("LDL", "local loads (spills)"), ("STL", "local stores (spills)"),
    ("MUFU", "special function unit"), ("F2FP", "fp32 -> 16-bit pack"), ("FFMA", "fp32 fma"), ("FMUL", "fp32 mul"),
    ("HMUL2", "packed 16-bit mul"), ("ELECT", "elect.sync"), ("R2UR", "vector -> uniform register moves"),
])


def main():
    from gdkvm_b200 import _build
    tag = sys.argv[1] if len(sys.argv) > 1 else None
    if tag is None:
        h = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
        dirty = subprocess.run(["git", "-C", ROOT, "status", "--porcelain", "--", "gdkvm_b200/csrc", "include"], capture_output=True, text=True).stdout.strip()
        tag = h + ("-dirty" if dirty else "")
    res = subprocess.run(_build.nvcc_command(extra=["-Xptxas", "-v"]), capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    ptxas = {}
    cur = None
    for ln in res.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", ln)
        if m:
            cur = m.group(1)
        elif cur and "Used" in ln:
            ptxas[cur] = ptxas.get(cur, "") + ln.split("ptxas info    :")[-1].strip()
        elif cur and "spill" in ln:
            ptxas[cur] = ln.strip() + "; " + ptxas.get(cur, "")
    sass = subprocess.run(["cuobjdump", "-sass", _build.LIB_PATH], capture_output=True, text=True).stdout
    kernels, name = collections.OrderedDict(), None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name and re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
            kernels[name].append(ln.rstrip())
    def demangle(n):      # without the trailing parameter list
        d = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
        depth = 0
        for i in range(len(d) - 1, -1, -1):
            depth += d[i] == ")"
            depth -= d[i] == "("
            if depth == 0 and d[i] == "(":
                return d[:i].replace("void ", "").replace("gdkvm::<unnamed>::", "")
        return d
    out = [f"# SASS summary of build `{tag}` -- `cuobjdump -sass gdkvm_b200/libgdkvm_gdr.so` (sm_100a)", "",
           "Written by `scripts/sass_summary.py` from a fresh build of the tree (nvcc " +
           subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-2].strip() + ").", ""]
    for n, lines in kernels.items():
        ops = collections.Counter()
        for ln in lines:
            m = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m:
                ops[m.group(1).split(".")[0]] += 1
        short = demangle(n)
        out += [f"## `{short}`", "", f"- {len(lines)} instructions; ptxas: {ptxas.get(n, 'n/a')}", ""]
        rows = [(k, ops[k], w) for k, w in WHAT.items() if ops.get(k)]
        if rows:
            out += ["| SASS opcode | count | what it is |", "|---|---|---|"] + [f"| {k} | {c} | {w} |" for k, c, w in rows] + [""]
        for fam in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG"):
            hits = [ln.strip() for ln in lines if re.search(r"\b" + fam + r"\b", ln)]
            if hits:
                out += [f"{fam} ({len(hits)}; first 4):", "```"] + [h.split(" /* 0x")[0].rstrip() for h in hits[:4]] + ["```", ""]
    path = os.path.join(ROOT, "profiles", f"sass_{tag}.md")
    open(path, "w").write("\n".join(out))
    print(path, {demangle(n): len(v) for n, v in kernels.items()})


if __name__ == "__main__":
    main()
