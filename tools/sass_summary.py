#!/usr/bin/env python
"""Per-kernel SASS mnemonic table of libveto_b200.so -> profiles/sass_summary.txt (run by __graft_entry__.build()).

What each kernel is built on, read from the shipped cubins (cuobjdump -sass), not from the source: tcgen05 tensor-core
instructions (UTCHMMA = kind::f16, UTCQMMA = kind::f8f6f4, .2CTA = cta_group::2), tensor-memory loads (LDTM), TMA bulk
tensor copies (UTMALDG), warp MMAs (HMMA), asynchronous global->shared copies (LDGSTS), atomics (ATOMS / ATOMG / RED).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "veto_b200", "lib", "libveto_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass_summary.txt")
COLS = [("UTCHMMA", r"^UTCHMMA"), ("UTCQMMA", r"^UTCQMMA"), (".2CTA", r"^UTC[HQ]MMA.*2CTA"), ("LDTM", r"^LDTM"),
        ("UTMALDG", r"^UTMALDG"), ("UTCBAR", r"^UTCBAR"), ("HMMA", r"^HMMA"), ("LDGSTS", r"^LDGSTS"), ("ATOMS", r"^ATOMS"),
        ("ATOMG/RED", r"^(ATOMG|RED)"), ("SHFL", r"^SHFL"), ("MUFU", r"^MUFU"), ("LDG", r"^LDG"), ("STG", r"^STG")]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    if not os.path.exists(LIB):
        print("sass_summary: library not built", file=sys.stderr)
        return 1
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            kernels[cur]["_total"] += 1
            for name, pat in COLS:
                if re.match(pat, m.group(1)):
                    kernels[cur][name] += 1
    names = demangle(list(kernels))

    def short(n):
        d = names.get(n, n)
        d = re.sub(r"\(anonymous namespace\)::", "", d)
        d = re.sub(r"^void ", "", d)
        d = re.sub(r"\(.*", "", d)
        return d.replace("veto::", "")[:46]

    with open(OUT, "w") as f:
        f.write(f"# cuobjdump -sass veto_b200/lib/libveto_b200.so : {len(kernels)} kernels, cubin arch {', '.join(arch)}\n")
        f.write("# instruction counts per kernel (static SASS): tcgen05 MMA (UTCHMMA kind::f16, UTCQMMA kind::f8f6f4; .2CTA = cta_group::2),\n"
                "# LDTM tcgen05.ld, UTMALDG TMA loads, UTCBAR tcgen05.commit, HMMA mma.sync, LDGSTS cp.async\n")
        f.write(f"{'kernel':46s} {'instr':>6s} " + " ".join(f"{c:>9s}" for c, _ in COLS) + "\n")
        tot = collections.Counter()
        for k, c in kernels.items():
            f.write(f"{short(k):46s} {c['_total']:6d} " + " ".join(f"{c[n]:9d}" for n, _ in COLS) + "\n")
            tot.update(c)
        f.write(f"{'TOTAL':46s} {tot['_total']:6d} " + " ".join(f"{tot[n]:9d}" for n, _ in COLS) + "\n")
    print(f"sass_summary: {len(kernels)} kernels -> {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
