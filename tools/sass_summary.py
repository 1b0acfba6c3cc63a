"""Per-kernel SASS mnemonic counts of libern_b200.so (evidence that the hot kernels are tcgen05 / TMEM / TMA code):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM
alloc, SYNCS = mbarrier, HMMA = legacy mma.sync, LDGSTS = cp.async, REDUX = warp reduce."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fashionern_aaai2024_b200", "libern_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "LDGSTS", "LDSM",
         "REDUX", "FFMA", "MUFU", "ATOMG", "RED", "LDG", "STG", "BAR")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Za-z0-9_]+)*)", line)
        if m and cur:
            kernels[cur]["_total"] += 1
            op = m.group(1)
            if op in WATCH:
                kernels[cur][op] += 1
                mods = m.group(2)
                if op in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR") and mods:
                    kernels[cur][op + mods] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    for name, c in kernels.items():
        d = demangle(name)
        d = re.sub(r"\(.*", "", d)[:150]
        keys = [k for k in c if k != "_total"]
        base = [k for k in keys if "." not in k]
        detail = [k for k in keys if "." in k]
        print(f"\n{d}\n  instructions: {c['_total']}")
        print("  " + "  ".join(f"{k}={c[k]}" for k in sorted(base, key=lambda k: WATCH.index(k))))
        if detail:
            print("  " + "  ".join(f"{k}={c[k]}" for k in sorted(detail)))


if __name__ == "__main__":
    main()
