"""SASS opcode histogram of every kernel in the built objects (build/csrc/*.o), for profiles/: shows at a glance that the matcher
is tcgen05/TMEM/TMA code (UTCHMMA, LDTM, UTMALDG, UTCBAR), the BA kernels bulk-copy staged fp64, and so on.
    python tools/sass_histogram.py matcher_tc ba remap > profiles/r02_sass_histogram.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
units = sys.argv[1:] or ["matcher_tc", "ba", "remap"]
for u in units:
    obj = os.path.join(ROOT, "build", "csrc", u + ".o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    fn, hist = None, None
    out = []
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            if fn: out.append((fn, hist))
            fn, hist = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn: hist[m.group(1)] += 1
    if fn: out.append((fn, hist))
    print(f"==== {u}.o (nvcc -gencode arch=compute_100a,code=sm_100a)")
    for fn, hist in out:
        name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
        tot = sum(hist.values())
        key = [k for k in ("UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "UCGABAR_ARV", "DFMA", "FFMA2", "FMNMX", "FMNMX3", "LDGSTS") if hist.get(k)]
        print(f"-- {name}: {tot} instructions; " + ", ".join(f"{k} {hist[k]}" for k in key))
        print("   " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(14)))
