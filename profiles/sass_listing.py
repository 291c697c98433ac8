"""Counts and lists the tcgen05 / TMEM / TMA / mbarrier instructions per kernel of one object file.
usage: python profiles/sass_listing.py build/obj/value_mlp.o csrc/value_mlp.cu > profiles/r02_value_mlp_sass_tcgen05.txt"""
import collections, re, subprocess, sys
obj, src = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC[A-Z0-9_.]+|LDTM[A-Za-z0-9_.]*|STTM[A-Za-z0-9_.]*|UTMA[A-Z0-9_.]+|SYNCS[A-Z0-9_.]*|UTMAPF[A-Z0-9_.]*)\b")
kern, lines, counts = None, collections.OrderedDict(), collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f_]+\w*?cu_[0-9a-f]+\d\d", "", m.group(1))
        kern = m.group(1).split("cu_")[-1][10:] if "cu_" in m.group(1) else m.group(1)
        lines[kern], counts[kern] = [], collections.Counter()
        continue
    if kern is None or "/*" not in ln:
        continue
    m = pat.search(ln)
    if m:
        counts[kern][m.group(1)] += 1
        lines[kern].append(ln.rstrip())
print(f"# tcgen05 / TMA / TMEM instructions in libtarl_b200.so's kernels of {src}")
print(f"# made with: python profiles/sass_listing.py {obj} {src}  (cuobjdump -sass, sm_100a, CUDA 12.9); counts per kernel,")
print("# then the full listing of every matching line.\n")
for k, c in counts.items():
    print(f"{k}: " + (", ".join(f"{n} x{v}" for n, v in sorted(c.items())) if c else "(none: fp32 pipe)"))
for k, ls in lines.items():
    if ls:
        print(f"\n## {k}")
        print("\n".join(ls))
