"""Per-kernel count of the SASS mnemonics that prove tcgen05 / TMEM / TMA use in libvitk.so (no GPU needed).
    python scripts/sass_summary.py [lib] > profiles/r02_sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "vit_torch_b200/libvitk.so"
pats = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "HMMA", "MUFU.EX2"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
name = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    for p in pats:
        if re.search(r"(?<![A-Z])" + re.escape(p), line):
            counts[name][p] += 1
            if p == "UTCHMMA.2CTA":
                break          # counted once, as the 2-CTA form
names = list(counts)
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {lib}: instructions per kernel. UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / st,")
print("# UTMALDG / UTMASTG = TMA load / store, UTMAPF = TMA L2 prefetch, UTCBAR = tcgen05.commit, HMMA = mma.sync, MUFU.EX2 = ex2")
tot = collections.Counter()
for n, d in sorted(zip(names, dem), key=lambda t: t[1]):
    c = counts[n]
    tot.update(c)
    if not any(c[p] for p in pats[:7]):
        continue
    short = re.sub(r"\(.*$", "", d).replace("void ", "")
    print(f"{short:70s} " + " ".join(f"{p}={c[p]}" for p in pats if c[p]))
print("# totals: " + " ".join(f"{p}={tot[p]}" for p in pats))
