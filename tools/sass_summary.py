"""Per-kernel counts of the Blackwell-native SASS mnemonics in the shipped library -> profiles/r02_sass_summary.txt."""
import collections, hashlib, re, subprocess, sys, datetime
so = "pytorch_nested-unet_b200/libnunet_b200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDGSTS", "HMMA")
counts, fn = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    for n in names:
        if re.search(r"\b%s\b" % n, line) or re.search(r"\b%s\." % n, line):
            counts[fn][n] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# cuobjdump -sass %s  (sha256 %s, %s)" % (so, hashlib.sha256(open(so, "rb").read()).hexdigest()[:16], datetime.date.today()))
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,")
print("# LDGSTS = cp.async, HMMA = legacy mma.sync (none).  No UTMASTG: the conv epilogues are bound by HBM WRITE bandwidth (3.85 TB/s write-only")
print("# vs 6.4 TB/s copy, profiles/r02_membw.log), not by store issue — see DESIGN.md section 3.1.")
tot = collections.Counter()
for (fn, c), d in zip(counts.items(), dem):
    tot.update(c)
    if any(c[n] for n in names):
        print("%-110s %s" % (re.sub(r"nunet::", "", d)[:110], " ".join("%s=%d" % (n, c[n]) for n in names if c[n])))
print("TOTAL " + " ".join("%s=%d" % (n, tot[n]) for n in names))
