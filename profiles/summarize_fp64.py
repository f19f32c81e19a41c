"""FP64 instruction counts per kernel of the two product builds (generic / fma sincosf restatement); run from the repo root."""
import subprocess, re, collections, sys
out = []
for l in ("libplayrho_b200", "libplayrho_b200_fma"):
    txt = subprocess.run(["cuobjdump", "-sass", "playrho_b200/%s.so" % l], capture_output=True, text=True).stdout
    cnt = collections.defaultdict(collections.Counter)
    f = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            f = re.sub(r"^_ZN2pr\d+", "", m.group(1)); f = re.sub(r"ENS_.*$|E[a-zA-Z0-9_]*$", "", f) if f.startswith("k_") else f
            continue
        m = re.search(r"\b(DFMA|DMUL|DADD)\b", line)
        if m and f:
            cnt[f][m.group(1)] += 1
    for f in sorted(cnt):
        c = cnt[f]
        out.append("%-22s %-34s DADD %4d  DMUL %4d  DFMA %4d  total %4d" % (l, f, c["DADD"], c["DMUL"], c["DFMA"], sum(c.values())))
print("\n".join(sorted(out, key=lambda r: (r.split()[1], r.split()[0]))))
