import csv, subprocess, sys, io, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
hd = rows[h]
si, src, ie, te = hd.index("# Samples"), hd.index("Source"), hd.index("Instructions Executed"), hd.index("Thread Instructions Executed")
mix = collections.Counter(); smp = collections.Counter(); thr = collections.Counter()
tot = 0
lines = []
for r in rows[h + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"): break
    t = r[src].strip().split()
    op = t[1] if t[0].startswith("@") else t[0]
    op = op.split(".")[0]
    n = int(r[ie] or 0); mix[op] += n; smp[op] += int(r[si] or 0); thr[op] += int(r[te] or 0); tot += n
    lines.append((n, r[src].strip()))
nw = max(n for n, _ in lines[:5])
print("warps", nw, "instr/warp", tot / nw)
for op, n in mix.most_common(28):
    print("%-10s %8.1f /warp  %5.1f%%  samples %5d  lanes %.1f" % (op, n / nw, 100 * n / tot, smp[op], thr[op] / max(n, 1)))
