"""Per-instruction shared-memory wavefronts of the LDS instructions of a kernel (.ncu-rep with --import-source on):
executed warp-instructions, wavefronts, ideal wavefronts -- the 64-bit gathers are the in-situ measurement of what
8-byte coordinate planes would cost against the 128-bit {x,y} gathers."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
hd = rows[h]
cols = {n: i for i, n in enumerate(hd)}
wf = [n for n in hd if "Wavefronts Shared" in n]
print("columns:", wf)
src, ie = cols["Source"], cols["Instructions Executed"]
def num(s):
    try: return float(s.replace(",", ""))
    except Exception: return 0.0
agg = collections.OrderedDict()
lines = []
for k, r in enumerate(rows[h + 1:]):
    if not r or r[0] in ("Kernel Name", "Address"): break
    t = r[src].strip()
    if "LDS" not in t and "STS" not in t: continue
    op = t.split()[0] if not t.startswith("@") else t.split()[1]
    vals = [num(r[cols[n]]) for n in wf]
    lines.append((k, t, num(r[ie]), vals))
    a = agg.setdefault(op, [0.0] + [0.0] * len(wf)); a[0] += num(r[ie])
    for j, v in enumerate(vals): a[1 + j] += v
print("%-22s %14s " % ("opcode", "warp-instr") + " ".join("%22s" % n[-22:] for n in wf))
for op, a in agg.items():
    print("%-22s %14.0f " % (op, a[0]) + " ".join("%22.0f" % v for v in a[1:]) + ("   per instr: " + " ".join("%.2f" % (v / a[0]) for v in a[1:]) if a[0] else ""))
print("--- LDS lines with the most wavefronts ---")
for k, t, e, vals in sorted(lines, key=lambda x: -x[3][0] if x[3] else 0)[:24]:
    print("#%-5d %-52s exec %10.0f  " % (k, t[:52], e) + " ".join("%12.0f" % v for v in vals) + ("  per instr %.2f" % (vals[0] / e) if e and vals else ""))
