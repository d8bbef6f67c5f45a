"""Read an `ncu --page source --csv` dump (SASS view): per-instruction executed counts and stall samples.
usage: python tools/ncu_src.py <source.csv> [top=25] [--list lo hi]   (lo/hi = instruction index range to list)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {k: i for i, k in enumerate(hdr)}
ins = []
for r in rows[hdr_i + 1:]:
    if len(r) < len(hdr):
        continue
    f = lambda k: float(r[col[k]].replace(",", "") or 0)
    stalls = {k[6:]: f(k) for k in hdr if k.startswith("stall_") and "(Not Issued)" not in k}
    ins.append((r[col["Source"]].strip(), f("# Samples"), f("Instructions Executed"), f("Thread Instructions Executed"), stalls))
tot_s = sum(i[1] for i in ins) or 1
tot_e = sum(i[2] for i in ins) or 1
print("instructions: %d static, %.0f executed (warp), samples %d" % (len(ins), tot_e, tot_s))
if "--list" in sys.argv:
    k = sys.argv.index("--list")
    lo, hi = int(sys.argv[k + 1]), int(sys.argv[k + 2])
    for n in range(lo, min(hi, len(ins))):
        s, smp, ex, tex, st = ins[n]
        top = max(st.items(), key=lambda kv: kv[1]) if st else ("", 0)
        print("%5d %-70s exec %5.2f%% smp %5.2f%% thr %4.1f %s" % (n, s[:70], 100 * ex / tot_e, 100 * smp / tot_s, tex / ex if ex else 0, top[0] if top[1] else ""))
else:
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    order = sorted(range(len(ins)), key=lambda n: -ins[n][1])[:top]
    for n in order:
        s, smp, ex, tex, st = ins[n]
        t3 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
        print("%5d %-70s smp %5.2f%% exec %5.2f%%  %s" % (n, s[:70], 100 * smp / tot_s, 100 * ex / tot_e, ", ".join("%s %d" % kv for kv in t3 if kv[1])))
    # opcode histogram weighted by executed count
    h = {}
    for s, smp, ex, tex, st in ins:
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        op = op.rstrip(";")
        h[op] = h.get(op, 0) + ex
    print("executed by opcode: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot_e) for k, v in sorted(h.items(), key=lambda kv: -kv[1])[:24]))
