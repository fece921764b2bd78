"""Hot spots of an `ncu --page source --csv` dump: stall samples per SASS instruction, grouped by region.
    python tools/ncu_hot.py gpurun_out/x_src.csv [top_n]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
base = int(data[0][0], 16)
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("total samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out = []
for r in data:
    n = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[s]] or 0), s) for s in stalls), reverse=True)[:2]
    out.append((int(r[0], 16) - base, n, r[1].strip(), st, int(r[col["Instructions Executed"]] or 0)))
for off, n, src, st, ex in sorted(out, key=lambda t: -t[1])[:top]:
    print("%6x %6d %5.1f%% ex=%9d %-55s %s" % (off, n, 100.0 * n / tot, ex, src[:55], " ".join("%s=%d" % (s[6:], v) for v, s in st if v)))
# coarse profile: samples per 0x400 bytes of code
print("--- samples by 1 KB code block")
blk = {}
for off, n, *_ in out:
    blk[off // 0x400] = blk.get(off // 0x400, 0) + n
for b in sorted(blk):
    if blk[b] > tot * 0.004:
        print("%6x %6d %5.1f%%" % (b * 0x400, blk[b], 100.0 * blk[b] / tot))
