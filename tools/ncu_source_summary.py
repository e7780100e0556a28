"""Summarise an `ncu --page source --csv` export: stall reasons overall and per code region (regions are
split at BAR.SYNC instructions), and the hottest SASS instructions."""
import csv, sys, collections
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = collections.Counter(); reg = []; cur = collections.Counter(); cur_inst = 0; cur_samp = 0
insts = []
region_id = 0
for r in rows[2:]:
    if len(r) < len(hdr): continue
    src = r[col["Source"]]
    samp = int(r[col["# Samples"]] or 0)
    ex = int(r[col["Instructions Executed"]] or 0)
    for s in stalls:
        v = int(r[col[s]] or 0); tot[s] += v; cur[s] += v
    cur_inst += ex; cur_samp += samp
    insts.append((samp, ex, src.strip(), region_id, r[col["Address"]]))
    if "BAR.SYNC" in src or "EXIT" in src:
        reg.append((region_id, cur_inst, cur_samp, cur)); cur = collections.Counter(); cur_inst = 0; cur_samp = 0; region_id += 1
reg.append((region_id, cur_inst, cur_samp, cur))
T = sum(tot.values())
NI = sum(i[1] for i in insts)
print("total samples", T, "total warp insts", NI)
for s, v in tot.most_common(10): print(f"  {s:28s} {v:8d} {100*v/T:5.1f}%")
print("regions (split at BAR.SYNC/EXIT):")
for rid, ni, ns, c in reg:
    if ns < T * 0.005: continue
    top = ", ".join(f"{k[6:]}={100*v/max(1,ns):.0f}%" for k, v in c.most_common(4))
    print(f"  region {rid:3d}: insts {ni:10d} ({100*ni/NI:4.1f}%) samples {ns:8d} ({100*ns/T:4.1f}%)  {top}")
print("hottest instructions:")
for samp, ex, src, rid, addr in sorted(insts, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"  r{rid:<3d} {samp:7d} {100*samp/T:4.1f}%  ex={ex:9d}  {src[:90]}")
