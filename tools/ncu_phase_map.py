"""Summarises an `ncu --page source --csv` dump: stall samples between marker instructions (barriers, TMEM loads,
UMMA/TMA issue, global stores) so the phases of a warp-specialised kernel can be read off.
Usage: python tools/ncu_phase_map.py src.csv [kernel-substring]"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ""
kern, data, hdr = None, {}, None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        kern = r[1]; data[kern] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and kern and len(r) == len(hdr):
        data[kern].append(dict(zip(hdr, r)))
MARK = ["LDTM", "BAR.SYNC", "SYNCS.PHASECHK", "SYNCS.ARRIVE", "UTCBAR", "STG", "ATOMS", "FENCE", "RED", "LDG", "EXIT",
        "ELECT"]
for k, v in data.items():
    if want not in k:
        continue
    print("=====", k[:60])
    tot = sum(int(x["# Samples"] or 0) for x in v)
    acc = n = 0
    last = None
    for x in v:
        src = x["Source"]; s = int(x["# Samples"] or 0)
        if any(t in src for t in ("UTCHMMA", "UTMALDG")):
            acc += s; n += 1; continue
        if any(t in src for t in MARK):
            if acc > 0.004 * tot:
                print("   ... %d instr, %.1f%%" % (n, 100 * acc / tot))
            acc = n = 0
            if s > 0.002 * tot or "SYNCS" in src or "LDTM" in src or "BAR" in src:
                print("%s %.1f%% %s" % (x["Address"][-5:], 100 * s / tot, src.strip()[:70]))
        else:
            acc += s; n += 1
    print("   ... %d instr, %.1f%%" % (n, 100 * acc / tot))
