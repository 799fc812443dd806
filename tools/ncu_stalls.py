#!/usr/bin/env python3
"""Per-function stall breakdown of one kernel from an .ncu-rep (source page): splits the SASS at RET/EXIT
boundaries (one region per non-inlined device function) and sums executed instructions and stall samples.
usage: tools/ncu_stalls.py report.ncu-rep"""
import csv, re, subprocess, sys, collections
def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = out.splitlines()
    rows = list(csv.reader(lines[1:]))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    regions, cur = [], {"n": 0, "ex": 0, "wide": 0, "st": collections.Counter(), "first": None}
    for r in rows[1:]:
        if len(r) < len(hdr): continue
        src = r[col["Source"]]
        ex = int(r[col["Instructions Executed"]] or 0)
        cur["n"] += 1; cur["ex"] += ex
        if cur["first"] is None: cur["first"] = r[col["Address"]]
        if "IMAD.WIDE" in src: cur["wide"] += ex
        for h in stall_cols: cur["st"][h] += int(r[col[h]] or 0)
        if re.search(r"\bRET\b|^\s*EXIT", src) and not src.strip().startswith("@"):
            regions.append(cur); cur = {"n": 0, "ex": 0, "wide": 0, "st": collections.Counter(), "first": None}
    if cur["n"]: regions.append(cur)
    tot_ex = sum(x["ex"] for x in regions); tot_s = sum(sum(x["st"].values()) for x in regions)
    print("regions (static instrs, share of executed, share of samples, MAC share, top stalls)")
    for x in regions:
        s = sum(x["st"].values())
        top = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(s, 1)) for k, v in x["st"].most_common(5))
        print("  %5d instrs  exec %5.1f%%  samples %5.1f%%  wide %4.1f%%  | %s" % (x["n"], 100.0 * x["ex"] / tot_ex, 100.0 * s / max(tot_s, 1), 100.0 * x["wide"] / max(x["ex"], 1), top))
if __name__ == "__main__":
    main(sys.argv[1])
