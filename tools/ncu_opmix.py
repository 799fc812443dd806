#!/usr/bin/env python3
"""Dynamic SASS opcode mix of one kernel from an .ncu-rep (source page, per-instruction execution counts).
usage: tools/ncu_opmix.py report.ncu-rep [kernel-index]"""
import csv, collections, re, subprocess, sys
def main(path, which=0):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = out.split('"Kernel Name",')[1:]
    blk = blocks[which]
    lines = blk.splitlines()
    print("kernel:", lines[0].strip('",'))
    rows = list(csv.reader(lines[1:]))
    hdr = rows[0]
    isrc, iex, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    mix, samples = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= iex: continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[isrc])
        if not m: continue
        op = m.group(1)
        mix[op] += int(r[iex]); samples[op] += int(r[ismp] or 0)
    tot = sum(mix.values()); ts = sum(samples.values())
    print("total warp-instructions %d, stall samples %d" % (tot, ts))
    for op, c in mix.most_common(24):
        print("  %-22s %6.2f%%  samples %5.2f%%" % (op, 100.0 * c / tot, 100.0 * samples[op] / max(ts, 1)))
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
