"""Aggregate the warp-stall samples of an `ncu --page source --csv` dump: share per stall reason, the instructions
that collect most samples, and the opcode mix."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    for i, r in enumerate(rows):
        if r and r[0] == "Address":
            hdr, data = r, rows[i + 1:]
            break
    if hdr is None:
        print("no source page")
        return
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    for r in data:
        for s in stalls:
            tot[s] += int(r[ix[s]] or 0)
    total = sum(tot.values()) or 1
    print("# warp-stall samples by reason (all samples %d)" % total)
    for s, v in tot.most_common(12):
        print("  %-26s %6.1f %%" % (s, 100.0 * v / total))
    print("# instructions with the most samples")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
        top = max(stalls, key=lambda s: int(r[ix[s]] or 0))
        print("  %6s samples  %-22s %s" % (r[ix["# Samples"]], top, r[1][:80]))
    ops = collections.Counter()
    for r in data:
        toks = [t for t in r[1].split() if not t.startswith("@")]
        if toks:
            ops[toks[0].rstrip(";").split(".")[0]] += int(r[ix["Instructions Executed"]] or 0)
    tt = sum(ops.values()) or 1
    print("# opcode mix (executed warp instructions)")
    print("  " + "  ".join("%s %.1f%%" % (k, 100.0 * v / tt) for k, v in ops.most_common(14)))


if __name__ == "__main__":
    main(sys.argv[1])
