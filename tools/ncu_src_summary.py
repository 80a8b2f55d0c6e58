#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` export: stall samples per code region (regions are
cut at BAR / mbarrier waits), top instructions by samples, shared-memory excess wavefronts.
usage: tools/ncu_src_summary.py src.csv [topN]"""
import csv
import sys


def main():
    path = sys.argv[1]
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]

    def num(r, k):
        try:
            return float(r[ix[k]])
        except Exception:
            return 0.0

    total = sum(num(r, "# Samples") for r in data)
    print(f"instructions {len(data)}  samples {total:.0f}")
    regions, cur = [], {"start": 0, "n": 0, "samples": 0.0, "exec": 0.0, "stalls": {}, "label": "prologue"}
    for i, r in enumerate(data):
        src = r[ix["Source"]]
        cur["n"] += 1
        cur["samples"] += num(r, "# Samples")
        cur["exec"] = max(cur["exec"], num(r, "Instructions Executed"))
        for c in stall_cols:
            cur["stalls"][c] = cur["stalls"].get(c, 0.0) + num(r, c)
        if src.startswith("BAR") or "SYNCS.PHASECHK" in src or src.startswith("EXIT"):
            regions.append(cur)
            cur = {"start": i + 1, "n": 0, "samples": 0.0, "exec": 0.0, "stalls": {}, "label": src[:28]}
    regions.append(cur)
    print("\nregions (cut after BAR / mbarrier wait):")
    for g in regions:
        if g["samples"] < 0.002 * total:
            continue
        top = sorted(g["stalls"].items(), key=lambda kv: -kv[1])[:4]
        tops = " ".join(f"{k[6:]}={v:.0f}" for k, v in top if v > 0)
        print(f"  @{g['start']:5d} n={g['n']:4d} after[{g['label']:28s}] samples={g['samples']:6.0f} ({100*g['samples']/total:4.1f}%) exec={g['exec']:.0f}  {tops}")
    print("\ntop instructions:")
    order = sorted(range(len(data)), key=lambda i: -num(data[i], "# Samples"))[:topn]
    for i in order:
        r = data[i]
        top = sorted(((c, num(r, c)) for c in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"  #{i:5d} {num(r, '# Samples'):6.0f}  {r[ix['Source']][:70]:70s} {top[0][0][6:]}={top[0][1]:.0f} {top[1][0][6:]}={top[1][1]:.0f}")
    ex = [(num(r, "L1 Wavefronts Shared Excessive"), i) for i, r in enumerate(data)]
    tot_ex = sum(e for e, _ in ex)
    tot_w = sum(num(r, "L1 Wavefronts Shared") for r in data)
    print(f"\nshared wavefronts {tot_w:.0f}, excessive {tot_ex:.0f}")
    for e, i in sorted(ex, reverse=True)[:12]:
        if e > 0:
            print(f"  #{i:5d} excess={e:.0f} of {num(data[i], 'L1 Wavefronts Shared'):.0f}  {data[i][ix['Source']][:60]}")


if __name__ == "__main__":
    main()
