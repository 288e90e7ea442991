"""Summarise `ncu -i X.ncu-rep --page source --csv` output: per kernel, instruction mix, stall samples, hot spots."""
import collections
import csv
import sys


def main(path, pattern=None, topn=18):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for kn in kernels:
        if pattern and pattern not in kn["name"]:
            continue
        hdr, data = kn["hdr"], kn["data"]
        si, ii, so = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
        tot_s = sum(int(r[si]) for r in data) or 1
        tot_i = sum(int(r[ii]) for r in data) or 1
        print("=" * 100)
        print(kn["name"][:160])
        print("SASS lines", len(data), "samples", tot_s, "warp-instructions", tot_i)
        op, ops = collections.Counter(), collections.Counter()
        for r in data:
            t = r[so].split()
            o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            o = o.split(".")[0]
            op[o] += int(r[ii])
            ops[o] += int(r[si])
        for k, v in op.most_common(topn):
            print(f"  {k:12s} inst {100 * v / tot_i:5.1f}%  samples {100 * ops[k] / tot_s:5.1f}%")
        cum = 0
        print("  -- markers (index, instr, samples, executed, cumulative sample %)")
        for i, r in enumerate(data):
            o = r[so]
            cum += int(r[si])
            if any(x in o for x in ["UTCHMMA", "LDTM", "UBLKCP", "BAR.SYNC", "SYNCS", "UTCBAR", "EXIT", "UTCATOMSWS"]):
                print(f"  {i:5d} {o.strip()[:64]:64s} {r[si]:>6s} {r[ii]:>9s} {100 * cum / tot_s:6.1f}")
        print("  -- top sampled instructions")
        for r in sorted(data, key=lambda r: -int(r[si]))[:12]:
            print(f"  {r[si]:>6s} {r[so].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
