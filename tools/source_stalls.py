"""Per-instruction warp-stall summary of one kernel from an `ncu --set full --import-source on` report:

    ncu -i gpurun_out/<capture>.ncu-rep --page source --csv > /tmp/src_page.csv
    python tools/source_stalls.py /tmp/src_page.csv > profiles/<name>_source_stalls.txt

Prints the stall-reason shares, the top instructions by stall samples with their dominant reason, shares by opcode, and
the shared-memory / L2 sector efficiency columns (how profiles/r01_spmm_group_v3_source_stalls.txt was made)."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    for b in blocks:
        h = b["hdr"]
        ix = {n: i for i, n in enumerate(h)}

        def f(r, n):
            try:
                return float(r[ix[n]])
            except (ValueError, IndexError, KeyError):
                return 0.0

        tot = sum(f(r, "# Samples") for r in b["rows"]) or 1.0
        inst = sum(f(r, "Instructions Executed") for r in b["rows"]) or 1.0
        print("kernel: %s" % b["name"])
        print("SASS instructions: %d, warp-level instructions executed: %.0f, stall samples: %.0f\n" % (len(b["rows"]), inst, tot))
        reasons = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        print("stall reasons over the whole kernel (share of samples):")
        for n, v in sorted(((n, sum(f(r, n) for r in b["rows"])) for n in reasons), key=lambda kv: -kv[1]):
            if v > 0:
                print("  %-24s %6.1f %%" % (n, 100 * v / tot))
        print("\ntop %d instructions by stall samples (share, cumulative, dominant reason, executed count, SASS):" % top)
        cum = 0.0
        for r in sorted(b["rows"], key=lambda r: -f(r, "# Samples"))[:top]:
            s = f(r, "# Samples")
            cum += s
            dom = max(reasons, key=lambda n: f(r, n))
            print("  %5.1f %%  %5.1f %%  %-16s %9.0f  %s" % (100 * s / tot, 100 * cum / tot, dom.replace("stall_", ""),
                                                           f(r, "Instructions Executed"), r[ix["Source"]].strip()))
        cls = {}
        for r in b["rows"]:
            words = r[ix["Source"]].split()
            op = (words[1] if words and words[0].startswith("@") and len(words) > 1 else (words[0] if words else "?")).split(".")[0]
            c = cls.setdefault(op, [0.0, 0.0])
            c[0] += f(r, "# Samples")
            c[1] += f(r, "Instructions Executed")
        print("\nby opcode (share of stall samples, share of executed instructions):")
        for op, (s, e) in sorted(cls.items(), key=lambda kv: -kv[1][0])[:14]:
            print("  %-10s %5.1f %%  %5.1f %%" % (op, 100 * s / tot, 100 * e / inst))
        w, wi = (sum(f(r, n) for r in b["rows"]) for n in ("L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"))
        g, gi = (sum(f(r, n) for r in b["rows"]) for n in ("L2 Theoretical Sectors Global", "L2 Theoretical Sectors Global Ideal"))
        print("\nshared-memory wavefronts %.0f (ideal %.0f); L2 theoretical sectors for global accesses %.0f (ideal %.0f)\n"
              % (w, wi, g, gi))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
