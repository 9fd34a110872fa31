"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/launch_summary.py file.csv [out.txt]"""
import collections
import csv
import re
import sys


def main(path, out=None, step=None):
    """step = k: only the launches of the k-th step of the run (a step starts at mr::k_build_init (round 1: mr::k_elements), the first kernel of the
    LBVH rebuild); bench.py's default run is 3 eager warm-ups, 1 capture warm-up, W graph replays, K timed replays, ..."""
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    agg = collections.OrderedDict()
    body = rows[hdr + 1:]
    if step is not None:
        starts = [i for i, r in enumerate(body) if len(r) > ki and ("k_build_init" in r[ki] or "k_elements" in r[ki])] + [len(body)]
        body = body[starts[step]:starts[step + 1]]
        path = "%s [step %d of %d]" % (path, step, len(starts) - 1)
    for r in body:
        if len(r) <= vi:
            continue
        name = r[ki]
        m = re.search(r"&mr::(\w+)", name)
        name = "k_foreach<%s>" % m.group(1) if m else re.sub(r"\(.*", "", name)[:60]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = ["# %s: %d launches, %.1f us total device time (cold-cache, serialised)" % (path, sum(a[0] for a in agg.values()), tot)]
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append("%9.1f us %5.1f%% x%-4d %8.1f us/launch  %s" % (t, 100 * t / tot, c, t / c, n))
    print("\n".join(lines))
    if out:
        open(out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None,
         int(sys.argv[3]) if len(sys.argv) > 3 else None)
