"""Per-CTA timeline of the decode hot kernel in a dependent stream of launches.

Build bcl_stream_ws.cu with -DBCL_WS_TRACE into a side library, run the bench with BCL_LIB_PATH pointing at it and pipe its output here:

    BCL_LIB_PATH=build_ab/lib_trace.so python bench.py --steps 400 --warmup 10 --repeats 1 --no-detail --no-cpu-baseline | python tools/ws_trace.py

Each traced launch prints one line per (CTA, finisher): `TL launch cta smid finisher t_start t_release t_end segments` (ns, %globaltimer).
The summary answers: how long after the release of griddepcontrol.wait does each CTA finish, and is the spread tied to the SM."""
import sys
from collections import defaultdict

import numpy as np


def main():
    rec = defaultdict(dict)
    for line in sys.stdin:
        if not line.startswith("TL "):
            continue
        n, cta, sm, f, t0, t1, t2, nseg = (int(v) for v in line.split()[1:])
        r = rec[n].setdefault(cta, {"sm": sm, "t0": t0, "t1": t1, "t2": t2, "nseg": nseg})
        r["t0"] = min(r["t0"], t0); r["t1"] = min(r["t1"], t1); r["t2"] = max(r["t2"], t2)
    per_sm = defaultdict(list)
    for n in sorted(rec):
        c = rec[n]
        ids = sorted(c)
        t0 = np.array([c[i]["t0"] for i in ids], dtype=np.float64)
        t1 = np.array([c[i]["t1"] for i in ids], dtype=np.float64)
        t2 = np.array([c[i]["t2"] for i in ids], dtype=np.float64)
        rel = t1.min()
        q = lambda a: "min %.2f p10 %.2f med %.2f p90 %.2f max %.2f" % tuple(np.percentile(a, [0, 10, 50, 90, 100]) / 1e3)
        print(f"launch {n}: {len(ids)} CTAs; us relative to the first release of the wait")
        print("  CTA start     :", q(t0 - rel))
        print("  wait released :", q(t1 - rel))
        print("  CTA end       :", q(t2 - rel))
        print("  step (first release -> last end) %.2f us; mean end %.2f us" % ((t2.max() - rel) / 1e3, (t2 - rel).mean() / 1e3))
        for i in ids:
            per_sm[c[i]["sm"]].append((c[i]["t2"] - rel) / 1e3)
    if per_sm:
        m = {s: np.mean(v) for s, v in per_sm.items()}
        order = sorted(m, key=m.get)
        sd = np.mean([np.std(v) for v in per_sm.values() if len(v) > 1])
        print("per-SM mean end time over the traced launches: spread across SMs %.2f us (std %.2f), mean std within an SM %.2f us" %
              (m[order[-1]] - m[order[0]], np.std(list(m.values())), sd))
        print("  earliest SMs:", [(s, round(m[s], 1)) for s in order[:8]])
        print("  latest SMs  :", [(s, round(m[s], 1)) for s in order[-8:]])


if __name__ == "__main__":
    main()
