"""Compare load-displacement curves written by scripts/collapse_1m.py: python scripts/compare_curves.py a.json b.json"""
import json
import sys

import numpy as np

a, b = (json.load(open(p)) for p in sys.argv[1:3])
same = a["iters"] == b["iters"]
print(f"Newton iterations per step equal: {same}  ({a['iters']} vs {b['iters']})")
for k in ("lout", "un", "peeqplot", "csrplot"):
    x, y = np.array(a[k]), np.array(b[k])
    if x.shape != y.shape:
        print(k, "shape differs", x.shape, y.shape)
        continue
    print(f"{k}: max relative difference {np.abs(x - y).max() / max(np.abs(y).max(), 1e-300):.2e}")
print(f"time: {a['analysis_s']:.1f} s (N={a['world']}) vs {b['analysis_s']:.1f} s (N={b['world']})")
