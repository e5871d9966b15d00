"""Host lattice chain timing on this box: single thread, and T threads side by side (does the pool scale?).
python tools/chain_timing.py [streams] [threads...]"""
import ctypes
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO = os.path.dirname(ROOT)
for p in (REPO, ROOT, os.path.join(ROOT, "tools"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import bench  # noqa: E402
import helpers  # noqa: E402
import oracle  # noqa: E402
import vbmodel  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
threads = [int(x) for x in sys.argv[2:]] or [1, 4, 8, 16]
mdir = bench.model_dir()
model = vbmodel.load_model_dir(mdir)
waves = bench.make_audio(n, 0)
lats = []
for w in waves:
    ref = oracle.recognize(model, w, stages=True, lattice=False)
    dec = oracle.decode(model, ref["loglikes"], lattice_beam=6.0)
    lats.append((dec["lattice"], helpers.oracle_lattice_start(dec)))
print("cpus", os.cpu_count(), "lattice arcs", [len(l[0]["src"]) for l in lats])
for lat, st in lats:
    print(helpers.lattice_text(mdir, lat, st, 6.0, stage=4))


def work(reps, out, k):
    t0 = time.perf_counter()
    for _ in range(reps):
        for lat, st in lats:
            helpers.lattice_text(mdir, lat, st, 6.0, stage=0)
    out[k] = time.perf_counter() - t0


helpers.lattice_text(mdir, lats[0][0], lats[0][1], 6.0, stage=0)
for T in threads:
    out = [0.0] * T
    ts = [threading.Thread(target=work, args=(10, out, k)) for k in range(T)]
    t0 = time.perf_counter()
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    wall = time.perf_counter() - t0
    print("threads %d: %.2f ms per lattice per thread, aggregate %.0f lattices/s" % (T, 1e3 * np.mean(out) / (10 * len(lats)), T * 10 * len(lats) / wall))
