"""Small run that touches every kernel (for compute-sanitizer): tiny model, lattice mode, partials, a resampled stream, a stream
that is cut by rule 5, more streams than lanes.  python tools/sanitize_run.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402

import vbmodel  # noqa: E402
import vosk  # noqa: E402

vosk.SetLogLevel(0)
with tempfile.TemporaryDirectory() as td:
    mdir = vbmodel.write_model_dir(td, "tiny", 0)
    model = vosk.BatchModel(mdir, options="num-channels=4,max-batch-size=3,max-seconds=24,partials=1")
    waves = [(vbmodel.synth_audio(1.7, 11), 16000.0), (vbmodel.synth_audio(0.9, 12), 16000.0), (vbmodel.synth_audio(21.5, 13), 16000.0),
             (vbmodel.synth_audio(1.2, 14)[::2].copy(), 8000.0), (vbmodel.synth_audio(0.02, 15), 16000.0)]
    recs = [vosk.BatchRecognizer(model, rate) for _, rate in waves]
    pos = [0] * len(recs)
    live = set(range(len(recs)))
    while live:
        for i in sorted(live):
            w = waves[i][0]
            piece = w[pos[i]:pos[i] + 4000]
            pos[i] += 4000
            if len(piece):
                recs[i].AcceptWaveform(piece.tobytes())
            else:
                recs[i].FinishStream()
                live.discard(i)
        recs[0].PartialResult()
    model.Wait()
    n = 0
    for r in recs:
        while True:
            t = r.Result()
            if not t:
                break
            n += 1
    st = model.Stats()
    print("results", n, "launches", int(st["launches"]), "truncated", st["truncated"], "lattice fallbacks", st["lattice_fallbacks"])
    assert n >= len(recs) + 1 and st["truncated"] == 0
    del recs, model
print("sanitize_run ok")
