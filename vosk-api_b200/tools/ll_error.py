"""Log-likelihood error of the engine (tensor-core and fp32 paths) against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO = os.path.dirname(ROOT)
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
import vbmodel, oracle, helpers, vosk
vosk.SetLogLevel(-1)
for arch in sys.argv[1:] or ["tiny", "small"]:
    root = "/tmp/vb_llerr_" + arch
    mdir = os.path.join(root, "model")
    if not os.path.exists(mdir):
        vbmodel.write_model_dir(root, arch, 0)
    model = vbmodel.load_model_dir(mdir)
    waves = [vbmodel.synth_audio(2.5, 40 + i) for i in range(3)]
    P = int(model["cfg"]["num-pdfs"])
    refs = [oracle.recognize(model, w, stages=True) for w in waves]
    res = {}
    for tc in (0, 1, 2):
        got, _ = helpers.run_engine(mdir, waves, options="num-channels=4,max-batch-size=4,max-seconds=8,tensor-cores=%d" % tc)
        res[tc] = got
        errs = [np.abs(g["loglikes"].reshape(-1, P) - r["loglikes"]).max() for g, r in zip(got, refs)]
        rms = [np.sqrt(np.mean((g["loglikes"].reshape(-1, P) - r["loglikes"]) ** 2)) for g, r in zip(got, refs)]
        same = [g["text"] == r["text"] for g, r in zip(got, refs)]
        print(arch, "tc=%d" % tc, "max abs err", ["%.2e" % e for e in errs], "rms", ["%.2e" % e for e in rms], "text equal", same)
    for tc in (1, 2):
        d = [np.abs(a["loglikes"] - b["loglikes"]).max() for a, b in zip(res[0], res[tc])]
        print(arch, "tc%d vs tc0 max abs diff" % tc, ["%.2e" % e for e in d])
