"""Where the log-likelihood error of long (bench-size) streams comes from: per-stage differences engine vs oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPO = os.path.dirname(ROOT)
for p in (REPO, ROOT, os.path.join(ROOT, "tools"), os.path.join(REPO, "oracle"), os.path.join(REPO, "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, vbmodel, oracle, helpers, vosk
vosk.SetLogLevel(-1)
arch = sys.argv[1] if len(sys.argv) > 1 else "small"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mdir = bench.model_dir(arch)
model = vbmodel.load_model_dir(mdir)
waves = bench.make_audio(n, 0)
P = int(model["cfg"]["num-pdfs"]); D = int(model["cfg"]["ivector-dim"])
refs = [oracle.recognize(model, w, stages=True, lattice=False) for w in waves]
for tc in (1, 0):
    got, _ = helpers.run_engine(mdir, waves, options="num-channels=%d,max-batch-size=%d,max-seconds=18,tensor-cores=%d,lattice=0" % (n, n, tc))
    for i, (g, r) in enumerate(zip(got, refs)):
        ll = g["loglikes"].reshape(-1, P); iv = g["ivectors"].reshape(-1, D)
        dm = np.abs(g["mfcc"] - r["mfcc"]).max(axis=1)
        div = np.abs(iv - r["ivectors"]).max(axis=1)
        dll = np.abs(ll - r["loglikes"]).max(axis=1)
        # the network alone: oracle forward on the ENGINE's features and i-vectors
        ll2 = oracle.nnet_forward(model, g["mfcc"], iv, r["iv_index"])
        dnn = np.abs(ll - ll2).max(axis=1)
        print("tc=%d stream %d (%.1f s): mfcc max %.2e | ivector per chunk %s | ll max %.2e (frame %d of %d) | ll vs oracle-net-on-engine-inputs max %.2e" %
              (tc, i, len(waves[i]) / 16000.0, dm.max(), " ".join("%.1e" % x for x in div), dll.max(), int(dll.argmax()), len(dll), dnn.max()))
        print("   ll err by 50-frame block:", " ".join("%.1e" % dll[k:k + 50].max() for k in range(0, len(dll), 50)))
