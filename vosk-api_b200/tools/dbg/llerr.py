"""Large-architecture log-likelihood error budget: engine (tc=1 / tc=0) and fp32 oracle against an fp64 numpy forward."""
import sys, os, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path[:0] = [os.path.join(ROOT, "vosk-api_b200"), os.path.join(ROOT, "vosk-api_b200", "tools"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import numpy as np, vbmodel, oracle, helpers, vosk
vosk.SetLogLevel(-1)
td = tempfile.mkdtemp()
mdir = vbmodel.write_model_dir(td, "large", 0, overrides=dict(vocab=3000, succ=8))
model = vbmodel.load_model_dir(mdir)
arch = dict(vbmodel.ARCHS["large"])
w = vbmodel.synth_audio(1.9, 1700)
ref = oracle.recognize(model, w, stages=True)
T = {k: np.asarray(v) for k, v in model["nnet"].items()}
ll64 = vbmodel.nnet_forward_numpy(T, arch, ref["mfcc"], ref["ivectors"], ref["iv_index"])
P = 6016
print("oracle fp32 vs fp64: max %.2e rms %.2e  |ll| max %.1f" % (np.abs(ref["loglikes"] - ll64).max(), np.sqrt(np.mean((ref["loglikes"] - ll64) ** 2)), np.abs(ll64).max()))
for tc in (1, 0):
    got, _ = helpers.run_engine(mdir, [w], options=f"num-channels=2,max-batch-size=2,max-seconds=6,tensor-cores={tc}")
    ll = got[0]["loglikes"].reshape(-1, P)
    print("engine tc=%d vs fp64: max %.2e rms %.2e ; vs oracle fp32: max %.2e ; ivector err %.1e" % (tc, np.abs(ll - ll64).max(), np.sqrt(np.mean((ll - ll64) ** 2)), np.abs(ll - ref["loglikes"]).max(), np.abs(got[0]["ivectors"].reshape(-1, 100) - ref["ivectors"]).max()))
