"""Generates tests/golden/lattice_tiebreak.npz: two raw lattices (small architecture, beam 20 / max-active 300, log-likelihoods
perturbed by seeded noise) on which MinimumBayesRisk's edit-distance recursion meets near-ties: the alignment of a competing
word to one sausage bin or its neighbour is decided by the last bits of sums of products, so a chain compiled with FMA
contraction and one compiled without it print different confidences.  Kaldi's own build has no FMA (-msse -msse2): the expected
texts stored here come from the oracle chain built with -ffp-contract=off (oracle/Makefile).  Found by fuzzing (the search loop
below); the GPU test test_max_active_binds met the same effect first.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in ("vosk-api_b200/tools", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, p))
import oracle  # noqa: E402
import vbmodel  # noqa: E402


def main(model_root="/tmp/vb_bench_models/small_0"):
    if not os.path.exists(os.path.join(model_root, "model", "graph", "HCLG.fst")):
        vbmodel.write_model_dir(model_root, "small", 0)
    model = vbmodel.load_model_dir(os.path.join(model_root, "model"))
    model["conf"]["max-active"] = "300"
    model["conf"]["min-active"] = "50"
    model["conf"]["beam"] = "20"
    rng = np.random.default_rng(5)
    out = {}
    for k, (sec, seed, pick) in enumerate([(2.0, 700, 32), (3.1, 701, 39)]):
        ref = oracle.recognize(model, vbmodel.synth_audio(sec, seed), stages=True, lattice=False)
        for trial in range(40):
            ll = (ref["loglikes"] + rng.normal(0, 3e-4, ref["loglikes"].shape)).astype(np.float32)
            if trial != pick:
                continue
            dec = oracle.decode(model, ll, lattice_beam=6.0)
            lat = dec["lattice"]
            for name in ("src", "dst", "arc", "final_state"):
                out["%s_%d" % (name, k)] = np.asarray(lat[name], dtype=np.int32)
            out["ac_%d" % k] = np.asarray(lat["ac"], dtype=np.float32)
            out["final_cost_%d" % k] = np.asarray(lat["final_cost"], dtype=np.float32)
            out["n_states_%d" % k] = np.int32(len(lat["tok_index"]))
            out["start_%d" % k] = np.int32(oracle.lattice_start(dec))
            out["text_%d" % k] = np.array(oracle.lattice_result(model, dec, 6.0))
    np.savez_compressed(os.path.join(HERE, "lattice_tiebreak.npz"), **out)
    print("wrote lattice_tiebreak.npz")


if __name__ == "__main__":
    main()
