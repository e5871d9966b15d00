"""CPU tests pinning the oracle (oracle/) against every independent check available offline."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_mfcc_matches_torchaudio_golden(oracle_lib):
    g = np.load(os.path.join(HERE, "golden", "mfcc_torchaudio.npz"))
    got = oracle_lib.mfcc(g["wave"])
    assert got.shape == g["mfcc"].shape
    # torchaudio computes in fp32 (its FFT differs from the oracle's double FFT by ~1e-3 of a cepstrum of magnitude ~100)
    assert np.abs(got - g["mfcc"]).max() < 2e-3


def test_mfcc_matches_numpy_fp64(oracle_lib):
    import vbmodel
    for seed, secs in ((1, 0.7), (2, 2.3)):
        w = vbmodel.synth_audio(secs, seed)
        assert np.abs(oracle_lib.mfcc(w) - vbmodel.mfcc_numpy(w)).max() < 2e-4


def test_mfcc_live_torchaudio(oracle_lib):
    torchaudio = pytest.importorskip("torchaudio")
    import torch
    import vbmodel
    w = vbmodel.synth_audio(1.0, 77)
    ref = torchaudio.compliance.kaldi.mfcc(torch.from_numpy(w.astype(np.float32))[None], dither=0.0, num_ceps=40, num_mel_bins=40, low_freq=20,
                                           high_freq=-400, energy_floor=0.0, use_energy=False, sample_frequency=16000).numpy()
    assert np.abs(oracle_lib.mfcc(w) - ref).max() < 2e-3


@pytest.mark.parametrize("n", [0, 1, 399, 400, 401, 559, 560, 8160, 16000])
def test_frame_count_snip_edges(oracle_lib, n):
    expect = 0 if n < 400 else 1 + (n - 400) // 160
    assert oracle_lib.num_frames(n) == expect
    w = np.zeros(n, dtype=np.int16) + 3
    assert oracle_lib.mfcc(w).shape[0] == expect


def test_nnet_matches_numpy_fp64_and_collapse(model_root, oracle_lib):
    """Oracle TDNN-F forward == independent fp64 numpy forward; collapsed tdnn1 == idct+batchnorm0+delta+affine."""
    import vbmodel
    m = vbmodel.load_model_dir(model_root("tiny"))
    arch = vbmodel.ARCHS["tiny"]
    w = vbmodel.synth_audio(1.9, 5)
    r = oracle_lib.recognize(m, w, frames_per_chunk=51, stages=True)
    T = {k: np.asarray(v) for k, v in m["nnet"].items()}
    ll = vbmodel.nnet_forward_numpy(T, arch, r["mfcc"], r["ivectors"], r["iv_index"])
    assert np.abs(ll - r["loglikes"]).max() < 5e-5
    # un-collapsed front end on one frame
    F = 40
    x = r["mfcc"].astype(np.float64)
    y = x @ T["raw.idct"].astype(np.float64).T * T["raw.bn0_scale"] + T["raw.bn0_offset"]
    t = 10
    feat = np.concatenate([y[t], y[t + 1] - y[t - 1], y[t - 2] - 2 * y[t] + y[t + 2], r["ivectors"][r["iv_index"][t + vbmodel.context_of(arch)[0] - 2]]])
    raw = T["raw.tdnn1.w"].astype(np.float64) @ feat + T["raw.tdnn1.b"]
    spl = np.concatenate([x[t - 2], x[t - 1], x[t], x[t + 1], x[t + 2], r["ivectors"][r["iv_index"][t + vbmodel.context_of(arch)[0] - 2]]])
    col = T["tdnn1.w"].astype(np.float64) @ spl + T["tdnn1.b"]
    assert np.abs(raw - col).max() < 1e-3 * max(1.0, np.abs(raw).max())


def test_ivector_solve_matches_scipy(model_root, oracle_lib):
    """i-vector of the oracle == scipy solve of independently accumulated statistics (fp64 numpy)."""
    import scipy.linalg
    import vbmodel
    m = vbmodel.load_model_dir(model_root("tiny"))
    iv = m["ivector"]
    w = vbmodel.synth_audio(1.2, 9)
    feats = oracle_lib.mfcc(w).astype(np.float64)
    T = len(feats)
    got = oracle_lib.ivectors(m, feats, np.array([T], dtype=np.int32), np.array([T], dtype=np.int32))[0]
    F, D = 40, int(m["cfg"]["ivector-dim"])
    gs = iv["cmvn"]
    norm = np.zeros_like(feats)
    for t in range(T):
        n = t + 1
        fg = min(600 - n, 200)
        norm[t] = feats[t] - (feats[:n].sum(0) + fg / gs[0, F] * gs[0, :F]) / (n + fg)
    lda = iv["lda"].astype(np.float64)
    xu = vbmodel.splice_frames(feats) @ lda[:, :-1].T + lda[:, -1]
    xn = vbmodel.splice_frames(norm) @ lda[:, :-1].T + lda[:, -1]
    du = iv["dubm"]
    ll = du["gconsts"][None] + xn @ du["means_invvars"].astype(np.float64).T - 0.5 * (xn ** 2) @ du["inv_vars"].astype(np.float64).T
    M, Si = iv["ie"]["M"].astype(np.float64), iv["ie"]["sigma_inv"].astype(np.float64)
    po = float(iv["ie"]["prior_offset"][0])
    lin = np.zeros(D); lin[0] = po
    quad = np.eye(D)
    nf = 0.0
    for t in range(T):
        top = np.argsort(-ll[t], kind="stable")[:5]
        p = np.exp(ll[t][top] - ll[t][top[0]]); p /= p.sum()
        p[1:][p[1:] < 0.025] = 0; p /= p.sum()
        for g, wgt in zip(top, p * 0.1):
            if wgt == 0:
                continue
            SiM = Si[g] @ M[g]
            lin += wgt * SiM.T @ xu[t]
            quad += wgt * M[g].T @ SiM
            nf += wgt
    ref = scipy.linalg.solve(quad, lin, assume_a="pos")
    ref[0] -= po
    assert nf < 100  # max_count not reached in this short utterance
    assert np.abs(got - ref).max() < 1e-4


def _brute_force_best(g, ll):
    """Dense Viterbi over ALL states with Bellman-Ford epsilon closure, no pruning."""
    S = g["num_states"]
    INF = np.float32(np.inf)
    eps = np.nonzero(g["arc_pdf"] < 0)[0]
    em = np.nonzero(g["arc_pdf"] >= 0)[0]

    def closure(c):
        for _ in range(16):
            cand = (c[g["arc_src"][eps]] + g["arc_w"][eps]).astype(np.float32)
            new = c.copy()
            np.minimum.at(new, g["arc_next"][eps], cand)
            if np.array_equal(new, c):
                break
            c = new
        return c
    c = np.full(S, INF, dtype=np.float32)
    c[g["start"]] = 0
    c = closure(c)
    for f in range(len(ll)):
        off = np.float32(-c.min())
        ac = (off - ll[f][g["arc_pdf"][em]]).astype(np.float32)
        cand = ((c[g["arc_src"][em]] + ac).astype(np.float32) + g["arc_w"][em]).astype(np.float32)
        n = np.full(S, INF, dtype=np.float32)
        np.minimum.at(n, g["arc_next"][em], cand)
        c = closure(n)
    return float((c + g["final"]).min())


def test_decoder_best_cost_matches_brute_force_viterbi(model_root, oracle_lib):
    import vbmodel
    m = vbmodel.load_model_dir(model_root("tiny"))
    for seed in (5, 6):
        r = oracle_lib.recognize(m, vbmodel.synth_audio(1.4, seed), stages=True, beam=1e9, max_active=10 ** 9, min_active=0)
        d = r["decode"]
        assert d["reached_final"]
        assert abs(_brute_force_best(m["graph"], r["loglikes"]) - d["best_cost"]) < 1e-3


def test_decoder_pruned_path_is_a_valid_path(model_root, oracle_lib):
    """With the production beam the best path must be a connected path of the graph whose recomputed cost is its cost."""
    import vbmodel
    m = vbmodel.load_model_dir(model_root("tiny"))
    g = m["graph"]
    r = oracle_lib.recognize(m, vbmodel.synth_audio(2.0, 8), stages=True)
    d = r["decode"]
    arcs = d["best_arcs"]
    assert g["arc_src"][arcs[0]] == g["start"]
    assert np.all(g["arc_next"][arcs[:-1]] == g["arc_src"][arcs[1:]])
    assert (g["arc_pdf"][arcs] >= 0).sum() == d["frames"]
    # token log invariants: offsets monotone, prev points into the previous or the same frame
    off = d["offsets"]
    assert np.all(np.diff(off) >= 0) and off[-1] == len(d["state"])
    frame_of = np.repeat(np.arange(len(off) - 1), np.diff(off))
    has_prev = d["prev"] >= 0
    eps = g["arc_pdf"][np.maximum(d["arc"], 0)] < 0
    assert np.all(frame_of[d["prev"][has_prev & ~eps]] == frame_of[has_prev & ~eps] - 1)
    assert np.all(frame_of[d["prev"][has_prev & eps]] == frame_of[has_prev & eps])


def test_lattice_pruning_matches_forward_backward(model_root, oracle_lib):
    """Independent pin of FinalizeDecoding's pruning: a link survives iff alpha(src) + link + beta(dst) is within
    lattice_beam of the best complete path, where alpha = the logged forward costs and beta = a dense backward
    Viterbi over the UNPRUNED link log (fp64)."""
    import vbmodel
    m = vbmodel.load_model_dir(model_root("tiny"))
    g = m["graph"]
    r = oracle_lib.recognize(m, vbmodel.synth_audio(1.6, 21), stages=True)
    lb = 2.5
    full = oracle_lib.decode(m, r["loglikes"], lattice_beam=1e30)["lattice"]
    pruned = oracle_lib.decode(m, r["loglikes"], lattice_beam=lb)
    d, lat = pruned, pruned["lattice"]
    assert len(lat["src"]) < len(full["src"]) and len(lat["src"]) > d["frames"]
    # backward costs over the unpruned lattice (states of `full` are in topological order except epsilon links inside a
    # frame, so relax to the fixed point)
    # token costs carry the accumulated per-frame cost offsets, lattice arcs do not (GetRawLattice takes them out)
    cum = np.concatenate([[0.0], np.cumsum(d["cost_offset"].astype(np.float64))])
    cost = d["cost"][full["tok_index"]].astype(np.float64) - cum[full["frame"]]
    w = (g["arc_w"][full["arc"]].astype(np.float64) + full["ac"].astype(np.float64))
    beta = np.full(len(cost), np.inf)
    beta[full["final_state"]] = full["final_cost"]
    for _ in range(10000):
        new = beta.copy()
        np.minimum.at(new, full["src"], w + beta[full["dst"]])
        if np.array_equal(new, beta):
            break
        beta = new
    best = (cost + beta).min()
    assert abs(best - (d["best_cost"] - cum[-1])) < 1e-3
    link_total = cost[full["src"]] + w + beta[full["dst"]] - best
    key_full = list(zip(full["tok_index"][full["src"]], full["arc"]))
    want = {k for k, t in zip(key_full, link_total) if t <= lb - 1e-3}
    maybe = {k for k, t in zip(key_full, link_total) if t <= lb + 1e-3}
    got = set(zip(lat["tok_index"][lat["src"]], lat["arc"]))
    assert want <= got <= maybe
    # every surviving state lies on a surviving path from the start to a final state
    reach = np.zeros(len(lat["tok_index"]), bool)
    reach[lat["final_state"]] = True
    for _ in range(10000):
        n = reach.copy()
        n[lat["src"][reach[lat["dst"]]]] = True
        if np.array_equal(n, reach):
            break
        reach = n
    assert reach.all()
    # the best path is inside the lattice
    assert set(int(a) for a in d["best_arcs"]) <= set(int(x) for x in lat["arc"])


def test_result_text_matches_reference_json_h(model_root, oracle_lib):
    """Oracle result text == the reference's own json.h output (golden fixture from oracle/_ref) for the same words/times."""
    import vbmodel
    m = vbmodel.load_model_dir(model_root("tiny"))
    gold = json.load(open(os.path.join(HERE, "golden", "json_ref.json")))
    # layout check on the golden itself (sorted keys, %f floats, inline arrays)
    assert gold[0]["dump"] == '{\n  "text" : ""\n}'
    assert gold[1]["dump"] == '{\n  "result" : [{\n      "conf" : 1.000000,\n      "end" : 1.110000,\n      "start" : 0.840000,\n      "word" : "one"\n    }],\n  "text" : "one"\n}'
    # oracle output for a decoded utterance re-assembled through the same structure
    r = oracle_lib.recognize(m, vbmodel.synth_audio(1.6, 3), stages=True, lattice=False)
    obj = json.loads(r["text"])
    words, b, e = oracle_lib.align_words(m, r["decode"]["best_arcs"])
    assert [m["words"][int(x)] for x in words] == [x["word"] for x in obj.get("result", [])]
    assert obj["text"] == " ".join(m["words"][int(x)] for x in words)
    expect = "{\n"
    if len(words):
        expect += '  "result" : [' + ", ".join(
            '{\n      "conf" : 1.000000,\n      "end" : %f,\n      "start" : %f,\n      "word" : "%s"\n    }' % (float(e[i]) * 0.03, float(b[i]) * 0.03, m["words"][int(words[i])])
            for i in range(len(words))) + "],\n"
    expect += '  "text" : "%s"\n}' % obj["text"]
    assert r["text"] == expect
    assert np.all(b[1:] >= e[:-1]) and np.all(e > b)


def test_result_text_live_reference_json_h(model_root, oracle_lib):
    """When oracle/_ref (built from /root/reference/src/json.h) is present, compare byte for byte."""
    import ctypes
    import vbmodel
    so = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libref_json.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference absent)")
    lib = ctypes.CDLL(so)
    lib.ref_json_result.restype = ctypes.c_void_p
    m = vbmodel.load_model_dir(model_root("tiny"))
    for seed in (3, 4):
        r = oracle_lib.recognize(m, vbmodel.synth_audio(1.6, seed), stages=True)
        # lattice mode: the MBR one-best (raw floats) written by the reference's own json.h, with PushLattice's arithmetic
        # round(t) * 0.03 + offset [REF src/batch_recognizer.cc:91-93], must give the oracle's lattice-mode text
        for offset in (0.0, 20.13):
            raw = [x.split() for x in oracle_lib.lattice_result(m, r["decode"], 6.0, stage=5).splitlines()]
            n = len(raw)
            ws = [m["words"][int(x[0])].encode() for x in raw]
            arr = (ctypes.c_char_p * max(n, 1))(*ws)
            dbl = lambda v: (ctypes.c_double * max(n, 1))(*v)
            off = float(np.float32(offset))
            rnd = lambda t: float(np.round(np.float32(t)))  # times are BaseFloat; halves cannot occur away from .5 ties of np.round vs C round
            p = lib.ref_json_result(n, arr, dbl([rnd(x[1]) * 0.03 + off for x in raw]), dbl([rnd(x[2]) * 0.03 + off for x in raw]),
                                    dbl([float(np.float32(x[3])) for x in raw]), b" ".join(ws))
            assert oracle_lib.lattice_result(m, r["decode"], 6.0, offset=offset) == ctypes.string_at(p).decode()
            assert n >= 2
        r = oracle_lib.recognize(m, vbmodel.synth_audio(1.6, seed), stages=True, lattice=False)
        words, b, e = oracle_lib.align_words(m, r["decode"]["best_arcs"])
        n = len(words)
        ws = [m["words"][int(x)].encode() for x in words]
        arr = (ctypes.c_char_p * max(n, 1))(*ws)
        dbl = lambda v: (ctypes.c_double * max(n, 1))(*v)
        p = lib.ref_json_result(n, arr, dbl([float(np.float32(x)) * 0.03 for x in b]), dbl([float(np.float32(x)) * 0.03 for x in e]),
                                dbl([1.0] * n), b" ".join(ws))
        assert r["text"] == ctypes.string_at(p).decode()


def test_empty_and_tiny_inputs(model_root, oracle_lib):
    import vbmodel
    m = vbmodel.load_model_dir(model_root("tiny"))
    for n in (0, 100, 399):
        assert oracle_lib.recognize(m, np.zeros(n, dtype=np.int16)) == '{\n  "text" : ""\n}'
    out = json.loads(oracle_lib.recognize(m, vbmodel.synth_audio(0.05, 1)[:480]))
    assert "text" in out
