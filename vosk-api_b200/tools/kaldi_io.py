"""Kaldi on-disk formats for the files of a Vosk model directory (test/bench tooling).

Writes the files `BatchModel::BatchModel()` reads [REF src/batch_model.cc:28-37,53-67,76-77] in the formats
the reference's loaders expect — `ReadKaldiObject(final.mdl)` = TransitionModel + nnet3 AmNnetSimple
[REF src/batch_model.cc:39-45], and the i-vector extractor files named by ivector.conf
[REF src/model.cc:251-256] — so that the engine's real-format loader (csrc/vb_kaldi.cc) can be exercised
offline.  Kaldi itself is absent from /root/reference and from this image, so the layouts
below restate Kaldi's published binary I/O conventions:

  * a binary file starts with "\\0B"; tokens are ASCII followed by one space;
  * basic types: one size byte (4 or 8) then the little-endian value; bool = 'T' / 'F';
  * Matrix: "FM " / "DM " + int32 rows + int32 cols + row-major data; Vector: "FV " / "DV " + int32 dim + data;
  * SpMatrix: "FP " / "DP " + int32 rows + packed lower triangle; integer vector: size byte, int32 count, raw data;
  * nnet3: "<Nnet3> \\n" config lines, blank line, "<NumComponents>", then "<ComponentName> name <Type> ... </Type>".

The network is written UN-collapsed, as `xconfig_to_configs.py` lays out the in-tree recipe
[REF training/local/chain/run_tdnn.sh:98-129] (idct, batchnorm0, spec-augment, delta descriptors, relu/batchnorm/
dropout components, TdnnComponent pairs with a Sum(Scale(0.75, .), .) bypass, and the xent branch): the loader has
to do what `CollapseModel` [REF src/batch_model.cc:46-48] does.
"""
from __future__ import annotations

import os
import shutil
import struct

import numpy as np

import vbmodel


class _W:
    def __init__(self, f):
        self.f = f

    def raw(self, b):
        self.f.write(b)

    def tok(self, s):
        self.f.write(s.encode() + b" ")

    def i32(self, v):
        self.f.write(b"\x04" + struct.pack("<i", int(v)))

    def f32(self, v):
        self.f.write(b"\x04" + struct.pack("<f", float(v)))

    def f64(self, v):
        self.f.write(b"\x08" + struct.pack("<d", float(v)))

    def boolean(self, v):
        self.f.write(b"T" if v else b"F")

    def intvec(self, v):
        v = np.asarray(v, dtype="<i4")
        self.f.write(b"\x04" + struct.pack("<i", len(v)) + v.tobytes())

    def mat(self, m, double=False):
        m = np.ascontiguousarray(np.asarray(m, dtype="<f8" if double else "<f4"))
        if m.ndim != 2:
            m = m.reshape(0, 0)
        self.f.write(b"DM " if double else b"FM ")
        self.i32(m.shape[0])
        self.i32(m.shape[1])
        self.f.write(m.tobytes())

    def vec(self, v, double=False):
        v = np.ascontiguousarray(np.asarray(v, dtype="<f8" if double else "<f4")).reshape(-1)
        self.f.write(b"DV " if double else b"FV ")
        self.i32(len(v))
        self.f.write(v.tobytes())

    def sp(self, m, double=True):
        m = np.asarray(m, dtype=np.float64)
        n = m.shape[0]
        packed = np.concatenate([m[i, :i + 1] for i in range(n)]) if n else np.zeros(0)
        self.f.write(b"DP " if double else b"FP ")
        self.i32(n)
        self.f.write(packed.astype("<f8" if double else "<f4").tobytes())


# --------------------------------------------------------------------------- #
# i-vector extractor files
# --------------------------------------------------------------------------- #
def write_matrix_file(path, m, double=False, binary=True):
    if binary:
        with open(path, "wb") as f:
            w = _W(f)
            w.raw(b"\0B")
            w.mat(m, double)
    else:  # Kaldi text matrix: " [\n  a b c\n  d e f ]\n"
        m = np.asarray(m)
        with open(path, "w") as f:
            f.write(" [\n")
            for i, row in enumerate(m):
                f.write("  " + " ".join(repr(float(x)) for x in row) + (" ]\n" if i == len(m) - 1 else "\n"))


def write_dubm(path, dubm):
    with open(path, "wb") as f:
        w = _W(f)
        w.raw(b"\0B")
        w.tok("<DiagGMM>")
        w.tok("<GCONSTS>"); w.vec(dubm["gconsts"])
        w.tok("<WEIGHTS>"); w.vec(dubm["weights"])
        w.tok("<MEANS_INVVARS>"); w.mat(dubm["means_invvars"])
        w.tok("<INV_VARS>"); w.mat(dubm["inv_vars"])
        w.tok("</DiagGMM>")


def write_ie(path, ie):
    M, S = np.asarray(ie["M"]), np.asarray(ie["sigma_inv"])
    with open(path, "wb") as f:
        w = _W(f)
        w.raw(b"\0B")
        w.tok("<IvectorExtractor>")
        w.tok("<w>"); w.mat(np.zeros((0, 0)), double=True)          # no weight projection (the usual configuration)
        w.tok("<w_vec>"); w.vec(ie["w"], double=True)
        w.tok("<M>"); w.i32(len(M))
        for g in range(len(M)):
            w.mat(M[g], double=True)
        w.tok("<SigmaInv>")
        for g in range(len(S)):
            w.sp(S[g], double=True)
        w.tok("<IvectorOffset>"); w.f64(float(np.asarray(ie["prior_offset"]).reshape(-1)[0]))
        w.tok("</IvectorExtractor>")


# --------------------------------------------------------------------------- #
# final.mdl = TransitionModel + AmNnetSimple
# --------------------------------------------------------------------------- #
def _write_transition_model(w: _W, tid2pdf, tid2phone):
    """Chain topology: one emitting state (forward pdf-class 0, self-loop pdf-class 1) + a final state; one tuple
    (phone, 0, forward pdf, self-loop pdf) per transition-state, which owns (self-loop tid, forward tid) in that order."""
    n_ts = (len(tid2pdf) - 1) // 2
    phones = sorted(set(int(p) for p in tid2phone[1:]))
    w.tok("<TransitionModel>")
    w.tok("<Topology>")
    w.intvec(phones)
    p2i = np.full(max(phones) + 1, -1, dtype=np.int32)
    p2i[phones] = 0
    w.intvec(p2i)
    w.i32(-1)                 # extended format marker: states carry a self-loop pdf class
    w.i32(1)                  # one topology entry shared by all phones
    w.i32(2)                  # two states
    w.i32(0); w.i32(1); w.i32(2); w.i32(0); w.f32(0.5); w.i32(1); w.f32(0.5)
    w.i32(-1); w.i32(-1); w.i32(0)
    w.tok("</Topology>")
    w.tok("<Tuples>")
    w.i32(n_ts)
    for ts in range(n_ts):
        w.i32(int(tid2phone[2 * ts + 1])); w.i32(0); w.i32(int(tid2pdf[2 * ts + 2])); w.i32(int(tid2pdf[2 * ts + 1]))
    w.tok("</Tuples>")
    w.tok("<LogProbs>")
    w.vec(np.concatenate([[0.0], np.full(2 * n_ts, np.log(0.5))]))
    w.tok("</LogProbs>")
    w.tok("</TransitionModel>")


def _updatable_head(w: _W, typ, l2=0.008):
    w.tok(f"<{typ}>")
    w.tok("<MaxChange>"); w.f32(0.75)
    w.tok("<L2Regularize>"); w.f32(l2)
    w.tok("<LearningRate>"); w.f32(0.001)


def _c_affine(w, W, b):
    _updatable_head(w, "NaturalGradientAffineComponent")
    w.tok("<LinearParams>"); w.mat(W)
    w.tok("<BiasParams>"); w.vec(b)
    w.tok("<RankIn>"); w.i32(20)
    w.tok("<RankOut>"); w.i32(80)
    w.tok("<UpdatePeriod>"); w.i32(4)
    w.tok("<NumSamplesHistory>"); w.f32(2000.0)
    w.tok("<Alpha>"); w.f32(4.0)
    w.tok("</NaturalGradientAffineComponent>")


def _c_linear(w, W):
    _updatable_head(w, "LinearComponent")
    w.tok("<Params>"); w.mat(W)
    w.tok("<OrthonormalConstraint>"); w.f32(-1.0)
    w.tok("<UseNaturalGradient>"); w.boolean(True)
    w.tok("<RankInOut>"); w.i32(20); w.i32(80)
    w.tok("<Alpha>"); w.f32(4.0)
    w.tok("<NumSamplesHistory>"); w.f32(2000.0)
    w.tok("<UpdatePeriod>"); w.i32(4)
    w.tok("</LinearComponent>")


def _c_tdnn(w, W, b, offsets):
    _updatable_head(w, "TdnnComponent")
    w.tok("<TimeOffsets>"); w.intvec(offsets)
    w.tok("<LinearParams>"); w.mat(W)
    w.tok("<BiasParams>"); w.vec(b if b is not None else np.zeros(0))
    w.tok("<OrthonormalConstraint>"); w.f32(-1.0 if b is None else 0.0)
    w.tok("<UseNaturalGradient>"); w.boolean(True)
    w.tok("<NumSamplesHistory>"); w.f32(2000.0)
    w.tok("<AlphaInOut>"); w.f32(4.0); w.f32(4.0)
    w.tok("<RankInOut>"); w.i32(20); w.i32(80)
    w.tok("</TdnnComponent>")


def _c_fixed_affine(w, W, b):
    w.tok("<FixedAffineComponent>")
    w.tok("<LinearParams>"); w.mat(W)
    w.tok("<BiasParams>"); w.vec(b)
    w.tok("</FixedAffineComponent>")


BN_EPS, BN_TARGET_RMS = 1e-3, 1.0


def _c_batchnorm(w, scale, offset):
    """Test-mode batchnorm y = (x - mean) * target_rms / sqrt(var + eps): stores mean / var that reproduce scale / offset."""
    scale, offset = np.asarray(scale, np.float64), np.asarray(offset, np.float64)
    var = (BN_TARGET_RMS / scale) ** 2 - BN_EPS
    mean = -offset / scale
    assert (var > -BN_EPS).all()
    w.tok("<BatchNormComponent>")
    w.tok("<Dim>"); w.i32(len(scale))
    w.tok("<BlockDim>"); w.i32(len(scale))
    w.tok("<Epsilon>"); w.f32(BN_EPS)
    w.tok("<TargetRms>"); w.f32(BN_TARGET_RMS)
    w.tok("<TestMode>"); w.boolean(True)
    w.tok("<Count>"); w.f64(12345.0)
    w.tok("<StatsMean>"); w.vec(mean, double=True)
    w.tok("<StatsVar>"); w.vec(var, double=True)
    w.tok("</BatchNormComponent>")


def _c_relu(w, dim):
    w.tok("<RectifiedLinearComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<ValueAvg>"); w.vec(np.full(dim, 0.3), double=True)
    w.tok("<DerivAvg>"); w.vec(np.full(dim, 0.5), double=True)
    w.tok("<Count>"); w.f64(1000.0)
    w.tok("<OderivRms>"); w.vec(np.full(dim, 0.01), double=True)
    w.tok("<OderivCount>"); w.f64(1000.0)
    w.tok("<NumDimsSelfRepaired>"); w.f64(0.0)
    w.tok("<NumDimsProcessed>"); w.f64(0.0)
    w.tok("<SelfRepairScale>"); w.f32(1e-5)
    w.tok("</RectifiedLinearComponent>")


def _c_logsoftmax(w, dim):
    w.tok("<LogSoftmaxComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<ValueAvg>"); w.vec(np.zeros(0), double=True)
    w.tok("<DerivAvg>"); w.vec(np.zeros(0), double=True)
    w.tok("<Count>"); w.f64(0.0)
    w.tok("<NumDimsSelfRepaired>"); w.f64(0.0)
    w.tok("<NumDimsProcessed>"); w.f64(0.0)
    w.tok("</LogSoftmaxComponent>")


def _c_noop(w, dim):
    w.tok("<NoOpComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<BackpropScale>"); w.f32(1.0)
    w.tok("</NoOpComponent>")


def _c_general_dropout(w, dim):
    w.tok("<GeneralDropoutComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<BlockDim>"); w.i32(dim)
    w.tok("<TimePeriod>"); w.i32(0)
    w.tok("<DropoutProportion>"); w.f32(0.0)
    w.tok("<TestMode>"); w.boolean(True)
    w.tok("<Continuous>"); w.boolean(True)
    w.tok("</GeneralDropoutComponent>")


def _c_spec_time_mask(w, dim):
    w.tok("<SpecAugmentTimeMaskComponent>")
    w.tok("<Dim>"); w.i32(dim)
    w.tok("<ZeroedProportion>"); w.f32(0.2)
    w.tok("<TimeMaskMaxFrames>"); w.i32(20)
    w.tok("</SpecAugmentTimeMaskComponent>")


def write_final_mdl(path, T, arch, priors=None):
    """T: the tensor dict of vbmodel.make_nnet (collapsed + raw.* pieces)."""
    F, I, H = arch["feat_dim"], arch["ivector_dim"], arch["hidden"]
    PS, PB, NP = arch["prefinal_small"], arch["prefinal_big"], arch["num_pdfs"]
    bys = vbmodel.BYPASS_SCALE
    cfg = [f"input-node name=ivector dim={I}", f"input-node name=input dim={F}"]
    comps = []  # (name, writer)

    def node(name, inp, comp=None):
        cfg.append(f"component-node name={name} component={comp or name} input={inp}")

    comps.append(("idct", lambda w: _c_fixed_affine(w, T["raw.idct"], np.zeros(F))))
    node("idct", "input")
    comps.append(("batchnorm0", lambda w: _c_batchnorm(w, T["raw.bn0_scale"], T["raw.bn0_offset"])))
    node("batchnorm0", "idct")
    comps.append(("spec-augment.freq-mask", lambda w: _c_general_dropout(w, F)))
    node("spec-augment.freq-mask", "batchnorm0")
    comps.append(("spec-augment.time-mask", lambda w: _c_spec_time_mask(w, F)))
    node("spec-augment.time-mask", "spec-augment.freq-mask")
    x = "spec-augment.time-mask"
    comps.append(("delta", lambda w: _c_noop(w, 3 * F)))
    node("delta", f"Append({x}, Sum(Scale(-1.0, Offset({x}, -1)), Offset({x}, 1)), "
                  f"Sum(Sum(Offset({x}, -2), Offset({x}, 2)), Scale(-2.0, {x})))")
    comps.append(("input2", lambda w: _c_noop(w, 3 * F + I)))
    node("input2", "Append(delta, ReplaceIndex(ivector, t, 0))")

    def relu_bn_dropout(prefix, dim, bn, bn_name="batchnorm"):
        comps.append((f"{prefix}.relu", lambda w: _c_relu(w, dim)))
        node(f"{prefix}.relu", f"{prefix}.affine")
        comps.append((f"{prefix}.{bn_name}", lambda w: _c_batchnorm(w, T[bn + ".bn_scale"], T[bn + ".bn_offset"])))
        node(f"{prefix}.{bn_name}", f"{prefix}.relu")

    comps.append(("tdnn1.affine", lambda w: _c_affine(w, T["raw.tdnn1.w"], T["raw.tdnn1.b"])))
    node("tdnn1.affine", "input2")
    relu_bn_dropout("tdnn1", H, "tdnn1")
    comps.append(("tdnn1.dropout", lambda w: _c_general_dropout(w, H)))
    node("tdnn1.dropout", "tdnn1.batchnorm")
    prev = "tdnn1.dropout"
    for k, s in enumerate(arch["strides"], start=2):
        nm = f"tdnnf{k}"
        lo = [-s, 0] if s else [0]
        ao = [0, s] if s else [0]
        comps.append((f"{nm}.linear", lambda w, nm=nm, lo=lo: _c_tdnn(w, T[f"{nm}.linear.w"], None, lo)))
        node(f"{nm}.linear", prev)
        comps.append((f"{nm}.affine", lambda w, nm=nm, ao=ao: _c_tdnn(w, T[f"{nm}.affine.w"], T[f"{nm}.affine.b"], ao)))
        node(f"{nm}.affine", f"{nm}.linear")
        relu_bn_dropout(nm, H, nm)
        comps.append((f"{nm}.dropout", lambda w: _c_general_dropout(w, H)))
        node(f"{nm}.dropout", f"{nm}.batchnorm")
        comps.append((f"{nm}.noop", lambda w: _c_noop(w, H)))
        node(f"{nm}.noop", f"Sum(Scale({bys}, {prev}), {nm}.dropout)")
        prev = f"{nm}.noop"
    comps.append(("prefinal-l", lambda w: _c_linear(w, T["prefinal_l.w"])))
    node("prefinal-l", prev)
    rng = np.random.default_rng(5)
    for br in ("chain", "xent"):
        p = f"prefinal-{br}"
        if br == "chain":
            Wa, ba, Wl = T["prefinal.affine.w"], T["prefinal.affine.b"], T["raw.prefinal.linear.w"]
            s1, o1 = T["prefinal.bn_scale"], T["prefinal.bn_offset"]
            s2, o2 = T["raw.prefinal.bn2_scale"], T["raw.prefinal.bn2_offset"]
            Wo, bo = T["output.w"], T["output.b"]
        else:  # the xent branch is present in a trained final.mdl but not on the path of output-node "output"
            Wa, ba, Wl = rng.standard_normal((PB, PS)) * 0.1, np.zeros(PB), rng.standard_normal((PS, PB)) * 0.1
            s1, o1, s2, o2 = np.ones(PB), np.zeros(PB), np.ones(PS), np.zeros(PS)
            Wo, bo = rng.standard_normal((NP, PS)) * 0.1, np.zeros(NP)
        comps.append((f"{p}.affine", lambda w, Wa=Wa, ba=ba: _c_affine(w, Wa, ba)))
        node(f"{p}.affine", "prefinal-l")
        comps.append((f"{p}.relu", lambda w: _c_relu(w, PB)))
        node(f"{p}.relu", f"{p}.affine")
        comps.append((f"{p}.batchnorm1", lambda w, s1=s1, o1=o1: _c_batchnorm(w, s1, o1)))
        node(f"{p}.batchnorm1", f"{p}.relu")
        comps.append((f"{p}.linear", lambda w, Wl=Wl: _c_linear(w, Wl)))
        node(f"{p}.linear", f"{p}.batchnorm1")
        comps.append((f"{p}.batchnorm2", lambda w, s2=s2, o2=o2: _c_batchnorm(w, s2, o2)))
        node(f"{p}.batchnorm2", f"{p}.linear")
        out = "output" if br == "chain" else "output-xent"
        comps.append((f"{out}.affine", lambda w, Wo=Wo, bo=bo: _c_affine(w, Wo, bo)))
        node(f"{out}.affine", f"{p}.batchnorm2")
        if br == "chain":
            cfg.append("output-node name=output input=output.affine objective=linear")
        else:
            comps.append(("output-xent.log-softmax", lambda w: _c_logsoftmax(w, NP)))
            node("output-xent.log-softmax", "output-xent.affine")
            cfg.append("output-node name=output-xent input=output-xent.log-softmax objective=linear")
    ctx = vbmodel.context_of(arch)[0]
    with open(path, "wb") as f:
        w = _W(f)
        w.raw(b"\0B")
        _write_transition_model(w, T["tid2pdf"], T["tid2phone"])
        w.tok("<Nnet3>")
        w.raw(b"\n")
        for line in cfg:
            w.raw(line.encode() + b"\n")
        w.raw(b"\n")
        w.tok("<NumComponents>"); w.i32(len(comps))
        for name, fn in comps:
            w.tok("<ComponentName>"); w.tok(name)
            fn(w)
        w.tok("</Nnet3>")
        w.tok("<LeftContext>"); w.i32(ctx)
        w.tok("<RightContext>"); w.i32(ctx)
        w.tok("<Priors>"); w.vec(priors if priors is not None else np.zeros(0))


def convert_model_dir(src_model, dst_root, priors=None, text_cmvn=True):
    """Re-writes a vbmodel (VBT container) model directory as <dst_root>/model in Kaldi's formats.
    The graph, word table, word-boundary and conf files already are the real formats and are copied."""
    m = os.path.join(dst_root, "model")
    if os.path.exists(m):
        shutil.rmtree(m)
    shutil.copytree(src_model, m)
    T = vbmodel.read_vbt(os.path.join(src_model, "am/final.mdl"))
    cfg = {}
    for line in vbmodel.vbt_text(T["config"]).splitlines():
        if line.strip():
            k, _, v = line.partition(" ")
            cfg[k] = v
    arch = dict(feat_dim=int(cfg["feat-dim"]), ivector_dim=int(cfg["ivector-dim"]), hidden=int(cfg["hidden-dim"]),
                bottleneck=int(cfg["bottleneck-dim"]), strides=[int(s) for s in cfg["tdnnf-strides"].split()],
                prefinal_small=int(cfg["prefinal-small"]), prefinal_big=int(cfg["prefinal-big"]), num_pdfs=int(cfg["num-pdfs"]))
    write_final_mdl(os.path.join(m, "am/final.mdl"), T, arch, priors)
    iv = os.path.join(src_model, "ivector")
    write_matrix_file(os.path.join(m, "ivector/final.mat"), vbmodel.read_vbt(os.path.join(iv, "final.mat"))["lda"])
    write_dubm(os.path.join(m, "ivector/final.dubm"), vbmodel.read_vbt(os.path.join(iv, "final.dubm")))
    write_ie(os.path.join(m, "ivector/final.ie"), vbmodel.read_vbt(os.path.join(iv, "final.ie")))
    write_matrix_file(os.path.join(m, "ivector/global_cmvn.stats"), vbmodel.read_vbt(os.path.join(iv, "global_cmvn.stats"))["stats"],
                      double=True, binary=not text_cmvn)
    return m


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("src_model")
    ap.add_argument("dst_root")
    a = ap.parse_args()
    print(convert_model_dir(a.src_model, a.dst_root))
