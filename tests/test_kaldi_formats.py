"""Kaldi on-disk formats: the same synthetic model written (a) in the generator's tensor container,
collapsed, and (b) as Kaldi files — final.mdl = TransitionModel + UN-collapsed nnet3 (idct, batchnorm0, spec-augment,
delta descriptors, TdnnComponent pairs with Sum(Scale()) bypass, xent branch), final.mat / final.dubm / final.ie binary,
global_cmvn.stats as a text matrix — must load to the same engine model.  The Kaldi layouts are restated from Kaldi's I/O
conventions (tools/kaldi_io.py, csrc/vb_kaldi.cc): Kaldi is absent here, so this pins reader against writer and, more
importantly, the loader's CollapseModel-equivalent folding against the generator's independent numpy fold."""
import ctypes
import os
import shutil

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    L = ctypes.CDLL(os.path.join(ROOT, "vosk-api_b200", "lib", "libvosk.so"))
    L.vosk_b200_model_tensor.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int64]
    L.vosk_b200_model_tensor.restype = ctypes.c_int64
    L.vosk_b200_last_error.restype = ctypes.c_char_p
    L.vosk_b200_model_check.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    return L


def tensor(lib, mdir, name, optional=False):
    n = lib.vosk_b200_model_tensor(mdir.encode(), name.encode(), None, 0)
    assert n >= 0, lib.vosk_b200_last_error()
    if n == 0:
        assert optional, name
        return None
    out = np.zeros(n, dtype=np.float64)
    assert lib.vosk_b200_model_tensor(mdir.encode(), name.encode(), out.ctypes.data, n) == n
    return out


@pytest.fixture(scope="module")
def pair(model_root, tmp_path_factory):
    import kaldi_io
    src = model_root("tiny")
    dst = kaldi_io.convert_model_dir(src, str(tmp_path_factory.mktemp("kaldi_tiny")))
    return src, dst


def test_kaldi_files_are_kaldi_files(pair):
    _, k = pair
    assert open(os.path.join(k, "am/final.mdl"), "rb").read(20).startswith(b"\0B<TransitionModel> ")
    assert open(os.path.join(k, "ivector/final.dubm"), "rb").read(12).startswith(b"\0B<DiagGMM> ")
    assert open(os.path.join(k, "ivector/final.ie"), "rb").read(21).startswith(b"\0B<IvectorExtractor> ")
    assert open(os.path.join(k, "ivector/final.mat"), "rb").read(5) == b"\0BFM "
    assert open(os.path.join(k, "ivector/global_cmvn.stats")).read(3) == " [\n"
    mdl = open(os.path.join(k, "am/final.mdl"), "rb").read()
    assert b"component-node name=tdnnf2.noop component=tdnnf2.noop input=Sum(Scale(0.75, tdnn1.dropout), tdnnf2.dropout)\n" in mdl
    assert b"output-node name=output-xent" in mdl and b"<TdnnComponent> " in mdl


def test_compiled_network_equals_the_collapsed_container(lib, pair):
    import vbmodel
    v, k = pair
    mv, mk = tensor(lib, v, "meta"), tensor(lib, k, "meta")
    np.testing.assert_array_equal(mv, mk)   # ops, context, pdfs, ivector dim, bypass scale, prior offset, gaussians
    assert int(mk[1]) == vbmodel.context_of(vbmodel.ARCHS["tiny"])[0]
    for i in range(int(mv[0])):
        np.testing.assert_array_equal(tensor(lib, v, f"op{i}.meta"), tensor(lib, k, f"op{i}.meta"), err_msg=f"op{i}")
        for part in ("w", "b", "bn_scale", "bn_offset"):
            a, b = tensor(lib, v, f"op{i}.{part}", True), tensor(lib, k, f"op{i}.{part}", True)
            if a is None or b is None:
                # a bias the container stores as all-zero may be dropped by the compiler and vice versa
                assert (a is None or not a.any()) and (b is None or not b.any()), (i, part)
                continue
            scale = max(1e-6, float(np.abs(a).max()))
            np.testing.assert_allclose(b, a, rtol=0, atol=3e-6 * scale, err_msg=f"op{i}.{part}")
    for name in ("tid2pdf", "tid2phone"):
        np.testing.assert_array_equal(tensor(lib, v, name), tensor(lib, k, name))


def test_ivector_extractor_files_load_identically(lib, pair):
    v, k = pair
    for name in ("iv.lda", "iv.weights", "iv.means_invvars", "iv.inv_vars", "iv.M", "iv.sigma_inv", "iv.cmvn"):
        np.testing.assert_array_equal(tensor(lib, v, name), tensor(lib, k, name), err_msg=name)
    # gconsts are stored by the writer and must also agree with the formula used when a file omits them
    np.testing.assert_allclose(tensor(lib, k, "iv.gconsts"), tensor(lib, v, "iv.gconsts"), rtol=0, atol=1e-4)


def test_dubm_without_gconsts_and_binary_cmvn(lib, pair, tmp_path):
    import kaldi_io
    import vbmodel
    v, k = pair
    dst = str(tmp_path / "model")
    shutil.copytree(k, dst)
    d = vbmodel.read_vbt(os.path.join(v, "ivector/final.dubm"))
    with open(os.path.join(dst, "ivector/final.dubm"), "wb") as f:
        w = kaldi_io._W(f)
        w.raw(b"\0B")
        w.tok("<DiagGMM>")
        w.tok("<WEIGHTS>"); w.vec(d["weights"])
        w.tok("<MEANS_INVVARS>"); w.mat(d["means_invvars"])
        w.tok("<INV_VARS>"); w.mat(d["inv_vars"])
        w.tok("</DiagGMM>")
    kaldi_io.write_matrix_file(os.path.join(dst, "ivector/global_cmvn.stats"),
                               vbmodel.read_vbt(os.path.join(v, "ivector/global_cmvn.stats"))["stats"], double=True, binary=True)
    np.testing.assert_allclose(tensor(lib, dst, "iv.gconsts"), tensor(lib, v, "iv.gconsts"), rtol=0, atol=2e-3)
    np.testing.assert_array_equal(tensor(lib, dst, "iv.cmvn"), tensor(lib, v, "iv.cmvn"))


def test_priors_are_folded_into_the_output_bias(lib, pair, tmp_path):
    import kaldi_io
    import vbmodel
    v, _ = pair
    npdf = vbmodel.ARCHS["tiny"]["num_pdfs"]
    pri = np.random.default_rng(3).dirichlet(np.full(npdf, 2.0))
    dst = kaldi_io.convert_model_dir(v, str(tmp_path), priors=pri)
    last = int(tensor(lib, v, "meta")[0]) - 1
    np.testing.assert_allclose(tensor(lib, dst, f"op{last}.b"), tensor(lib, v, f"op{last}.b") - np.log(pri), rtol=0, atol=1e-5)


def test_loader_reports_what_it_cannot_express(lib, pair, tmp_path):
    _, k = pair
    buf = ctypes.create_string_buffer(2048)

    def broken(edit):
        dst = str(tmp_path / "m")
        if os.path.exists(dst):
            shutil.rmtree(dst)
        shutil.copytree(k, dst)
        p = os.path.join(dst, "am/final.mdl")
        data = edit(open(p, "rb").read())
        open(p, "wb").write(data)
        assert lib.vosk_b200_model_check(dst.encode(), buf, 2048) == -1
        return buf.value.decode()

    assert "truncated" in broken(lambda b: b[: len(b) // 2])
    # a nonlinearity the engine has no epilogue for
    msg = broken(lambda b: b.replace(b"RectifiedLinearComponent>", b"SigmoidComponent>"))
    assert "unsupported nnet3 component type SigmoidComponent" in msg
    # a descriptor function outside the supported set (the config section is plain text lines)
    msg = broken(lambda b: b.replace(b"ReplaceIndex(ivector, t, 0)", b"Const(1.0, 16)"))
    assert "unsupported descriptor function Const" in msg
    # neither format
    assert "unrecognised" in broken(lambda b: b"XXXX" + b[4:])


def _write_custom_mdl(path, cfg_lines, comps, n_pdfs, hmm_triples=False):
    """A final.mdl with an arbitrary nnet3 graph; hmm_triples: plain-HMM topology (no self-loop pdf class) and <Triples>."""
    import kaldi_io
    with open(path, "wb") as f:
        w = kaldi_io._W(f)
        w.raw(b"\0B")
        n_ts = n_pdfs // 2
        tid2pdf = np.zeros(2 * n_ts + 1, dtype=np.int32)
        tid2phone = np.zeros(2 * n_ts + 1, dtype=np.int32)
        for ts in range(n_ts):
            tid2pdf[2 * ts + 1], tid2pdf[2 * ts + 2] = 2 * ts + 1, 2 * ts
            tid2phone[2 * ts + 1] = tid2phone[2 * ts + 2] = 1 + ts % 7
        if not hmm_triples:
            kaldi_io._write_transition_model(w, tid2pdf, tid2phone)
        else:
            w.tok("<TransitionModel>")
            w.tok("<Topology>")
            w.intvec(list(range(1, 8)))
            w.intvec([-1] + [0] * 7)
            w.i32(1); w.i32(2)
            w.i32(0); w.i32(2); w.i32(0); w.f32(0.5); w.i32(1); w.f32(0.5)   # state 0: pdf class 0, self loop + forward
            w.i32(-1); w.i32(0)
            w.tok("</Topology>")
            w.tok("<Triples>")
            w.i32(n_ts)
            for ts in range(n_ts):
                w.i32(1 + ts % 7); w.i32(0); w.i32(ts)
            w.tok("</Triples>")
            w.tok("<LogProbs>"); w.vec(np.zeros(2 * n_ts + 1)); w.tok("</LogProbs>")
            w.tok("</TransitionModel>")
        w.tok("<Nnet3>")
        w.raw(b"\n")
        for line in cfg_lines:
            w.raw(line.encode() + b"\n")
        w.raw(b"\n")
        w.tok("<NumComponents>"); w.i32(len(comps))
        for name, fn in comps:
            w.tok("<ComponentName>"); w.tok(name)
            fn(w)
        w.tok("</Nnet3>")
        w.tok("<LeftContext>"); w.i32(0)
        w.tok("<RightContext>"); w.i32(0)
        w.tok("<Priors>"); w.vec(np.zeros(0))


def test_compiler_on_an_lda_style_network_against_direct_evaluation(lib, model_root, tmp_path):
    """A graph the TDNN-F recipe does not produce: spliced fixed 'lda' affine over Append(Offset...) + a scaled i-vector, plain
    AffineComponent, ReLU + batchnorm, a dim-range-node, a bypass over a no-op, <Triples> + HMM topology.  The compiled op chain,
    evaluated in numpy, must equal the direct evaluation of the graph as written."""
    import kaldi_io
    rng = np.random.default_rng(11)
    F, I, P = 40, 16, 96
    Wl, bl = rng.standard_normal((64, 3 * F + I)) * 0.1, rng.standard_normal(64) * 0.1
    W1, b1 = rng.standard_normal((48, 64)) * 0.2, rng.standard_normal(48) * 0.1
    s1, o1 = rng.uniform(0.5, 1.5, 48), rng.standard_normal(48) * 0.1
    W2, b2 = rng.standard_normal((48, 2 * 32)) * 0.2, rng.standard_normal(48) * 0.1     # TdnnComponent over the dim-range, offsets -3, 3
    s2, o2 = rng.uniform(0.5, 1.5, 48), rng.standard_normal(48) * 0.1
    Wo, bo = rng.standard_normal((P, 48)) * 0.2, rng.standard_normal(P) * 0.1
    cfg = [f"input-node name=ivector dim={I}", f"input-node name=input dim={F}",
           "component-node name=lda component=lda input=Append(Offset(input, -1), input, Offset(input, 1), Scale(0.5, ReplaceIndex(ivector, t, 0)))",
           "component-node name=l1.affine component=l1.affine input=lda",
           "component-node name=l1.relu component=l1.relu input=l1.affine",
           "component-node name=l1.bn component=l1.bn input=l1.relu",
           "dim-range-node name=l1.part input-node=l1.bn dim-offset=8 dim=32",
           "component-node name=l2.affine component=l2.affine input=l1.part",
           "component-node name=l2.relu component=l2.relu input=l2.affine",
           "component-node name=l2.bn component=l2.bn input=l2.relu",
           "component-node name=l2.noop component=l2.noop input=Sum(l2.bn, Scale(0.66, l1.bn))",
           "component-node name=out.affine component=out.affine input=l2.noop",
           "output-node name=output input=out.affine objective=linear"]

    def affine(W, b):
        def fn(w):
            w.tok("<AffineComponent>")
            w.tok("<LearningRate>"); w.f32(0.001)
            w.tok("<LinearParams>"); w.mat(W)
            w.tok("<BiasParams>"); w.vec(b)
            w.tok("<IsGradient>"); w.boolean(False)
            w.tok("</AffineComponent>")
        return fn
    comps = [("lda", lambda w: kaldi_io._c_fixed_affine(w, Wl, bl)), ("l1.affine", affine(W1, b1)),
             ("l1.relu", lambda w: kaldi_io._c_relu(w, 48)), ("l1.bn", lambda w: kaldi_io._c_batchnorm(w, s1, o1)),
             ("l2.affine", lambda w: kaldi_io._c_tdnn(w, W2, b2, [-3, 3])), ("l2.relu", lambda w: kaldi_io._c_relu(w, 48)),
             ("l2.bn", lambda w: kaldi_io._c_batchnorm(w, s2, o2)), ("l2.noop", lambda w: kaldi_io._c_noop(w, 48)),
             ("out.affine", affine(Wo, bo))]
    dst = str(tmp_path / "model")
    shutil.copytree(model_root("tiny"), dst)
    _write_custom_mdl(os.path.join(dst, "am/final.mdl"), cfg, comps, P, hmm_triples=True)
    meta = tensor(lib, dst, "meta")
    n_ops, ctx = int(meta[0]), int(meta[1])
    assert n_ops == 3 and ctx == 4 and abs(meta[4] - 0.66) < 1e-6
    # <Triples> with a plain HMM topology: both transitions of a state emit the state's single pdf
    t2p = tensor(lib, dst, "tid2pdf")
    np.testing.assert_array_equal(t2p[1:], np.repeat(np.arange(P // 2), 2))
    # evaluate the compiled chain and the graph as written on random input (interior frames only)
    T = 40
    x = rng.standard_normal((T, F)).astype(np.float32).astype(np.float64)
    iv = rng.standard_normal(I).astype(np.float32).astype(np.float64)
    nodes = [x]
    for i in range(n_ops):
        m = tensor(lib, dst, f"op{i}.meta")
        in_node, byp, uses_iv, relu, K, N, noff = (int(v) for v in m[:7])
        offs = [int(v) for v in m[7:7 + noff]]
        W = tensor(lib, dst, f"op{i}.w").reshape(N, K)
        b = tensor(lib, dst, f"op{i}.b", True)
        src = nodes[in_node]
        ar = np.arange(T)
        cols = [src[np.clip(ar + o, 0, T - 1)] for o in offs]
        if uses_iv:
            cols.append(np.tile(iv, (T, 1)))
        z = np.concatenate(cols, 1) @ W.T + (b if b is not None else 0.0)
        if relu:
            z = np.maximum(z, 0.0) * tensor(lib, dst, f"op{i}.bn_scale") + tensor(lib, dst, f"op{i}.bn_offset")
        if byp >= 0:
            z = z + meta[4] * nodes[byp]
        nodes.append(z)
    ar = np.arange(T)
    sp = lambda a, o: a[np.clip(ar + o, 0, T - 1)]
    lda = np.concatenate([sp(x, -1), x, sp(x, 1), np.tile(0.5 * iv, (T, 1))], 1) @ Wl.T + bl
    h1 = np.maximum(lda @ W1.T + b1, 0.0) * s1 + o1
    part = h1[:, 8:40]
    h2 = np.maximum(np.concatenate([sp(part, -3), sp(part, 3)], 1) @ W2.T + b2, 0.0) * s2 + o2
    want = (h2 + 0.66 * h1) @ Wo.T + bo
    np.testing.assert_allclose(nodes[-1][ctx:T - ctx], want[ctx:T - ctx], rtol=0, atol=2e-5)


def test_odd_ivector_and_output_dimensions_are_padded_exactly(lib, model_root):
    """Real models have i-vector dimensions such as 30 and arbitrary pdf counts; the kernels want multiples of 4 / 16.  The
    loader pads with zero extractor columns, zero weight columns and zero output rows (both file formats)."""
    import kaldi_io
    import vbmodel
    import tempfile
    src = model_root("tiny", overrides=dict(ivector_dim=14, num_pdfs=90), tag="_odd")
    T = vbmodel.read_vbt(os.path.join(src, "am/final.mdl"))
    ie = vbmodel.read_vbt(os.path.join(src, "ivector/final.ie"))
    with tempfile.TemporaryDirectory() as td:
        for mdir in (src, kaldi_io.convert_model_dir(src, td)):
            meta = tensor(lib, mdir, "meta")
            assert int(meta[2]) == 90 and int(meta[3]) == 16          # pdfs stay 90, the i-vector is carried as 16
            m0 = tensor(lib, mdir, "op0.meta")
            K, N = int(m0[4]), int(m0[5])
            assert K == 5 * 40 + 16
            w0 = tensor(lib, mdir, "op0.w").reshape(N, K)
            np.testing.assert_allclose(w0[:, :214], T["tdnn1.w"], rtol=0, atol=3e-6 * np.abs(T["tdnn1.w"]).max())
            assert not w0[:, 214:].any()
            last = int(meta[0]) - 1
            ml = tensor(lib, mdir, f"op{last}.meta")
            assert int(ml[5]) == 96
            wl = tensor(lib, mdir, f"op{last}.w").reshape(96, int(ml[4]))
            np.testing.assert_allclose(wl[:90], T["output.w"], rtol=0, atol=3e-6 * np.abs(T["output.w"]).max())
            assert not wl[90:].any() and not tensor(lib, mdir, f"op{last}.b")[90:].any()
            M = tensor(lib, mdir, "iv.M").reshape(ie["M"].shape[0], ie["M"].shape[1], 16)
            np.testing.assert_array_equal(M[:, :, :14], ie["M"].astype(np.float64))
            assert not M[:, :, 14:].any()


def test_reader_on_hand_assembled_bytes(lib, model_root, tmp_path):
    """Byte strings put together by hand from Kaldi's I/O rules (not through tools/kaldi_io.py): binary marker, token + space,
    size-prefixed basic types, FM / FV / DM headers — so reader and writer cannot agree on a private dialect."""
    import struct
    dst = str(tmp_path / "model")
    shutil.copytree(model_root("tiny"), dst)
    F = 40
    lda = (np.arange(F * (7 * F + 1), dtype=np.float32).reshape(F, 7 * F + 1) % 17 - 8) / 16.0
    with open(os.path.join(dst, "ivector/final.mat"), "wb") as f:
        f.write(b"\x00B" + b"FM " + b"\x04" + struct.pack("<i", F) + b"\x04" + struct.pack("<i", 7 * F + 1) + lda.astype("<f4").tobytes())
    G = 16
    w = np.full(G, 1.0 / G, dtype=np.float32)
    iv = (1.0 + (np.arange(G * F, dtype=np.float32).reshape(G, F) % 5) / 4.0)
    miv = ((np.arange(G * F, dtype=np.float32).reshape(G, F) % 7) - 3.0) / 2.0
    with open(os.path.join(dst, "ivector/final.dubm"), "wb") as f:
        f.write(b"\x00B<DiagGMM> ")
        f.write(b"<WEIGHTS> FV \x04" + struct.pack("<i", G) + w.astype("<f4").tobytes())
        f.write(b"<MEANS_INVVARS> FM \x04" + struct.pack("<i", G) + b"\x04" + struct.pack("<i", F) + miv.astype("<f4").tobytes())
        f.write(b"<INV_VARS> FM \x04" + struct.pack("<i", G) + b"\x04" + struct.pack("<i", F) + iv.astype("<f4").tobytes())
        f.write(b"</DiagGMM> ")
    stats = np.arange(2 * (F + 1), dtype=np.float64).reshape(2, F + 1)
    with open(os.path.join(dst, "ivector/global_cmvn.stats"), "wb") as f:
        f.write(b"\x00B" + b"DM " + b"\x04" + struct.pack("<i", 2) + b"\x04" + struct.pack("<i", F + 1) + stats.astype("<f8").tobytes())
    np.testing.assert_array_equal(tensor(lib, dst, "iv.lda").reshape(F, -1), lda.astype(np.float64))
    np.testing.assert_array_equal(tensor(lib, dst, "iv.inv_vars").reshape(G, F), iv.astype(np.float64))
    np.testing.assert_array_equal(tensor(lib, dst, "iv.means_invvars").reshape(G, F), miv.astype(np.float64))
    np.testing.assert_array_equal(tensor(lib, dst, "iv.cmvn").reshape(2, F + 1), stats)
    # gconsts were absent: log w - 0.5 (F log 2pi - sum log inv_var + sum mean_invvar^2 / inv_var)
    want = np.log(w.astype(np.float64)) - 0.5 * (F * np.log(2 * np.pi) - np.log(iv).sum(1) + (miv.astype(np.float64) ** 2 / iv).sum(1))
    np.testing.assert_allclose(tensor(lib, dst, "iv.gconsts"), want, rtol=0, atol=1e-4)


def test_compressed_matrices_expand_as_kaldi_does(lib, model_root, tmp_path):
    """Kaldi's CompressedMatrix layouts (CM: one byte per element, per-column percentile headers, stored by columns; CM2: two bytes;
    CM3: one byte), assembled by hand; the expansion is Kaldi's float arithmetic, restated here in numpy."""
    import struct
    F, S = 40, 7 * 40 + 1
    rng = np.random.default_rng(5)
    mn, rg = np.float32(-3.5), np.float32(9.25)
    inc16, f32 = np.float32(1.52590218966964e-05), np.float32

    def check(payload, want):
        dst = str(tmp_path / ("model_%d" % len(os.listdir(tmp_path))))
        shutil.copytree(model_root("tiny"), dst)
        with open(os.path.join(dst, "ivector/final.mat"), "wb") as f:
            f.write(b"\x00B" + payload)
        np.testing.assert_array_equal(tensor(lib, dst, "iv.lda").reshape(F, S), want.astype(np.float64))

    head = struct.pack("<ffii", float(mn), float(rg), F, S)
    # CM3: one byte per element, row-major
    b3 = rng.integers(0, 256, size=(F, S), dtype=np.uint8)
    check(b"CM3 " + head + b3.tobytes(), mn + rg * f32(1.0 / 255.0) * b3.astype(np.float32))
    # CM2: two bytes per element, row-major
    b2 = rng.integers(0, 65536, size=(F, S)).astype("<u2")
    check(b"CM2 " + head + b2.tobytes(), mn + rg * inc16 * b2.astype(np.float32))
    # CM: per-column percentiles (4 x uint16), then the bytes column by column
    pc = np.sort(rng.integers(0, 65536, size=(S, 4)), axis=1).astype("<u2")
    b1 = rng.integers(0, 256, size=(S, F), dtype=np.uint8)  # [col][row]
    p = mn + rg * inc16 * pc.astype(np.float32)  # [col][4]
    v = b1.astype(np.float32)
    p0, p25, p75, p100 = (p[:, k:k + 1] for k in range(4))
    lo = p0 + (p25 - p0) * v * f32(1.0 / 64.0)
    mid = p25 + (p75 - p25) * (v - f32(64)) * f32(1.0 / 128.0)
    hi = p75 + (p100 - p75) * (v - f32(192)) * f32(1.0 / 63.0)
    want = np.where(b1 <= 64, lo, np.where(b1 <= 192, mid, hi)).T
    check(b"CM " + head + pc.tobytes() + b1.tobytes(), want)
