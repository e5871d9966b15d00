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
