"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vosk-api_b200", "lib", "libvosk.so")


def declared_symbols(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vosk_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(ROOT, "vosk-api_b200")])
    return ctypes.CDLL(LIB)


@pytest.mark.parametrize("header", ["vosk_api.h", "vosk_b200.h"])
def test_every_declared_symbol_is_exported(lib, header):
    syms = declared_symbols(header)
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), s


def test_reference_batch_symbols_present(lib):
    # the 11 batch symbols + gpu init + log level of the reference header [REF src/vosk_api.h:287-346]
    for s in ["vosk_set_log_level", "vosk_gpu_init", "vosk_gpu_thread_init", "vosk_batch_model_new", "vosk_batch_model_free",
              "vosk_batch_model_wait", "vosk_batch_recognizer_new", "vosk_batch_recognizer_free", "vosk_batch_recognizer_accept_waveform",
              "vosk_batch_recognizer_set_nlsml", "vosk_batch_recognizer_finish_stream", "vosk_batch_recognizer_front_result",
              "vosk_batch_recognizer_pop", "vosk_batch_recognizer_get_pending_chunks"]:
        assert hasattr(lib, s), s


def test_cdef_parses_like_the_reference_builder():
    """The reference builds its cdef from `cpp vosk_api.h` [REF python/vosk_builder.py:6-11]; ours must parse the same way."""
    from cffi import FFI
    ffi = FFI()
    text = os.popen("cpp " + os.path.join(ROOT, "include", "vosk_api.h")).read()
    ffi.cdef(text)
    c = ffi.dlopen(LIB)
    assert c.vosk_batch_model_new is not None
    c.vosk_set_log_level(-1)


def test_null_handles_are_safe(lib):
    lib.vosk_batch_recognizer_front_result.restype = ctypes.c_char_p
    assert lib.vosk_batch_recognizer_front_result(None) == b""
    lib.vosk_batch_recognizer_pop(None)
    lib.vosk_batch_recognizer_accept_waveform(None, b"ab", 2)
    lib.vosk_batch_model_wait(None)
    assert lib.vosk_batch_recognizer_get_pending_chunks(None) == 0
    lib.vosk_batch_recognizer_new.restype = ctypes.c_void_p
    assert not lib.vosk_batch_recognizer_new(None, ctypes.c_float(16000.0))


def test_fails_loudly_without_gpu_or_model(lib, tmp_path):
    """No CPU fallback: model creation returns NULL (with a logged reason) when no device / no model is usable."""
    import torch
    lib.vosk_batch_model_new_ex.restype = ctypes.c_void_p
    lib.vosk_b200_last_error.restype = ctypes.c_char_p
    lib.vosk_set_log_level(-1)
    h = lib.vosk_batch_model_new_ex(str(tmp_path / "nope").encode(), b"")
    assert not h
    assert b"cannot open" in lib.vosk_b200_last_error()
    if not torch.cuda.is_available():
        sys.path.insert(0, os.path.join(ROOT, "vosk-api_b200", "tools"))
        import vbmodel
        mdir = vbmodel.write_model_dir(str(tmp_path), "tiny", 0)
        h = lib.vosk_batch_model_new_ex(mdir.encode(), b"num-channels=2,max-batch-size=2")
        assert not h
        assert b"CUDA" in lib.vosk_b200_last_error()


def test_reference_python_package_loads_our_library(tmp_path):
    """The unmodified reference cffi package dlopens libvosk.so next to itself [REF python/vosk/__init__.py:17-32]."""
    ref = "/root/reference/python/vosk/__init__.py"
    if not os.path.exists(ref):
        pytest.skip("reference not present on this machine")
    pkg = tmp_path / "vosk"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(open(ref).read())
    # what `python setup.py` would generate from vosk_builder.py, built from OUR header
    from cffi import FFI
    ffi = FFI()
    ffi.set_source("vosk.vosk_cffi", None)
    ffi.cdef(os.popen("cpp " + os.path.join(ROOT, "include", "vosk_api.h")).read())
    ffi.emit_python_code(str(pkg / "vosk_cffi.py"))
    os.symlink(LIB, pkg / "libvosk.so")
    code = ("import sys; sys.path.insert(0, %r); import vosk; vosk.GpuInit(); vosk.GpuThreadInit(); "
            "print('ok', hasattr(vosk, 'BatchModel'), hasattr(vosk, 'BatchRecognizer'))" % str(tmp_path))
    env = dict(os.environ, PYTHONPATH="")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert "ok True True" in out.stdout, out.stderr[-800:]


def test_every_symbol_of_the_reference_header_resolves(lib):
    """Eager binders (JNA Native.register [REF java/lib/src/main/java/org/vosk/LibVosk.java:38-41], cgo, P/Invoke) bind the
    whole reference header at load time: every function it declares must be exported — the batch / GPU half implemented,
    the CPU recognizer half as "not built" stubs (NULL / -1 / ""), the mirror image of [REF src/vosk_api.cc:198-282]."""
    import re
    def protos(text):  # declared functions (comments stripped): name -> prototype with single spaces
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        return {re.search(r"(vosk_[a-z_0-9]+)\s*\(", p).group(1): re.sub(r"\s+", " ", p).strip()
                for p in re.findall(r"[A-Za-z_][A-Za-z_ \*]*\bvosk_[a-z_0-9]+\s*\([^;{]*\)\s*;", text)}
    ours = protos(open(os.path.join(ROOT, "include", "vosk_api.h")).read())
    names = list(ours)
    ref = "/root/reference/src/vosk_api.h"
    if os.path.exists(ref):
        theirs = protos(open(ref).read())
        assert set(theirs) == set(ours), sorted(set(theirs) ^ set(ours))
        for n in theirs:  # identical prototypes
            assert theirs[n] == ours[n], (theirs[n], ours[n])
    assert len(set(names)) == 35  # 21 CPU + 14 batch / GPU / log functions
    for n in set(names):
        assert hasattr(lib, n), n
    lib.vosk_set_log_level(-2)
    lib.vosk_model_new.restype = ctypes.c_void_p
    lib.vosk_recognizer_new.restype = ctypes.c_void_p
    lib.vosk_recognizer_result.restype = ctypes.c_char_p
    assert not lib.vosk_model_new(b"model")
    assert not lib.vosk_recognizer_new(None, ctypes.c_float(16000.0))
    assert lib.vosk_recognizer_accept_waveform(None, b"ab", 2) == -1
    assert lib.vosk_model_find_word(None, b"x") == -1
    assert lib.vosk_recognizer_result(None) == b""
    lib.vosk_recognizer_free(None)
    lib.vosk_model_free(None)
