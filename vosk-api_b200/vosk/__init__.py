"""Python mirror of the reference's `vosk` package for the batch path.

Same class and method names as the reference binding [REF python/vosk/__init__.py:185-235]
(`GpuInit`, `GpuThreadInit`, `BatchModel`, `BatchRecognizer` with `AcceptWaveform`, `Result`,
`FinishStream`, `GetPendingChunks`, `Wait`) so `python/example/test_gpu_batch.py` runs unchanged
against this package.  cffi ABI mode, cdef taken from include/vosk_api.h exactly as the reference
builds its cdef from src/vosk_api.h [REF python/vosk_builder.py:6-11].  Fails loudly when
libvosk.so (the CUDA engine) is missing: there is no CPU fallback.
"""
import os
import re

from cffi import FFI

_here = os.path.dirname(os.path.abspath(__file__))
_root = os.path.dirname(_here)
_inc = os.path.join(os.path.dirname(_root), "include")


def _cdef_text():
    text = ""
    for name in ("vosk_api.h", "vosk_b200.h"):
        src = open(os.path.join(_inc, name)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = "\n".join(l for l in src.splitlines() if not l.strip().startswith("#") and "extern \"C\"" not in l and l.strip() != "}")
        text += src + "\n"
    return text


_ffi = FFI()
_ffi.cdef(_cdef_text())


def _load():
    path = os.environ.get("VOSK_B200_LIB", os.path.join(_root, "lib", "libvosk.so"))
    if not os.path.exists(path):
        raise ImportError(f"{path} not found: build it with `make -C {_root}` (CUDA engine; no CPU fallback)")
    return _ffi.dlopen(path)


_c = _load()
LIB_PATH = os.environ.get("VOSK_B200_LIB", os.path.join(_root, "lib", "libvosk.so"))


def SetLogLevel(level):
    return _c.vosk_set_log_level(level)


def GpuInit():
    _c.vosk_gpu_init()


def GpuThreadInit():
    _c.vosk_gpu_thread_init()


class BatchModel(object):
    def __init__(self, *args, **kw):
        # reference: BatchModel() takes no argument and loads ./model [REF python/vosk/__init__.py:199-203]
        model_path = kw.get("model_path") or (args[0] if args else None)
        options = kw.get("options", "")
        if model_path is None and not options:
            self._handle = _c.vosk_batch_model_new()
        else:
            self._handle = _c.vosk_batch_model_new_ex((model_path or "model").encode(), options.encode())
        if self._handle == _ffi.NULL:
            raise Exception("Failed to create a model: " + _ffi.string(_c.vosk_b200_last_error()).decode())

    def __del__(self):
        if getattr(self, "_handle", None) not in (None, _ffi.NULL):
            _c.vosk_batch_model_free(self._handle)
            self._handle = _ffi.NULL

    def Wait(self):
        _c.vosk_batch_model_wait(self._handle)

    # ---- additive surface (include/vosk_b200.h) ----
    def SamplesPerChunk(self):
        return _c.vosk_batch_model_samples_per_chunk(self._handle)

    def Stats(self):
        buf = _ffi.new("double[81]")
        n = _c.vosk_batch_model_stats(self._handle, buf, 81)
        keys = ["audio_seconds", "steps", "lanes", "launches", "tokens", "arcs_emitting", "arcs_epsilon", "tokens_new",
                "ms_feat", "ms_ivector", "ms_nnet", "ms_search", "gemm_launches",
                "lane_cycles_sum", "lane_cycles_max", "max_tokens_per_frame", "lane_launches", "host_launch_ms",
                "arcs_staged", "links", "lattice_arcs"]
        keys += ["cyc_%s_%s" % (v, ph) for v in ("heavy", "light") for ph in ("cutoff", "rank", "log", "gather", "insert", "closure", "finalize", "x")]
        keys += ["resample_segments", "truncated", "lattice_fallbacks", "post_ms", "post_jobs", "post_threads", "ms_prune", "host_complete_ms", "host_fetch_ms", "h2d_bytes", "d2h_bytes"]
        keys += ["slowest_%s_%s" % (v, ph) for v in ("t1024", "t512", "t256") for ph in ("cutoff", "rank", "log", "gather", "insert", "closure", "finalize", "x")]
        keys += ["slowest_cycles_%s" % v for v in ("t1024", "t512", "t256")] + ["slowest_tokens_%s" % v for v in ("t1024", "t512", "t256")]
        keys += ["lane_launches_%s" % v for v in ("t1024", "t512", "t256")]
        return {k: buf[i] for i, k in enumerate(keys[:n])}

    def Latency(self, reset=False):
        buf = _ffi.new("double[5]")
        _c.vosk_batch_model_latency(self._handle, buf, int(reset))
        return dict(p50=buf[0], p90=buf[1], p99=buf[2], mean=buf[3], count=int(buf[4]))

    def FeedStreams(self, waves, bytes_per_call=8000, threads=8, want_results=True, passes=1):
        """Native multi-threaded feeder (vosk_b200_feed_streams_passes): waves = list of int16 numpy arrays, fed `passes` times
        over before the one Wait; returns the last pass's result texts (self.feed_mismatches = streams of earlier passes whose
        text differed)."""
        import numpy as np
        arrs = [np.ascontiguousarray(w, dtype=np.int16) for w in waves]
        n = len(arrs)
        ptrs = _ffi.new("int16_t *[]", [_ffi.cast("int16_t *", a.ctypes.data) for a in arrs])
        lens = _ffi.new("int[]", [len(a) for a in arrs])
        res = _ffi.new("char *[]", n) if want_results else _ffi.NULL
        bad = _ffi.new("int *")
        rc = _c.vosk_b200_feed_streams_passes(self._handle, ptrs, lens, n, int(bytes_per_call), int(threads), int(passes), res, bad)
        self.feed_mismatches = int(bad[0])
        if rc != 0:
            raise Exception("feed_streams failed")
        out = []
        if want_results:
            for i in range(n):
                out.append(_ffi.string(res[i]).decode("utf-8"))
                _c.vosk_b200_free(res[i])
        return out

    def ResetStats(self):
        _c.vosk_batch_model_reset_stats(self._handle)

    def SetTiming(self, on):
        _c.vosk_batch_model_set_timing(self._handle, int(on))

    def SetSlots(self, n):
        _c.vosk_batch_model_set_slots(self._handle, int(n))

    def RunResident(self, audio, lengths=None, passes=1):
        """audio: C-contiguous int16 numpy array [streams, samples] (+ optional valid length per row);
        returns (device_ms, [result text]).  passes > 1: the streams are decoded that many times over, back to back
        (vosk_batch_model_run_resident_passes); self.resident_mismatches = streams whose text differed between passes."""
        import numpy as np
        a = np.ascontiguousarray(audio, dtype=np.int16)
        lp = _ffi.NULL
        if lengths is not None:
            ln = np.ascontiguousarray(lengths, dtype=np.int32)
            lp = _ffi.cast("int *", ln.ctypes.data)
        bad = _ffi.new("int *")
        ms = _c.vosk_batch_model_run_resident_passes(self._handle, _ffi.cast("int16_t *", a.ctypes.data), a.shape[0], a.shape[1], lp, int(passes), bad)
        self.resident_mismatches = int(bad[0])
        if ms < 0:
            raise RuntimeError("run_resident failed")
        return ms, [_ffi.string(_c.vosk_batch_model_resident_result(self._handle, i)).decode() for i in range(a.shape[0])]


class BatchRecognizer(object):
    def __init__(self, *args):
        self._model = args[0]
        self._handle = _c.vosk_batch_recognizer_new(args[0]._handle, args[1])
        if self._handle == _ffi.NULL:
            raise Exception("Failed to create a recognizer")

    def __del__(self):
        if getattr(self, "_handle", None) not in (None, _ffi.NULL):
            _c.vosk_batch_recognizer_free(self._handle)
            self._handle = _ffi.NULL

    def AcceptWaveform(self, data):
        _c.vosk_batch_recognizer_accept_waveform(self._handle, data, len(data))

    def Result(self):
        ptr = _c.vosk_batch_recognizer_front_result(self._handle)
        res = _ffi.string(ptr).decode("utf-8")
        _c.vosk_batch_recognizer_pop(self._handle)
        return res

    def FinishStream(self):
        _c.vosk_batch_recognizer_finish_stream(self._handle)

    def GetPendingChunks(self):
        return _c.vosk_batch_recognizer_get_pending_chunks(self._handle)

    def SetNLSML(self, on):
        _c.vosk_batch_recognizer_set_nlsml(self._handle, int(on))

    # ---- additive surface: partial results (model option partials=1) ----
    def PartialResult(self):
        return _ffi.string(_c.vosk_batch_recognizer_partial_result(self._handle)).decode()

    def PartialFrames(self):
        return _c.vosk_batch_recognizer_partial_frames(self._handle)

    # ---- additive surface: test taps ----
    def DebugCapture(self):
        _c.vosk_batch_recognizer_debug_capture(self._handle)

    def DebugGet(self, what, dtype):
        import numpy as np
        n = _c.vosk_batch_recognizer_debug_get(self._handle, what.encode(), _ffi.NULL, 0)
        if n < 0:
            raise KeyError(what)
        out = np.zeros(n // np.dtype(dtype).itemsize, dtype=dtype)
        if n:
            _c.vosk_batch_recognizer_debug_get(self._handle, what.encode(), _ffi.cast("void *", out.ctypes.data), n)
        return out
