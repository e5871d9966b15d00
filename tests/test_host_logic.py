"""CPU tests of the host-side logic: model loaders, chunk schedule, result writer, resampler, sharding, gloo reduce."""
import ctypes
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vosk-api_b200", "lib", "libvosk.so")


@pytest.fixture(scope="module")
def lib():
    return ctypes.CDLL(LIB)


def test_engine_loader_reads_the_model_dir(lib, model_root):
    """The engine's C++ loaders (VBT container, const FST, conf files) see the same model as the numpy reader."""
    import vbmodel
    mdir = model_root("tiny")
    buf = ctypes.create_string_buffer(1024)
    assert lib.vosk_b200_model_check(mdir.encode(), buf, 1024) == 0, buf.value
    info = dict(kv.split("=") for kv in buf.value.decode().split())
    m = vbmodel.load_model_dir(mdir)
    g = m["graph"]
    assert int(info["states"]) == g["num_states"] and int(info["arcs"]) == g["num_arcs"]
    assert int(info["eps"]) == int((g["arc_pdf"] < 0).sum())
    assert abs(float(info["wsum"]) - float(g["arc_w"].astype(np.float64).sum())) < 1e-2
    assert int(info["context"]) == vbmodel.context_of(vbmodel.ARCHS["tiny"])[0]
    assert float(info["beam"]) == 13.0 and int(info["max_active"]) == 7000 and int(info["fpc"]) == 51


def test_vector_and_const_fst_load_identically(lib, model_root, tmp_path):
    """HCLG.fst as OpenFst 'vector' must load to the same graph as the 'const' file."""
    import shutil
    import struct
    import vbmodel
    src = model_root("tiny")
    dst = str(tmp_path / "model")
    shutil.copytree(src, dst)
    fst = vbmodel.read_fst(os.path.join(src, "graph", "HCLG.fst"))
    ns = len(fst["final"])
    with open(os.path.join(dst, "graph", "HCLG.fst"), "wb") as f:
        f.write(struct.pack("<i", vbmodel.FST_MAGIC))
        for s in (b"vector", b"standard"):
            f.write(struct.pack("<i", len(s)) + s)
        f.write(struct.pack("<iiQqqq", 2, 0, 0, fst["start"], ns, int(fst["row"][-1])))
        for s in range(ns):
            lo, hi = int(fst["row"][s]), int(fst["row"][s + 1])
            f.write(struct.pack("<fq", float(fst["final"][s]), hi - lo))
            for a in range(lo, hi):
                f.write(struct.pack("<iifi", int(fst["ilabel"][a]), int(fst["olabel"][a]), float(fst["weight"][a]), int(fst["next"][a])))
    a, b = ctypes.create_string_buffer(1024), ctypes.create_string_buffer(1024)
    assert lib.vosk_b200_model_check(src.encode(), a, 1024) == 0 and lib.vosk_b200_model_check(dst.encode(), b, 1024) == 0
    assert a.value == b.value
    again = vbmodel.read_fst(os.path.join(dst, "graph", "HCLG.fst"))
    for k in ("ilabel", "olabel", "weight", "next", "final"):
        np.testing.assert_array_equal(again[k], fst[k])


def test_loader_rejects_corrupt_files(lib, model_root, tmp_path):
    import shutil
    dst = str(tmp_path / "model")
    shutil.copytree(model_root("tiny"), dst)
    p = os.path.join(dst, "graph", "HCLG.fst")
    data = open(p, "rb").read()
    open(p, "wb").write(data[: len(data) // 2])
    buf = ctypes.create_string_buffer(1024)
    assert lib.vosk_b200_model_check(dst.encode(), buf, 1024) == -1
    assert b"truncated" in buf.value
    os.remove(os.path.join(dst, "am", "final.mdl"))
    assert lib.vosk_b200_model_check(dst.encode(), buf, 1024) == -1


def test_generator_is_deterministic(tmp_path):
    import vbmodel
    a = vbmodel.write_model_dir(str(tmp_path / "a"), "tiny", 0)
    b = vbmodel.write_model_dir(str(tmp_path / "b"), "tiny", 0)
    for rel in ("am/final.mdl", "graph/HCLG.fst", "ivector/final.ie", "graph/words.txt"):
        assert open(os.path.join(a, rel), "rb").read() == open(os.path.join(b, rel), "rb").read()
    g = vbmodel.load_model_dir(a)["graph"]
    # canonical CSR: emitting arcs first, then epsilons; every arc's source is the state whose range holds it
    for s in (0, 1, 5, g["num_states"] - 1):
        lo, mid, hi = g["e_begin"][s], g["eps_begin"][s], g["e_begin"][s + 1]
        assert np.all(g["arc_pdf"][lo:mid] >= 0) and np.all(g["arc_pdf"][mid:hi] < 0)
        assert np.all(g["arc_src"][lo:hi] == s)
    assert np.all(g["arc_w"][g["arc_pdf"] < 0] >= 0)  # the token log requires non-negative epsilon weights


def test_chunk_plan_properties(oracle_lib):
    """Chunk schedule shared by oracle and engine: monotone, complete, chunk-size independent frame count."""
    ctx = 26
    for n in (0, 399, 400, 8159, 8160, 8161, 16320, 50000):
        for fpc in (51, 9, 12):
            ends, avail, iv = oracle_lib.chunk_plan(n, fpc, ctx)
            T = oracle_lib.num_frames(n)
            assert len(ends) == n // (fpc * 160) + 1
            assert ends[-1] == T and avail[-1] == T
            assert np.all(np.diff(ends) >= 0) and np.all(ends <= avail)
            assert len(iv) == max(T + 2 * (ctx - 2), 0)
            assert np.all(np.diff(iv) >= 0) and (len(iv) == 0 or iv[-1] == len(ends) - 1)


def test_result_writer_matches_reference_json_h(lib):
    """Engine result text == the reference's json.h dump (golden fixture generated from oracle/_ref)."""
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "json_ref.json")))
    for c in gold:
        n = len(c["words"])
        # the engine formats frame indices; pick cases whose times are whole frames
        begin = [int(round(x / 0.03)) for x in c["start"]]
        end = [int(round(x / 0.03)) for x in c["end"]]
        if any(abs(b * 0.03 - x) > 1e-9 for b, x in zip(begin, c["start"])) or any(abs(e * 0.03 - x) > 1e-9 for e, x in zip(end, c["end"])):
            continue
        words = (ctypes.c_char_p * max(n, 1))(*[w.encode() for w in c["words"]])
        ib = (ctypes.c_int * max(n, 1))(*begin)
        ie = (ctypes.c_int * max(n, 1))(*end)
        cf = (ctypes.c_float * max(n, 1))(*c["conf"])
        out = ctypes.create_string_buffer(4096)
        lib.vosk_b200_format_result(words, ib, ie, cf, n, ctypes.c_float(0.0), out, 4096)
        if all(abs(float(np.float32(x)) - x) < 1e-12 for x in c["conf"]):
            assert out.value.decode() == c["dump"]
        else:
            assert json.loads(out.value.decode())["text"] == c["text"]


def test_resampler(lib):
    lib.vosk_b200_resample.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
    x = (1000 * np.sin(2 * np.pi * 440 * np.arange(8000) / 8000.0)).astype(np.float32)
    out = np.zeros(20000, dtype=np.float32)
    n = lib.vosk_b200_resample(x.ctypes.data, len(x), 8000.0, out.ctypes.data, len(out))
    assert n == 16000  # Kaldi GetNumOutputSamples with flush=true: every input tick yields out/in samples
    y = out[:n]
    ref = 1000 * np.sin(2 * np.pi * 440 * np.arange(n) / 16000.0)
    assert np.abs(y[200:-200] - ref[200:-200]).max() < 15.0  # band-limited interpolation away from the flushed edges
    z = np.zeros(100, dtype=np.float32)
    x16 = np.arange(100, dtype=np.float32)
    assert lib.vosk_b200_resample(x16.ctypes.data, 100, 16000.0, z.ctypes.data, 100) == 100
    np.testing.assert_array_equal(z, x16)  # 16 kHz in: identity
    n = lib.vosk_b200_resample(x.ctypes.data, 441, 44100.0, out.ctypes.data, len(out))
    assert n == 160


def test_stream_sharding_rule(lib):
    lib.vosk_b200_device_for_stream.argtypes = [ctypes.c_ulonglong, ctypes.c_int]
    for g in (1, 2, 4, 8):
        owners = [lib.vosk_b200_device_for_stream(i, g) for i in range(64)]
        assert set(owners) == set(range(g))
        assert max(np.bincount(owners)) - min(np.bincount(owners)) == 0


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    audio, secs = 100.0 * (rank + 1), 2.0 + rank          # rank 0: 100 s in 2 s, rank 1: 200 s in 3 s
    q.put((rank, bench.job_throughput(audio, secs), bench.reduce_over_ranks(secs, "max"), bench.reduce_over_ranks(audio, "sum")))
    dist.destroy_process_group()


def test_world_size_2_aggregation_gloo():
    """N>1 path of bench.py: utterance-sharded ranks, whole-job value = sum(audio) / max(time), no data-path collective."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    for rank, val, tmax, asum in res:
        assert tmax == 3.0 and asum == 300.0 and abs(val - 100.0) < 1e-9


def test_mfcc_conf_is_read_and_checked(lib, model_root, tmp_path):
    """conf/mfcc.conf [REF src/batch_model.cc:75-76]: band edges are taken from it (7600 Hz = Nyquist - 400), a feature geometry
    the kernels are not built for is refused at load time instead of silently decoding with other features."""
    import shutil
    dst = str(tmp_path / "model")
    shutil.copytree(model_root("tiny"), dst)
    buf = ctypes.create_string_buffer(1024)
    conf = os.path.join(dst, "conf", "mfcc.conf")
    base = open(conf).read()
    open(conf, "w").write(base.replace("--high-freq=-400", "--high-freq=7600") + "--sample-frequency=16000\n")
    assert lib.vosk_b200_model_check(dst.encode(), buf, 1024) == 0, buf.value
    open(conf, "w").write(base.replace("--num-mel-bins=40", "--num-mel-bins=23"))
    assert lib.vosk_b200_model_check(dst.encode(), buf, 1024) == -1 and b"num-mel-bins" in buf.value
    open(conf, "w").write(base.replace("--low-freq=20", "--low-freq=9000"))
    assert lib.vosk_b200_model_check(dst.encode(), buf, 1024) == -1 and b"low-freq" in buf.value


@pytest.mark.parametrize("rate,n", [(8000, 4000), (44100, 4410), (22050, 3000), (48000, 4800), (11025, 2000)])
def test_resampler_matches_torchaudio(lib, rate, n):
    """Independent pin of the LinearResample restatement (filter design, phase tables, zero-padded edges of a flushed call,
    output count): torchaudio's sinc_interp_hann resampler is derived from Kaldi's LinearResample; with rolloff 1.0 and filter
    width 6 it is the filter the reference asks for — cutoff min(rate, 16000) / 2, num_zeros 6 [REF src/batch_recognizer.cc:27-29]."""
    torch = pytest.importorskip("torch")
    torchaudio = pytest.importorskip("torchaudio")
    lib.vosk_b200_resample.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
    x = (np.random.default_rng(rate).standard_normal(n) * 1000).astype(np.float32)
    out = np.zeros(3 * n + 64, dtype=np.float32)
    m = lib.vosk_b200_resample(x.ctypes.data, n, float(rate), out.ctypes.data, len(out))
    ref = torchaudio.functional.resample(torch.from_numpy(x)[None].double(), rate, 16000, lowpass_filter_width=6, rolloff=1.0)[0].numpy()
    assert m == len(ref)
    assert np.abs(out[:m] - ref).max() < 2e-6 * np.abs(ref).max() + 1e-3   # fp32 products / sums against fp64
