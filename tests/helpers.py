"""Shared helpers of the parity tests: run streams through the engine's C ABI with the test taps on."""
import json

import numpy as np


def feed_round_robin(recs, waves, bytes_per_call=8000):
    """Feeds like the reference's only batch driver [REF python/example/test_gpu_batch.py:27-51]."""
    pos = [0] * len(recs)
    ended = set()
    while len(ended) < len(recs):
        for i, r in enumerate(recs):
            if i in ended:
                continue
            data = waves[i][pos[i]:pos[i] + bytes_per_call // 2].tobytes()
            pos[i] += bytes_per_call // 2
            if not data:
                r.FinishStream()
                ended.add(i)
                continue
            r.AcceptWaveform(data)


def run_engine(model_dir, waves, options="", capture=True, bytes_per_call=8000, rate=16000.0):
    import vosk
    opts = "debug-capture=1," + options if capture else options
    model = vosk.BatchModel(model_dir, options=opts)
    recs = [vosk.BatchRecognizer(model, float(rate)) for _ in waves]
    if capture:
        for r in recs:
            r.DebugCapture()
    feed_round_robin(recs, waves, bytes_per_call)
    model.Wait()
    out = []
    for r in recs:
        d = {"text": r.Result()}
        if capture:
            d["mfcc"] = r.DebugGet("mfcc", np.float32).reshape(-1, 40)
            d["ivectors"] = r.DebugGet("ivectors", np.float32)
            d["loglikes"] = r.DebugGet("loglikes", np.float32)
            for k in ("frame_off", "tok_state", "tok_arc", "tok_prev"):
                d[k] = r.DebugGet(k, np.int32)
            d["tok_cost"] = r.DebugGet("tok_cost", np.float32)
            d["error"] = int(r.DebugGet("error", np.int32)[0])
            if "lattice=0" not in options:  # lattice generation is the default result path
                d["lat_hdr"] = r.DebugGet("lat_hdr", np.int32)
                d["lat_links"] = r.DebugGet("lat_links", np.int32).reshape(-1, 4)
                d["lat_final"] = r.DebugGet("lat_final", np.int32).reshape(-1, 2)
                d["lat_tok_frame"] = r.DebugGet("lat_tok_frame", np.int32)
                d["lat_tok_state"] = r.DebugGet("lat_tok_state", np.int32)
        out.append(d)
    stats = model.Stats()
    del recs
    del model
    return out, stats


def canonical_tokens(frame_off, state, cost, arc, prev):
    """Per frame: sort by state id and renumber the back pointers (DESIGN.md canonical token order)."""
    n = len(state)
    new_index = np.full(n, -1, dtype=np.int64)
    order_all = np.zeros(n, dtype=np.int64)
    for f in range(len(frame_off) - 1):
        lo, hi = int(frame_off[f]), int(frame_off[f + 1])
        o = lo + np.argsort(state[lo:hi], kind="stable")
        order_all[lo:hi] = o
        new_index[o] = np.arange(lo, hi)
    p = prev[order_all].astype(np.int64)
    p = np.where(p >= 0, new_index[np.maximum(p, 0)], p)
    return state[order_all], cost[order_all], arc[order_all], p


def canonical_lattice(frame, state, src, dst, arc, ac_bits, final_state, final_bits):
    """Order-free form of a raw lattice: states as sorted (frame, graph state), links as sorted
    (src frame, src state, dst frame, dst state, arc, acoustic-cost bits), finals as sorted (frame, state, cost bits)."""
    states = np.stack([frame, state], 1).astype(np.int64)
    states = states[np.lexsort(states.T[::-1])]
    links = np.stack([frame[src], state[src], frame[dst], state[dst], arc, ac_bits], 1).astype(np.int64)
    links = links[np.lexsort(links.T[::-1])]
    fin = np.stack([frame[final_state], state[final_state], final_bits], 1).astype(np.int64)
    fin = fin[np.lexsort(fin.T[::-1])]
    return states, links, fin


def words_of(text):
    return json.loads(text).get("text", "")


_LAT_HOOK = None


def oracle_lattice_text(model, dec, lattice_beam, offset=0.0, nlsml=False, rc=None):
    """Expected lattice-mode result text: the ORACLE's chain (oracle/orc_lattice.cc) on the oracle's raw lattice."""
    import oracle
    return oracle.lattice_result(model, dec, lattice_beam, offset=offset, nlsml=nlsml, rc=rc)


def lattice_text(model_dir, oracle_lattice, start, lattice_beam, stage=0):
    """ENGINE host lattice chain (vosk_b200_lattice_result, no GPU involved) on a raw lattice given as oracle.decode()["lattice"]."""
    import ctypes
    import os
    global _LAT_HOOK
    if _LAT_HOOK is None:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        lib = ctypes.CDLL(os.path.join(root, "vosk-api_b200", "lib", "libvosk.so"))
        f = lib.vosk_b200_lattice_result
        f.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                      ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
        f.restype = ctypes.c_int
        _LAT_HOOK = f
    lat = oracle_lattice
    src = np.ascontiguousarray(lat["src"], dtype=np.int32)
    dst = np.ascontiguousarray(lat["dst"], dtype=np.int32)
    arc = np.ascontiguousarray(lat["arc"], dtype=np.int32)
    ac = np.ascontiguousarray(lat["ac"], dtype=np.float32)
    fs = np.ascontiguousarray(lat["final_state"], dtype=np.int32)
    fc = np.ascontiguousarray(lat["final_cost"], dtype=np.float32)
    cap = 1 << 22
    buf = ctypes.create_string_buffer(cap)
    n = _LAT_HOOK(model_dir.encode(), len(lat["tok_index"]), int(start), len(src), src.ctypes.data, dst.ctypes.data, arc.ctypes.data,
                  ac.ctypes.data, len(fs), fs.ctypes.data, fc.ctypes.data, lattice_beam, stage, buf, cap)
    assert 0 <= n < cap, buf.value
    return buf.value.decode()


def oracle_lattice_start(dec):
    lat = dec["lattice"]
    return int(np.flatnonzero((lat["frame"] == 0) & (dec["arc"][lat["tok_index"]] < 0))[0])


def results_close(a, b, conf_tol=2e-2):
    """Two result texts: same words and word times, confidences within conf_tol."""
    ja, jb = json.loads(a), json.loads(b)
    if ja.get("text") != jb.get("text"):
        return False
    for x, y in zip(ja.get("result", []), jb.get("result", [])):
        if x["word"] != y["word"] or abs(x["start"] - y["start"]) > 1e-6 or abs(x["end"] - y["end"]) > 1e-6 or abs(x["conf"] - y["conf"]) > conf_tol:
            return False
    return True
