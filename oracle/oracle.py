"""ctypes front end of the CPU oracle (liboracle.so).  TEST INFRASTRUCTURE ONLY — see oracle/oracle.h.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
never by anything under vosk-api_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        so = os.environ.get("ORC_SO") or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_num_frames.argtypes = [C.c_int64]
        _LIB.orc_mfcc.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        _LIB.orc_ivectors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _LIB.orc_nnet_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.orc_decode.restype = C.c_void_p
        _LIB.orc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        _LIB.orc_decoder_free.argtypes = [C.c_void_p]
        _LIB.orc_decoder_num_frames.argtypes = [C.c_void_p]
        _LIB.orc_decoder_num_tokens.argtypes = [C.c_void_p]
        _LIB.orc_decoder_num_tokens.restype = C.c_int64
        _LIB.orc_decoder_tokens.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        _LIB.orc_decoder_best_path.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _LIB.orc_decoder_cost_offsets.argtypes = [C.c_void_p, C.c_void_p]
        _LIB.orc_decoder_num_links.argtypes = [C.c_void_p]
        _LIB.orc_decoder_num_links.restype = C.c_int64
        _LIB.orc_decoder_lattice.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB.orc_decoder_lattice.restype = C.c_int64
        _LIB.orc_align_words.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _LIB.orc_result_json.restype = C.c_void_p
        _LIB.orc_result_json.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_int]
        _LIB.orc_free.argtypes = [C.c_void_p]
        _LIB.orc_lattice_result.restype = C.c_void_p
        _LIB.orc_lattice_result.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_float, C.c_double, C.c_float,
                                            C.c_int, C.c_int, C.c_int]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _IvecParams(C.Structure):
    _fields_ = [("feat_dim", C.c_int), ("ivec_dim", C.c_int), ("num_gauss", C.c_int),
                ("splice_left", C.c_int), ("splice_right", C.c_int), ("num_gselect", C.c_int),
                ("min_post", C.c_float), ("posterior_scale", C.c_float), ("max_count", C.c_float),
                ("cmn_window", C.c_int), ("global_frames", C.c_int),
                ("lda", C.c_void_p), ("gconsts", C.c_void_p), ("means_invvars", C.c_void_p),
                ("inv_vars", C.c_void_p), ("M", C.c_void_p), ("sigma_inv", C.c_void_p),
                ("prior_offset", C.c_float), ("global_cmvn", C.c_void_p)]


class _NnetParams(C.Structure):
    _fields_ = [("feat_dim", C.c_int), ("ivec_dim", C.c_int), ("hidden", C.c_int), ("bottleneck", C.c_int),
                ("num_tdnnf", C.c_int), ("strides", C.c_void_p),
                ("prefinal_small", C.c_int), ("prefinal_big", C.c_int), ("num_pdfs", C.c_int),
                ("bypass_scale", C.c_float), ("tensors", C.c_void_p)]


class _Graph(C.Structure):
    _fields_ = [("num_states", C.c_int), ("num_arcs", C.c_int), ("start", C.c_int),
                ("final_cost", C.c_void_p), ("e_begin", C.c_void_p), ("eps_begin", C.c_void_p),
                ("arc_w", C.c_void_p), ("arc_next", C.c_void_p), ("arc_pdf", C.c_void_p), ("arc_src", C.c_void_p)]


class _Opts(C.Structure):
    _fields_ = [("beam", C.c_float), ("max_active", C.c_int), ("min_active", C.c_int), ("beam_delta", C.c_float)]


class _ResultCtx(C.Structure):
    _fields_ = [("arc_ilabel", C.c_void_p), ("arc_olabel", C.c_void_p), ("tid2phone", C.c_void_p),
                ("phone_type", C.c_void_p), ("num_phones", C.c_int), ("words", C.c_void_p), ("num_words", C.c_int)]


_PHONE_TYPES = {"nonword": 1, "begin": 2, "end": 3, "internal": 4, "singleton": 5}


def num_frames(n):
    return lib().orc_num_frames(int(n))


def mfcc(wave):
    w = np.ascontiguousarray(wave, dtype=np.int16)
    T = num_frames(len(w))
    out = np.zeros((T, 40), dtype=np.float32)
    if T:
        lib().orc_mfcc(_p(w), len(w), _p(out))
    return out


def nnet_tensor_list(model):
    n = model["nnet"]
    strides = [int(s) for s in model["cfg"]["tdnnf-strides"].split()]
    names = ["tdnn1.w", "tdnn1.b", "tdnn1.bn_scale", "tdnn1.bn_offset"]
    for k in range(2, 2 + len(strides)):
        names += [f"tdnnf{k}.linear.w", f"tdnnf{k}.affine.w", f"tdnnf{k}.affine.b", f"tdnnf{k}.bn_scale", f"tdnnf{k}.bn_offset"]
    names += ["prefinal_l.w", "prefinal.affine.w", "prefinal.affine.b", "prefinal.bn_scale", "prefinal.bn_offset",
              "prefinal.linear.w", "prefinal.linear.b", "output.w", "output.b"]
    return strides, [_f32(n[k]) for k in names]


def model_context(model):
    return 2 + sum(int(s) for s in model["cfg"]["tdnnf-strides"].split())


def chunk_plan(num_samples, frames_per_chunk, ctx):
    """Chunk schedule shared by oracle and engine (DESIGN.md 'chunk semantics').

    Returns (ends, avail, iv_index): chunk k (k < nfull) closes after (k+1)*fpc*160 samples; the last
    chunk flushes the remainder.  i-vector k sees spliced frames [0, ends[k]); tdnn1 frame t uses the
    i-vector of the first chunk in which mfcc(t+2) exists (or the last chunk for the padded tail)."""
    cs = frames_per_chunk * 160
    nfull = num_samples // cs
    T = num_frames(num_samples)
    avail = [num_frames((k + 1) * cs) for k in range(nfull)] + [T]
    ends = [max(a - 3, 0) for a in avail[:-1]] + [T]
    nchunks = nfull + 1
    lo, hi = -(ctx - 2), T - 1 + (ctx - 2)
    iv_index = np.zeros(max(hi - lo + 1, 0), dtype=np.int32)
    for t in range(lo, hi + 1):
        k = nchunks - 1
        for c in range(nfull):
            if avail[c] - 2 > t:
                k = c
                break
        iv_index[t - lo] = k
    return np.asarray(ends, dtype=np.int32), np.asarray(avail, dtype=np.int32), iv_index


def ivectors(model, feats, ends, avail):
    iv, cfg = model["ivector"], model["cfg"]
    F, D = int(cfg["feat-dim"]), int(cfg["ivector-dim"])
    keep = dict(lda=_f32(iv["lda"]), gconsts=_f32(iv["dubm"]["gconsts"]),
                mi=_f32(iv["dubm"]["means_invvars"]), ivr=_f32(iv["dubm"]["inv_vars"]),
                M=_f32(iv["ie"]["M"]), si=_f32(iv["ie"]["sigma_inv"]),
                cm=np.ascontiguousarray(iv["cmvn"], dtype=np.float64))
    p = _IvecParams(F, D, keep["gconsts"].shape[0], 3, 3, 5, 0.025, 0.1, 100.0, 600, 200,
                    _p(keep["lda"]), _p(keep["gconsts"]), _p(keep["mi"]), _p(keep["ivr"]), _p(keep["M"]), _p(keep["si"]),
                    float(iv["ie"]["prior_offset"][0]), _p(keep["cm"]))
    feats = _f32(feats)
    ends, avail = _i32(ends), _i32(avail)
    out = np.zeros((len(ends), D), dtype=np.float32)
    lib().orc_ivectors(C.byref(p), _p(feats), len(feats), _p(ends), _p(avail), len(ends), _p(out))
    return out


def nnet_forward(model, feats, ivecs, iv_index):
    cfg = model["cfg"]
    strides, tensors = nnet_tensor_list(model)
    st = _i32(strides)
    ptrs = (C.c_void_p * len(tensors))(*[t.ctypes.data for t in tensors])
    p = _NnetParams(int(cfg["feat-dim"]), int(cfg["ivector-dim"]), int(cfg["hidden-dim"]), int(cfg["bottleneck-dim"]),
                    len(strides), _p(st), int(cfg["prefinal-small"]), int(cfg["prefinal-big"]), int(cfg["num-pdfs"]),
                    float(cfg["bypass-scale"]), C.cast(ptrs, C.c_void_p))
    feats, ivecs, iv_index = _f32(feats), _f32(ivecs), _i32(iv_index)
    T = len(feats)
    out = np.zeros(((T + 2) // 3, int(cfg["num-pdfs"])), dtype=np.float32)
    if T:
        lib().orc_nnet_forward(C.byref(p), _p(feats), T, _p(ivecs), _p(iv_index), _p(out))
    return out


def decode_opts(model, **over):
    c = model["conf"]
    o = dict(beam=float(c.get("beam", 13.0)), max_active=int(c.get("max-active", 7000)),
             min_active=int(c.get("min-active", 200)), beam_delta=0.5)
    o.update(over)
    return o


def decode(model, loglikes, **over):
    g = model["graph"]
    o = decode_opts(model, **over)
    G = _Graph(g["num_states"], g["num_arcs"], g["start"], _p(g["final"]), _p(g["e_begin"]), _p(g["eps_begin"]),
               _p(g["arc_w"]), _p(g["arc_next"]), _p(g["arc_pdf"]), _p(g["arc_src"]))
    O = _Opts(o["beam"], o["max_active"], o["min_active"], o["beam_delta"])
    ll = _f32(loglikes)
    L = lib()
    h = L.orc_decode(C.byref(G), C.byref(O), _p(ll), ll.shape[0], ll.shape[1] if ll.ndim == 2 else int(model["cfg"]["num-pdfs"]))
    try:
        nf = L.orc_decoder_num_frames(h)
        nt = L.orc_decoder_num_tokens(h)
        offsets = np.zeros(nf + 2, dtype=np.int64)
        state = np.zeros(nt, dtype=np.int32)
        cost = np.zeros(nt, dtype=np.float32)
        arc = np.zeros(nt, dtype=np.int32)
        prev = np.zeros(nt, dtype=np.int64)
        L.orc_decoder_tokens(h, _p(offsets), _p(state), _p(cost), _p(arc), _p(prev))
        cap = nf + 4 * 1024 + 16 * nf
        arcs = np.zeros(cap, dtype=np.int32)
        tot = C.c_float()
        rf = C.c_int()
        n = L.orc_decoder_best_path(h, _p(arcs), cap, C.byref(tot), C.byref(rf))
        out = dict(frames=nf, offsets=offsets, state=state, cost=cost, arc=arc, prev=prev,
                   best_arcs=arcs[:n].copy(), best_cost=tot.value, reached_final=bool(rf.value))
        if "lattice_beam" in over or "lattice-beam" in model["conf"]:
            lb = float(over.get("lattice_beam", model["conf"].get("lattice-beam", 6.0)))
            cap = int(L.orc_decoder_num_links(h))
            tok_index = np.zeros(max(nt, 1), dtype=np.int64)
            lsrc = np.zeros(max(cap, 1), dtype=np.int64)
            ldst = np.zeros(max(cap, 1), dtype=np.int64)
            larc = np.zeros(max(cap, 1), dtype=np.int32)
            lac = np.zeros(max(cap, 1), dtype=np.float32)
            fst = np.zeros(max(nt, 1), dtype=np.int64)
            fco = np.zeros(max(nt, 1), dtype=np.float32)
            ns = C.c_int64()
            nfin = C.c_int64()
            nl = L.orc_decoder_lattice(h, C.byref(G), lb, C.byref(ns), _p(tok_index), _p(lsrc), _p(ldst), _p(larc), _p(lac),
                                       cap, _p(fst), _p(fco), C.byref(nfin))
            coff = np.zeros(max(nf, 1), dtype=np.float32)
            L.orc_decoder_cost_offsets(h, _p(coff))
            out["cost_offset"] = coff[:nf].copy()
            frame_of = np.searchsorted(offsets, tok_index[:ns.value], side="right") - 1
            out["lattice"] = dict(tok_index=tok_index[:ns.value].copy(), frame=frame_of.astype(np.int32),
                                  state=state[tok_index[:ns.value]].copy(), src=lsrc[:nl].copy(), dst=ldst[:nl].copy(),
                                  arc=larc[:nl].copy(), ac=lac[:nl].copy(), final_state=fst[:nfin.value].copy(),
                                  final_cost=fco[:nfin.value].copy(), num_links_unpruned=cap)
        return out
    finally:
        L.orc_decoder_free(h)


class ResultCtx:
    def __init__(self, model):
        g = model["graph"]
        self.keep = [g["arc_ilabel"], g["arc_olabel"], _i32(model["nnet"]["tid2phone"])]
        nph = max(model["word_boundary"]) if model["word_boundary"] else 0
        pt = np.zeros(nph + 1, dtype=np.int32)
        for ph, kind in model["word_boundary"].items():
            pt[ph] = _PHONE_TYPES[kind]
        self.keep.append(pt)
        nw = max(model["words"]) + 1
        self.strs = [model["words"].get(i, "").encode() for i in range(nw)]
        self.arr = (C.c_char_p * nw)(*self.strs)
        self.ctx = _ResultCtx(_p(self.keep[0]), _p(self.keep[1]), _p(self.keep[2]), _p(pt), nph,
                              C.cast(self.arr, C.c_void_p), nw)


def align_words(model, arcs, rc=None):
    rc = rc or ResultCtx(model)
    arcs = _i32(arcs)
    cap = len(arcs) + 1
    w, b, e = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    n = lib().orc_align_words(C.byref(rc.ctx), _p(arcs), len(arcs), _p(w), _p(b), _p(e), cap)
    return w[:n], b[:n], e[:n]


def result_json(model, arcs, offset=0.0, nlsml=False, rc=None):
    rc = rc or ResultCtx(model)
    arcs = _i32(arcs)
    ptr = lib().orc_result_json(C.byref(rc.ctx), _p(arcs), len(arcs), offset, int(nlsml))
    s = C.string_at(ptr).decode()
    lib().orc_free(ptr)
    return s


def lattice_start(dec):
    """Lattice state of the start token (frame 0, no incoming arc)."""
    lat = dec["lattice"]
    return int(np.flatnonzero((lat["frame"] == 0) & (dec["arc"][lat["tok_index"]] < 0))[0])


def lattice_result(model, dec, lattice_beam=None, lm_scale=0.9, offset=0.0, nlsml=False, stage=0, phone_pass=True, rc=None):
    """The reference's result chain on the raw lattice of decode(..., lattice_beam=...): phone-pruned determinization, graph
    scale 0.9, word alignment, MBR, text [REF src/batch_recognizer.cc:43-107,138-149]."""
    rc = rc or ResultCtx(model)
    lat = dec["lattice"]
    lb = float(lattice_beam if lattice_beam is not None else model["conf"].get("lattice-beam", 6.0))
    g = model["graph"]
    src, dst = np.ascontiguousarray(lat["src"], dtype=np.int64), np.ascontiguousarray(lat["dst"], dtype=np.int64)
    arc, ac = _i32(lat["arc"]), _f32(lat["ac"])
    fs, fc = np.ascontiguousarray(lat["final_state"], dtype=np.int64), _f32(lat["final_cost"])
    arc_w = _f32(g["arc_w"])
    flags = model["nnet"].get("tid_flags")
    flags = None if flags is None else np.ascontiguousarray(flags, dtype=np.uint8)
    ptr = lib().orc_lattice_result(C.byref(rc.ctx), _p(arc_w), None if flags is None else _p(flags), len(model["nnet"]["tid2phone"]),
                                   len(lat["tok_index"]), lattice_start(dec), len(src), _p(src), _p(dst), _p(arc), _p(ac), len(fs), _p(fs),
                                   _p(fc), lb, lm_scale, offset, int(nlsml), stage, int(phone_pass))
    s = C.string_at(ptr).decode()
    lib().orc_free(ptr)
    return s


def recognize(model, wave, frames_per_chunk=51, stages=False, rc=None, lattice=True, **over):
    """Whole reference path for one stream: int16 samples -> result text (one segment, no endpointing).

    lattice=True (the reference's only result path [REF src/batch_recognizer.cc:43-107,138-149]): raw lattice within
    lattice_beam -> phone-pruned determinization -> graph scale 0.9 -> word alignment -> MBR text.  lattice=False: the text of
    the best path (= the MBR result of a linear lattice, confidence 1), which is what the engine's lattice=0 mode returns."""
    ctx = model_context(model)
    feats = mfcc(wave)
    ends, avail, iv_index = chunk_plan(len(wave), frames_per_chunk, ctx)
    ivecs = ivectors(model, feats, ends, avail)
    ll = nnet_forward(model, feats, ivecs, iv_index)
    rc = rc or ResultCtx(model)
    if lattice and "lattice_beam" not in over:
        over["lattice_beam"] = float(model["conf"].get("lattice-beam", 6.0))
    dec = decode(model, ll, **over) if len(ll) else dict(best_arcs=np.zeros(0, dtype=np.int32))
    text_best = result_json(model, dec["best_arcs"], rc=rc)
    text = lattice_result(model, dec, over["lattice_beam"], rc=rc) if lattice and "lattice" in dec else text_best
    if stages:
        return dict(mfcc=feats, ivectors=ivecs, loglikes=ll, decode=dec, text=text, text_best=text_best, iv_index=iv_index)
    return text


# Kaldi OnlineEndpointConfig defaults, rules 1-4: (must_contain_nonsilence, min_trailing_silence, max_relative_cost, min_utterance_length)
ENDPOINT_RULES = [(False, 5.0, np.inf, 0.0), (True, 0.5, 2.0, 0.0), (True, 1.0, 8.0, 0.0), (True, 2.0, np.inf, 0.0)]


def endpoint_detected(model, dec, rules, silence_phones):
    """kaldi::EndpointDetected(config, tmodel, frame_shift, decoder) on a decoder that has consumed the log-likelihoods `dec`
    was run on [REF src/recognizer.cc:318]: TrailingSilenceLength (best path without final costs, emitting arcs of silence
    phones counted back from the newest frame), FinalRelativeCost, then rules 1-4 in single precision."""
    g = model["graph"]
    lo, hi = int(dec["offsets"][-2]), int(dec["offsets"][-1])
    if hi <= lo:
        return False
    cost = dec["cost"][lo:hi]
    i = lo + int(np.argmin(cost))
    fin = g["final"][dec["state"][lo:hi]].astype(np.float32)
    with_final = (cost + fin)[np.isfinite(fin)]
    relative = np.float32(with_final.min() - cost.min()) if len(with_final) else np.float32(np.inf)
    tid2phone = model["nnet"]["tid2phone"]
    sil = 0
    while i >= 0 and dec["arc"][i] >= 0:
        il = int(g["arc_ilabel"][dec["arc"][i]])
        if il != 0:
            if int(tid2phone[il]) in silence_phones:
                sil += 1
            else:
                break
        i = int(dec["prev"][i])
    shift = np.float32(0.03)
    utt, trail = np.float32(dec["frames"]) * shift, np.float32(sil) * shift
    nonsil = utt > trail
    for must, min_trail, max_rel, min_utt in rules:
        if (nonsil or not must) and trail >= np.float32(min_trail) and relative <= np.float32(min(max_rel, 1e30)) and utt >= np.float32(min_utt):
            return True
    return False


def recognize_segments(model, wave, frames_per_chunk=51, rule5_seconds=20.0, rc=None, silence_phones=None, rules=None, lattice=True, loglikes=None,
                       **over):
    """The batch path with reset_on_endpoint [REF src/batch_model.cc:72]: Kaldi's endpoint rules tested after every chunk; a
    segment is finalized there, the search starts over on the next chunk while features and i-vector carry on, and result
    times are offset by the segment start (GetTimeOffsetSeconds [REF src/batch_recognizer.cc:146-147]).  With an empty
    silence-phone list only rule 5 (decoded length >= 20 s) can fire; with silence_phones (a set of phone ids) rules 1-4 run
    too (`rules` defaults to Kaldi's).  lattice: result text through the lattice chain (default, as the reference) or from the
    best path.  loglikes: log-likelihoods to search instead of the oracle's own (a parity test passes the engine's, so that
    the texts can be compared exactly).  Returns the list of result texts."""
    ctx = model_context(model)
    ends, avail, iv_index = chunk_plan(len(wave), frames_per_chunk, ctx)
    if loglikes is None:
        feats = mfcc(wave)
        ivecs = ivectors(model, feats, ends, avail)
        ll = nnet_forward(model, feats, ivecs, iv_index)
    else:
        ll = np.ascontiguousarray(loglikes, dtype=np.float32)
    rc = rc or ResultCtx(model)
    rule5 = int(np.ceil(rule5_seconds / 0.03 - 1e-6)) if rule5_seconds > 0 else 0
    texts, seg_start = [], 0
    for k, a in enumerate(avail):
        last = k == len(avail) - 1
        dec = len(ll) if last else ((int(a) - ctx + 2) // 3 if a > ctx else 0)
        close = last or (rule5 and dec - seg_start >= rule5)
        d = None
        if not close and silence_phones and dec > seg_start:
            d = decode(model, ll[seg_start:dec], **over)
            close = endpoint_detected(model, d, rules or ENDPOINT_RULES, silence_phones)
        if close:
            seg = ll[seg_start:dec]
            lb = float(over.get("lattice_beam", model["conf"].get("lattice-beam", 6.0)))
            if len(seg) and (d is None or (lattice and "lattice" not in d)):
                d = decode(model, seg, **dict(over, lattice_beam=lb)) if lattice else decode(model, seg, **over)
            elif d is None:
                d = dict(best_arcs=np.zeros(0, dtype=np.int32))
            off = float(np.float32(seg_start * 0.03))
            if lattice and "lattice" in d:
                texts.append(lattice_result(model, d, lb, offset=off, rc=rc))
            else:
                texts.append(result_json(model, d["best_arcs"], offset=off, rc=rc))
            seg_start = dec
    return texts
