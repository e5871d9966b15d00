"""Synthetic Vosk model directories for the batch hot path (test/bench tooling).

Real Vosk models cannot be downloaded offline, so every model here is a
random-init network of the *named* architecture with a synthetic HCLG
(BASELINE.json north_star).  The directory layout and file names are the ones
the reference hard-codes in `BatchModel::BatchModel()`
[REF src/batch_model.cc:28-37,76-77] and `Model::ConfigureV2()`
[REF src/model.cc:176-206,251-256]:

    model/conf/{model,mfcc,ivector}.conf      --key=value lines (Kaldi ParseOptions style)
    model/am/final.mdl                        transition tables + collapsed TDNN-F parameters
    model/graph/HCLG.fst                      OpenFst *const* StdArc FST (binary, version 2)
    model/graph/words.txt                     "<word> <id>" per line
    model/graph/phones/word_boundary.int      "<phone-id> nonword|begin|end|internal|singleton"
    model/ivector/{final.ie,final.dubm,final.mat,global_cmvn.stats,splice.conf,online_cmvn.conf}

`HCLG.fst`, `words.txt`, `word_boundary.int` and the conf files use the real
on-disk formats.  `final.mdl` and the binary i-vector files use a small named
tensor container ("VBT1", documented in `write_vbt`) because Kaldi's own
nnet3/ivector binary formats cannot be produced or checked offline (SURVEY.md
§8f rank 2: real-format loaders come after the synthetic-format path).

The network architecture follows the in-tree recipe
[REF training/local/chain/run_tdnn.sh:98-129]; parameters are stored in the
form the reference uses at run time, i.e. after `CollapseModel`
[REF src/batch_model.cc:46-48]: idct + batchnorm0 + delta-layer are folded
into `tdnn1`, and the trailing batchnorm of prefinal-chain into its linear.
The un-collapsed pieces are kept under `raw.*` so a test can check the fold.
"""
from __future__ import annotations

import io
import os
import struct

import numpy as np

# --------------------------------------------------------------------------- #
# VBT1 container: magic "VBT1", u32 n, then n × {u32 name_len, name, u32 dtype,
# u32 ndim, u64 dims[ndim], raw little-endian data}.  dtype: 0=f32 1=i32 2=f64 3=u8
# --------------------------------------------------------------------------- #
_DT = {0: np.float32, 1: np.int32, 2: np.float64, 3: np.uint8}
_DT_INV = {np.dtype(v): k for k, v in _DT.items()}


def write_vbt(path, tensors: dict):
    with open(path, "wb") as f:
        f.write(b"VBT1")
        f.write(struct.pack("<I", len(tensors)))
        for name, arr in tensors.items():
            if isinstance(arr, str):
                arr = np.frombuffer(arr.encode(), dtype=np.uint8)
            arr = np.ascontiguousarray(arr)
            code = _DT_INV[arr.dtype]
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<II", code, arr.ndim))
            for d in arr.shape:
                f.write(struct.pack("<Q", d))
            f.write(arr.tobytes())


def read_vbt(path) -> dict:
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:4] == b"VBT1", path
    (n,) = struct.unpack_from("<I", buf, 4)
    p = 8
    for _ in range(n):
        (ln,) = struct.unpack_from("<I", buf, p); p += 4
        name = buf[p:p + ln].decode(); p += ln
        code, ndim = struct.unpack_from("<II", buf, p); p += 8
        dims = struct.unpack_from("<%dQ" % ndim, buf, p); p += 8 * ndim
        dt = np.dtype(_DT[code])
        cnt = int(np.prod(dims)) if ndim else 1
        arr = np.frombuffer(buf, dtype=dt, count=cnt, offset=p).reshape(dims)
        p += cnt * dt.itemsize
        out[name] = arr
    return out


def vbt_text(arr) -> str:
    return bytes(arr).decode()


# --------------------------------------------------------------------------- #
# Architectures
# --------------------------------------------------------------------------- #
ARCHS = {
    # in-tree recipe [REF training/local/chain/run_tdnn.sh:98-129]; N_pdf fixed at 2496 (<= 2500 leaves, :79)
    "small": dict(feat_dim=40, ivector_dim=40, hidden=512, bottleneck=96,
                  strides=[1, 1, 1, 0, 3, 3, 3, 3, 3, 3, 3],
                  prefinal_small=192, prefinal_big=512, num_pdfs=2496,
                  num_gauss=512, vocab=8000, succ=10),
    # en-us-0.22: architecture NOT in the reference; stated assumption = librispeech tdnn_1d style
    "large": dict(feat_dim=40, ivector_dim=100, hidden=1536, bottleneck=160,
                  strides=[1, 1, 1, 0, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3],
                  prefinal_small=256, prefinal_big=1536, num_pdfs=6016,
                  num_gauss=512, vocab=200000, succ=64),
    # tiny: unit-test sized (same topology family)
    "tiny": dict(feat_dim=40, ivector_dim=16, hidden=64, bottleneck=32,
                 strides=[1, 0, 3],
                 prefinal_small=32, prefinal_big=64, num_pdfs=96,
                 num_gauss=16, vocab=60, succ=4),
}

BYPASS_SCALE = 0.75  # [REF training/local/chain/run_tdnn.sh:90]
SUBSAMPLE = 3        # [REF src/batch_model.cc:82]


def context_of(arch):
    """(left,right) frames of model context: delta +-2, each tdnnf +-stride."""
    c = 2 + sum(arch["strides"])
    return c, c


# --------------------------------------------------------------------------- #
# MFCC (numpy, whole utterance) — generator-side twin used to fit BN0/UBM stats
# and as an independent cross-check in tests.  Options [REF training/conf/mfcc.conf:1-7].
# --------------------------------------------------------------------------- #
def mel_scale(f):
    return 1127.0 * np.log(1.0 + f / 700.0)


def mel_banks(num_bins=40, nfft=512, sr=16000.0, low=20.0, high=-400.0):
    nyq = sr / 2
    if high <= 0:
        high += nyq
    nb = nfft // 2
    bw = sr / nfft
    ml, mh = mel_scale(low), mel_scale(high)
    delta = (mh - ml) / (num_bins + 1)
    W = np.zeros((num_bins, nb), dtype=np.float64)
    mel = mel_scale(bw * np.arange(nb))
    for j in range(num_bins):
        l, c, r = ml + j * delta, ml + (j + 1) * delta, ml + (j + 2) * delta
        up = (mel - l) / (c - l)
        dn = (r - mel) / (r - c)
        w = np.where(mel <= c, up, dn)
        w[(mel <= l) | (mel >= r)] = 0.0
        W[j] = w
    return W.astype(np.float32)


def dct_matrix(n=40):
    m = np.zeros((n, n), dtype=np.float64)
    m[0, :] = np.sqrt(1.0 / n)
    k = np.arange(1, n)[:, None]
    j = np.arange(n)[None, :]
    m[1:, :] = np.sqrt(2.0 / n) * np.cos(np.pi / n * (j + 0.5) * k)
    return m


def lifter_coeffs(n=40, q=22.0):
    return 1.0 + 0.5 * q * np.sin(np.pi * np.arange(n) / q)


def povey_window(n=400):
    a = 2 * np.pi / (n - 1)
    return np.power(0.5 - 0.5 * np.cos(a * np.arange(n)), 0.85)


def mfcc_numpy(wave: np.ndarray) -> np.ndarray:
    """float64 whole-utterance MFCC; wave = int16-valued samples (not normalised)."""
    x = np.asarray(wave, dtype=np.float64)
    n = len(x)
    if n < 400:
        return np.zeros((0, 40), dtype=np.float32)
    nf = 1 + (n - 400) // 160
    idx = np.arange(400)[None, :] + 160 * np.arange(nf)[:, None]
    fr = x[idx]
    fr = fr - fr.mean(axis=1, keepdims=True)
    pre = np.empty_like(fr)
    pre[:, 1:] = fr[:, 1:] - 0.97 * fr[:, :-1]
    pre[:, 0] = fr[:, 0] - 0.97 * fr[:, 0]
    pre *= povey_window()[None, :]
    spec = np.fft.rfft(pre, n=512, axis=1)
    power = (spec.real ** 2 + spec.imag ** 2)[:, :256]
    mel = power @ mel_banks().astype(np.float64).T
    mel = np.log(np.maximum(mel, np.finfo(np.float32).eps))
    cep = mel @ dct_matrix().T
    cep *= lifter_coeffs()[None, :]
    return cep.astype(np.float32)


# --------------------------------------------------------------------------- #
# Synthetic speech-like audio: harmonics + syllabic envelope +
# formant-band noise, sigma~350, low-level noise floor (never digital silence).
# --------------------------------------------------------------------------- #
def synth_audio(seconds: float, seed: int, sr=16000) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sr))
    t = np.arange(n) / sr
    f0 = rng.uniform(90, 220) * (1.0 + 0.08 * np.sin(2 * np.pi * rng.uniform(0.3, 0.9) * t + rng.uniform(0, 6)))
    ph = 2 * np.pi * np.cumsum(f0) / sr
    sig = np.zeros(n)
    for h in range(1, 6):
        sig += rng.uniform(0.3, 1.0) / h * np.sin(h * ph + rng.uniform(0, 6))
    env = 0.55 + 0.45 * np.sin(2 * np.pi * 4.0 * t + rng.uniform(0, 6))
    env *= 0.6 + 0.4 * np.sin(2 * np.pi * rng.uniform(0.4, 1.1) * t + rng.uniform(0, 6))
    sig *= np.clip(env, 0.02, None)
    noise = rng.standard_normal(n)
    spec = np.fft.rfft(noise)
    fr = np.fft.rfftfreq(n, 1 / sr)
    shape = np.zeros_like(fr)
    for fc in (rng.uniform(400, 900), rng.uniform(1200, 2300), rng.uniform(2600, 3600)):
        shape += np.exp(-0.5 * ((fr - fc) / 180.0) ** 2)
    fn = np.fft.irfft(spec * shape, n=n)
    fn /= (fn.std() + 1e-9)
    sig = sig / (sig.std() + 1e-9) + 0.5 * fn * np.clip(env, 0.05, None)
    sig = 350.0 * sig / (sig.std() + 1e-9)
    edge = int(0.3 * sr)
    gate = np.ones(n)
    if n > 2 * edge:
        gate[:edge] = 0.0
        gate[-edge:] = 0.0
    sig = sig * gate + 5.0 * rng.standard_normal(n)
    return np.clip(np.round(sig), -4000, 4000).astype(np.int16)


# --------------------------------------------------------------------------- #
# Phones / transition tables
# --------------------------------------------------------------------------- #
N_BASE_PHONES = 41            # non-silence base phones
N_PHONES = 5 + 4 * N_BASE_PHONES  # ids 1..169 (0 = <eps>); 1..5 = SIL, SIL_B, SIL_E, SIL_I, SIL_S
POS_B, POS_E, POS_I, POS_S = 0, 1, 2, 3


def phone_id(base, pos):
    return 6 + 4 * (base - 1) + pos


def make_transition_tables(num_pdfs):
    """Chain topology: one emitting HMM state per phone with a forward pdf and a
    self-loop pdf; transition-state ts owns tids (2ts+1 = self-loop, 2ts+2 = forward),
    pdfs (2ts+1 = self-loop pdf, 2ts = forward pdf)."""
    n_ts = num_pdfs // 2
    base = max(1, n_ts // N_PHONES)
    nvar = np.full(N_PHONES + 1, base, dtype=np.int64)
    nvar[0] = 0
    extra = n_ts - base * N_PHONES
    if extra > 0:
        nvar[1:1 + extra] += 1
    elif extra < 0:  # fewer tstates than phones (tiny): share
        nvar[:] = 0
        nvar[1:] = 1
    ts_base = np.concatenate([[0], np.cumsum(nvar)[:-1]])
    if extra < 0:
        ts_base = np.arange(N_PHONES + 1) % n_ts
        ts_base[0] = 0
        nvar[1:] = 1
    ntid = 2 * n_ts
    tid2pdf = np.zeros(ntid + 1, dtype=np.int32)
    tid2phone = np.zeros(ntid + 1, dtype=np.int32)
    tid2pdf[0] = -1
    ts = np.arange(n_ts)
    tid2pdf[2 * ts + 1] = 2 * ts + 1
    tid2pdf[2 * ts + 2] = 2 * ts
    # owner phone of each tstate
    owner = np.zeros(n_ts, dtype=np.int32)
    if extra >= 0:
        owner = np.repeat(np.arange(N_PHONES + 1), nvar).astype(np.int32)
    # for the shared (tiny) case the phone of a tid is ambiguous; word alignment then
    # uses the first phone mapped to it, which tests for 'tiny' do not rely on.
    else:
        for ph in range(N_PHONES, 0, -1):
            owner[ts_base[ph]] = ph
    tid2phone[2 * ts + 1] = owner
    tid2phone[2 * ts + 2] = owner
    return nvar, ts_base, tid2pdf, tid2phone


def tstate_of(phone, prev_phone, nvar, ts_base):
    """Context-dependent tstate: variant chosen by the previous phone."""
    return ts_base[phone] + (prev_phone * 7 + 3) % np.maximum(nvar[phone], 1)


# --------------------------------------------------------------------------- #
# OpenFst const FST writer/reader (StdArc, version 2 = unaligned) — SURVEY.md A11
# --------------------------------------------------------------------------- #
FST_MAGIC = 2125659606
_STATE_DT = np.dtype([("final", "<f4"), ("pos", "<u4"), ("narcs", "<u4"), ("nieps", "<u4"), ("noeps", "<u4")])
_ARC_DT = np.dtype([("ilabel", "<i4"), ("olabel", "<i4"), ("weight", "<f4"), ("next", "<i4")])


def write_const_fst(path, start, final, src, ilabel, olabel, weight, dst):
    ns = len(final)
    order = np.argsort(src, kind="stable")
    src, ilabel, olabel, weight, dst = (a[order] for a in (src, ilabel, olabel, weight, dst))
    narcs = np.bincount(src, minlength=ns).astype(np.uint32)
    pos = np.concatenate([[0], np.cumsum(narcs)[:-1]]).astype(np.uint32)
    states = np.zeros(ns, dtype=_STATE_DT)
    states["final"] = final
    states["pos"] = pos
    states["narcs"] = narcs
    states["nieps"] = np.bincount(src[ilabel == 0], minlength=ns)
    states["noeps"] = np.bincount(src[olabel == 0], minlength=ns)
    arcs = np.zeros(len(src), dtype=_ARC_DT)
    arcs["ilabel"], arcs["olabel"], arcs["weight"], arcs["next"] = ilabel, olabel, weight, dst
    with open(path, "wb") as f:
        f.write(struct.pack("<i", FST_MAGIC))
        for s in (b"const", b"standard"):
            f.write(struct.pack("<i", len(s)))
            f.write(s)
        f.write(struct.pack("<ii", 2, 0))           # version 2 (unaligned), flags 0 (no symbol tables)
        f.write(struct.pack("<Q", 0x0000000000000003))  # properties: expanded|mutable bits only (informational)
        f.write(struct.pack("<qqq", start, ns, len(src)))
        f.write(states.tobytes())
        f.write(arcs.tobytes())


def read_fst(path):
    """Returns dict(start, final[ns], row[ns+1], ilabel, olabel, weight, next) in FILE arc order.
    Handles const (v1 aligned / v2) and vector FSTs."""
    buf = open(path, "rb").read()
    p = 0
    (magic,) = struct.unpack_from("<i", buf, p); p += 4
    assert magic == FST_MAGIC
    (ln,) = struct.unpack_from("<i", buf, p); p += 4
    ftype = buf[p:p + ln].decode(); p += ln
    (ln,) = struct.unpack_from("<i", buf, p); p += 4
    atype = buf[p:p + ln].decode(); p += ln
    assert atype == "standard"
    version, flags = struct.unpack_from("<ii", buf, p); p += 8
    p += 8  # properties
    start, ns, na = struct.unpack_from("<qqq", buf, p); p += 24
    assert (flags & 3) == 0, "embedded symbol tables not supported by this reader"
    if ftype == "const":
        if version == 1 and p % 16:
            p += 16 - p % 16
        st = np.frombuffer(buf, dtype=_STATE_DT, count=ns, offset=p); p += ns * _STATE_DT.itemsize
        if version == 1 and p % 16:
            p += 16 - p % 16
        arcs = np.frombuffer(buf, dtype=_ARC_DT, count=na, offset=p)
        row = np.concatenate([st["pos"].astype(np.int64), [na]])
        final = st["final"].copy()
    elif ftype == "vector":
        final = np.zeros(ns, dtype=np.float32)
        row = np.zeros(ns + 1, dtype=np.int64)
        chunks = []
        for s in range(ns):
            (fw,) = struct.unpack_from("<f", buf, p); p += 4
            (n,) = struct.unpack_from("<q", buf, p); p += 8
            final[s] = fw
            row[s + 1] = row[s] + n
            chunks.append(np.frombuffer(buf, dtype=_ARC_DT, count=n, offset=p)); p += 16 * n
        arcs = np.concatenate(chunks) if chunks else np.zeros(0, dtype=_ARC_DT)
    else:
        raise ValueError("unsupported fst type " + ftype)
    return dict(start=int(start), final=final, row=row,
                ilabel=arcs["ilabel"].copy(), olabel=arcs["olabel"].copy(),
                weight=arcs["weight"].copy(), next=arcs["next"].copy())


def csr_canonical(fst, tid2pdf):
    """Canonical CSR arc numbering used by BOTH the oracle and the engine (it defines
    the tie-break of equal-cost arrivals): states in id order; within a state the emitting
    arcs (ilabel != 0) in file order, then the epsilon-input arcs in file order."""
    row = fst["row"]
    ns = len(row) - 1
    na = int(row[-1])
    src = np.repeat(np.arange(ns, dtype=np.int64), np.diff(row))
    is_eps = (fst["ilabel"] == 0).astype(np.int64)
    order = np.lexsort((np.arange(na), is_eps, src))
    il, ol, w, nx = (fst[k][order] for k in ("ilabel", "olabel", "weight", "next"))
    n_all = np.diff(row)
    n_eps = np.bincount(src[fst["ilabel"] == 0], minlength=ns)
    e_begin = row[:-1].astype(np.int32)
    eps_begin = (row[:-1] + n_all - n_eps).astype(np.int32)
    pdf = np.where(il > 0, tid2pdf[np.maximum(il, 0)], -1).astype(np.int32)
    return dict(num_states=ns, num_arcs=na, start=fst["start"],
                final=fst["final"].astype(np.float32),
                e_begin=np.concatenate([e_begin, [na]]).astype(np.int32),
                eps_begin=eps_begin,
                arc_w=w.astype(np.float32), arc_next=nx.astype(np.int32),
                arc_pdf=pdf, arc_ilabel=il.astype(np.int32), arc_olabel=ol.astype(np.int32),
                arc_src=src[order].astype(np.int32))


# --------------------------------------------------------------------------- #
# Lexicon, bigram G and the HCLG-shaped decoding graph
# --------------------------------------------------------------------------- #
_ONSETS = ["b", "d", "f", "g", "h", "k", "l", "m", "n", "p", "r", "s", "t", "v", "z", "ch", "sh", "th", "st", "tr"]
_VOWELS = ["a", "e", "i", "o", "u", "ai", "ou", "ee"]


def make_words(vocab):
    words = []
    no, nv = len(_ONSETS), len(_VOWELS)
    for i in range(vocab):
        k, s = i, ""
        for _ in range(2 + (i >= no * nv) + (i >= (no * nv) ** 2)):
            s += _ONSETS[k % no] + _VOWELS[(k // no) % nv]
            k //= no * nv
        words.append(s)
    assert len(set(words)) == vocab
    return words


def make_lexicon(vocab, rng):
    """Unique base-phone sequences, returned position-dependent, 0-padded [vocab, Lmax]."""
    lmax = 9
    need = vocab
    seqs = np.zeros((0, lmax), dtype=np.int64)
    while len(seqs) < need:
        m = int((need - len(seqs)) * 1.3) + 64
        ln = np.clip(np.round(rng.normal(5.2, 1.7, m)), 1, lmax).astype(np.int64)
        ph = rng.integers(1, N_BASE_PHONES + 1, size=(m, lmax))
        ph[np.arange(lmax)[None, :] >= ln[:, None]] = 0
        seqs = np.concatenate([seqs, ph])
        _, first = np.unique(seqs, axis=0, return_index=True)
        seqs = seqs[np.sort(first)]
    seqs = seqs[:need]
    ln = (seqs > 0).sum(1)
    pos = np.full(seqs.shape, POS_I)
    pos[:, 0] = POS_B
    pos[np.arange(need), ln - 1] = POS_E
    pos[ln == 1, 0] = POS_S
    pd = np.where(seqs > 0, phone_id(seqs, pos), 0)
    return pd, ln


def make_graph(arch, seed):
    """Builds an HCLG-shaped graph directly: a back-off bigram G whose per-history word
    sets are expanded into phone prefix trees of 1-state chain HMMs (forward tid on the
    arc into a phone state, self-loop tid on it), tropical-pushed LM costs, word label on
    the arc where the word becomes unique, epsilon arcs for word-end -> history hub,
    history hub -> unigram hub (back-off) and silence return."""
    rng = np.random.default_rng(seed)
    V, K = arch["vocab"], arch["succ"]
    nvar, ts_base, tid2pdf, tid2phone = make_transition_tables(arch["num_pdfs"])
    lex, lex_len = make_lexicon(V, rng)              # words 1..V -> lex[w-1]
    # unigram (Zipf) and bigram successors
    rank = rng.permutation(V)
    uni = 1.0 / (rank + 10.0)
    uni /= uni.sum()
    uni_cost = -np.log(uni)
    succ = rng.choice(V, size=(V, K), p=uni) + 1      # may repeat inside a row; dedupe below
    succ.sort(axis=1)
    dup = np.zeros_like(succ, dtype=bool)
    dup[:, 1:] = succ[:, 1:] == succ[:, :-1]
    q = rng.uniform(0.2, 1.0, size=(V, K)) * uni[succ - 1]
    q[dup] = 0.0
    q /= q.sum(1, keepdims=True)
    big_cost = -np.log(0.7 * np.maximum(q, 1e-30))
    backoff_cost = -np.log(rng.uniform(0.2, 0.4, size=V))
    # rows = (hub, word, cost); hub 0 = unigram hub U, hub h = history of word h
    hub_u = np.zeros(V, dtype=np.int64)
    row_hub = np.concatenate([hub_u, np.repeat(np.arange(1, V + 1), K)[~dup.ravel()]])
    row_word = np.concatenate([np.arange(1, V + 1), succ.ravel()[~dup.ravel()]])
    row_cost = np.concatenate([uni_cost, big_cost.ravel()[~dup.ravel()]])
    P = lex[row_word - 1]
    L = lex_len[row_word - 1]
    lmax = P.shape[1]
    order = np.lexsort(tuple(P[:, d] for d in range(lmax - 1, -1, -1)) + (row_hub,))
    row_hub, row_word, row_cost, P, L = row_hub[order], row_word[order], row_cost[order], P[order], L[order]
    R = len(row_hub)
    valid = np.arange(lmax)[None, :] < L[:, None]
    neq = np.ones((R, lmax), dtype=bool)
    neq[1:] = P[1:] != P[:-1]
    hub_change = np.ones(R, dtype=bool)
    hub_change[1:] = row_hub[1:] != row_hub[:-1]
    new = (np.logical_or.accumulate(neq, axis=1) | hub_change[:, None]) & valid
    flat_id = np.cumsum(new.ravel()).reshape(R, lmax) - 1
    n_nodes = int(new.sum())
    node = np.zeros((R, lmax), dtype=np.int64)
    ar = np.arange(R)
    for d in range(lmax):
        last_new = np.maximum.accumulate(np.where(new[:, d], ar, 0))
        node[:, d] = flat_id[last_new, d]
    node[~valid] = -1
    cnt = np.bincount(node[valid], minlength=n_nodes)
    node_min = np.full(n_nodes, np.inf)
    rr = np.broadcast_to(ar[:, None], (R, lmax))
    np.minimum.at(node_min, node[valid], row_cost[rr[valid]])
    dstar = (np.where(valid, cnt[np.maximum(node, 0)], 0) > 1).sum(1)
    assert np.all(dstar < L), "duplicate pronunciation inside one hub"
    # state numbering
    S0, U = 0, 1
    hub_state = np.concatenate([[U], np.arange(2, V + 2)])
    node_state = V + 2 + np.arange(n_nodes)
    sil_state = V + 2 + n_nodes
    ns = sil_state + 1
    LOG2 = float(-np.log(0.5))
    # tree arcs: one per node, defined at its 'new' (row, depth)
    ri, di = np.nonzero(new)
    nid = node[ri, di]
    parent_state = np.where(di == 0, hub_state[row_hub[ri]], node_state[node[ri, np.maximum(di - 1, 0)]])
    parent_min = np.where(di == 0, 0.0, node_min[node[ri, np.maximum(di - 1, 0)]])
    prev_ph = np.where(di == 0, 0, P[ri, np.maximum(di - 1, 0)])
    ts = tstate_of(P[ri, di], prev_ph, nvar, ts_base)
    a_src = [parent_state]
    a_il = [2 * ts + 2]
    a_ol = [np.where(di == dstar[ri], row_word[ri], 0)]
    a_w = [LOG2 + (node_min[nid] - parent_min)]
    a_dst = [node_state[nid]]
    # self loops
    a_src.append(node_state[nid]); a_il.append(2 * ts + 1); a_ol.append(np.zeros_like(nid))
    a_w.append(np.full(len(nid), LOG2)); a_dst.append(node_state[nid])
    # word end -> history hub (eps)
    leaf = node[ar, L - 1]
    a_src.append(node_state[leaf]); a_il.append(np.zeros(R, dtype=np.int64)); a_ol.append(np.zeros(R, dtype=np.int64))
    a_w.append(np.zeros(R)); a_dst.append(hub_state[row_word])
    # history hub -> unigram hub (back-off eps)
    hs = np.arange(1, V + 1)
    a_src.append(hub_state[hs]); a_il.append(np.zeros(V, dtype=np.int64)); a_ol.append(np.zeros(V, dtype=np.int64))
    a_w.append(backoff_cost); a_dst.append(np.full(V, U))
    # start -> U (eps), U -> SIL (emitting), SIL self loop, SIL -> U (eps)
    sil_ts = ts_base[1]
    a_src.append(np.array([S0, U, sil_state, sil_state]))
    a_il.append(np.array([0, 2 * sil_ts + 2, 2 * sil_ts + 1, 0]))
    a_ol.append(np.zeros(4, dtype=np.int64))
    a_w.append(np.array([0.0, LOG2 + 1.0, LOG2, 0.0]))
    a_dst.append(np.array([U, sil_state, sil_state, U]))
    src = np.concatenate(a_src).astype(np.int64)
    il = np.concatenate(a_il).astype(np.int32)
    ol = np.concatenate(a_ol).astype(np.int32)
    w = np.concatenate(a_w).astype(np.float32)
    dst = np.concatenate(a_dst).astype(np.int32)
    final = np.full(ns, np.inf, dtype=np.float32)
    final[hub_state] = 0.0
    return dict(start=S0, final=final, src=src, ilabel=il, olabel=ol, weight=w, dst=dst,
                tid2pdf=tid2pdf, tid2phone=tid2phone, words=make_words(V))


# --------------------------------------------------------------------------- #
# Acoustic model parameters
# --------------------------------------------------------------------------- #
def nnet_layers(arch):
    """Op list of the collapsed network: (name, in_node, offsets, relu_bn, bypass_node, uses_ivec)."""
    ops = [("tdnn1", 0, [-2, -1, 0, 1, 2], True, -1, True)]
    cur = 1
    for k, s in enumerate(arch["strides"], start=2):
        lo = [-s, 0] if s else [0]
        ao = [0, s] if s else [0]
        ops.append((f"tdnnf{k}.linear", cur, lo, False, -1, False))
        ops.append((f"tdnnf{k}.affine", cur + 1, ao, True, cur, False))
        cur += 2
    ops.append(("prefinal_l", cur, [0], False, -1, False))
    ops.append(("prefinal.affine", cur + 1, [0], True, -1, False))
    ops.append(("prefinal.linear", cur + 2, [0], False, -1, False))
    ops.append(("output", cur + 3, [0], False, -1, False))
    return ops


def _bn_name(op_name):
    return op_name.replace(".affine", "")


def nnet_forward_numpy(T, arch, mfcc, ivecs, iv_index, calibrate_rng=None):
    """float64 whole-utterance forward of the collapsed network over ALL frames (independent check of
    the oracle; also used by the generator to fit the test-mode batchnorm statistics).
    Returns loglikes at t = 0,3,6,...  T is the tensor dict (modified in place when calibrating)."""
    ctx = context_of(arch)[0]
    n = len(mfcc)
    Lt = n + 2 * ctx
    idx = np.clip(np.arange(Lt) - ctx, 0, n - 1)
    acts = [mfcc.astype(np.float64)[idx]]
    iv_rows = np.zeros(Lt, dtype=np.int64)
    iv_rows[2:Lt - 2] = iv_index
    ar = np.arange(Lt)
    for name, src, offs, relu_bn, byp, uses_iv in nnet_layers(arch):
        x = acts[src]
        cols = [x[np.clip(ar + o, 0, Lt - 1)] for o in offs]
        if uses_iv:
            cols.append(np.asarray(ivecs, dtype=np.float64)[iv_rows])
        X = np.concatenate(cols, axis=1)
        z = X @ T[name + ".w"].astype(np.float64).T
        if name + ".b" in T:
            z = z + T[name + ".b"].astype(np.float64)
        if relu_bn:
            z = np.maximum(z, 0.0)
            bn = _bn_name(name)
            if calibrate_rng is not None:
                core = z[ctx:Lt - ctx]
                sc = 1.0 / np.sqrt(core.var(0) + 1e-3) * calibrate_rng.uniform(0.9, 1.1, z.shape[1])
                T[bn + ".bn_scale"] = sc.astype(np.float32)
                T[bn + ".bn_offset"] = (-core.mean(0) * sc + 0.05 * calibrate_rng.standard_normal(z.shape[1])).astype(np.float32)
            z = z * T[bn + ".bn_scale"].astype(np.float64) + T[bn + ".bn_offset"].astype(np.float64)
        if byp >= 0:
            z = z + BYPASS_SCALE * acts[byp]
        acts.append(z)
    return acts[-1][ctx:Lt - ctx:3]


def make_nnet(arch, seed, feat_stats, calib_feats=None):
    """feat_stats = (mean[40], std[40]) of idct(mfcc) on synthetic audio, used for batchnorm0.
    calib_feats: MFCCs on which the test-mode batchnorm statistics are fitted (as training would)."""
    rng = np.random.default_rng(seed)
    F, I, H, B = arch["feat_dim"], arch["ivector_dim"], arch["hidden"], arch["bottleneck"]
    T = {}

    def w(o, i, gain=1.0):
        return (rng.standard_normal((o, i)) * gain / np.sqrt(i))

    def bn(n):
        return np.ones(n), np.zeros(n)

    # --- raw (un-collapsed) front end ---
    idct = np.linalg.inv(np.diag(lifter_coeffs(F)) @ dct_matrix(F))   # [REF run_tdnn.sh:103] cepstral-lifter=22
    mean, std = feat_stats
    bn0_s = 1.0 / np.maximum(std, 1e-3)
    bn0_o = -mean * bn0_s
    w1 = w(H, 3 * F + I)                                               # over Append(delta(3F), ivector)
    w1[:, F:2 * F] *= 0.7                                              # keep the delta terms at unit scale
    w1[:, 2 * F:3 * F] *= 0.4
    b1 = 0.1 * rng.standard_normal(H)
    T["raw.idct"] = idct
    T["raw.bn0_scale"], T["raw.bn0_offset"] = bn0_s, bn0_o
    T["raw.tdnn1.w"], T["raw.tdnn1.b"] = w1, b1
    # --- collapse: y(t) = A0 mfcc(t) + c0 ; delta = [y, y(t+1)-y(t-1), y(t-2)-2y(t)+y(t+2)] ---
    A0 = bn0_s[:, None] * idct
    c0 = bn0_o
    Wy, Wd, Wdd, Wi = w1[:, :F], w1[:, F:2 * F], w1[:, 2 * F:3 * F], w1[:, 3 * F:]
    blocks = [Wdd @ A0, -(Wd @ A0), (Wy - 2.0 * Wdd) @ A0, Wd @ A0, Wdd @ A0]   # offsets -2..+2
    T["tdnn1.w"] = np.concatenate(blocks + [Wi], axis=1)               # [H, 5F+I]
    # constant term: Wy c0 + Wd (c0-c0) + Wdd (c0 - 2c0 + c0) = Wy c0
    T["tdnn1.b"] = b1 + Wy @ c0
    T["tdnn1.bn_scale"], T["tdnn1.bn_offset"] = bn(H)
    for k, s in enumerate(arch["strides"], start=2):
        nin = 2 if s > 0 else 1
        T[f"tdnnf{k}.linear.w"] = w(B, nin * H)
        T[f"tdnnf{k}.affine.w"] = w(H, nin * B, 1.2)
        T[f"tdnnf{k}.affine.b"] = 0.1 * rng.standard_normal(H)
        T[f"tdnnf{k}.bn_scale"], T[f"tdnnf{k}.bn_offset"] = bn(H)
    PS, PB, NP = arch["prefinal_small"], arch["prefinal_big"], arch["num_pdfs"]
    T["prefinal_l.w"] = w(PS, H, 0.8)
    T["prefinal.affine.w"] = w(PB, PS, 1.2)
    T["prefinal.affine.b"] = 0.1 * rng.standard_normal(PB)
    T["prefinal.bn_scale"], T["prefinal.bn_offset"] = bn(PB)
    lin = w(PS, PB)
    s2 = rng.uniform(0.9, 1.1, PS)
    o2 = 0.05 * rng.standard_normal(PS)
    T["raw.prefinal.linear.w"], T["raw.prefinal.bn2_scale"], T["raw.prefinal.bn2_offset"] = lin, s2, o2
    T["prefinal.linear.w"] = s2[:, None] * lin
    T["prefinal.linear.b"] = o2
    T["output.w"] = w(NP, PS, 3.0)      # include-log-softmax=false [REF run_tdnn.sh:125]: raw affine outputs
    T["output.b"] = 0.5 * rng.standard_normal(NP)
    T = {k: np.asarray(v, dtype=np.float32) for k, v in T.items()}
    if calib_feats is not None:
        ctx = context_of(arch)[0]
        n = len(calib_feats)
        iv = 0.5 * rng.standard_normal((4, I))
        iv_index = (np.arange(n + 2 * (ctx - 2)) * 4 // (n + 2 * (ctx - 2))).astype(np.int64)
        nnet_forward_numpy(T, arch, calib_feats, iv, iv_index, calibrate_rng=rng)
    return T


def nnet_config_text(arch):
    return "\n".join([
        "arch tdnnf",
        f"feat-dim {arch['feat_dim']}", f"ivector-dim {arch['ivector_dim']}",
        f"hidden-dim {arch['hidden']}", f"bottleneck-dim {arch['bottleneck']}",
        "tdnnf-strides " + " ".join(str(s) for s in arch["strides"]),
        f"prefinal-small {arch['prefinal_small']}", f"prefinal-big {arch['prefinal_big']}",
        f"num-pdfs {arch['num_pdfs']}", f"bypass-scale {BYPASS_SCALE}",
        f"frame-subsampling-factor {SUBSAMPLE}", ""])


# --------------------------------------------------------------------------- #
# i-vector extractor
# --------------------------------------------------------------------------- #
def splice_frames(x, left=3, right=3):
    T = len(x)
    idx = np.clip(np.arange(T)[:, None] + np.arange(-left, right + 1)[None, :], 0, T - 1)
    return x[idx].reshape(T, -1)


def make_ivector_extractor(arch, seed, feats):
    """feats: [T,40] sample MFCCs from synthetic audio used to place the UBM."""
    rng = np.random.default_rng(seed)
    F, D, G = arch["feat_dim"], arch["ivector_dim"], arch["num_gauss"]
    feats = feats.astype(np.float64)
    T = len(feats)
    gmean = feats.mean(0)
    stats = np.zeros((2, F + 1))
    stats[0, :F] = feats.sum(0); stats[0, F] = T
    stats[1, :F] = (feats ** 2).sum(0)
    lda = rng.standard_normal((F, 7 * F)) / np.sqrt(7.0)
    lda /= np.maximum(np.tile(feats.std(0), 7), 1e-3)[None, :]
    lda_off = -lda @ np.tile(gmean, 7) * 0.5
    lda_mat = np.concatenate([lda, lda_off[:, None]], axis=1)          # [F, 7F+1], last col = offset
    xn = splice_frames(feats - gmean) @ lda.T + lda_off
    xu = splice_frames(feats) @ lda.T + lda_off
    pick = rng.choice(T, size=G, replace=T < G)
    means = xn[pick] + 0.3 * rng.standard_normal((G, F)) * xn.std(0)
    var = np.tile(xn.var(0) * 1.5 + 1e-3, (G, 1)) * rng.uniform(0.7, 1.3, (G, F))
    wts = rng.dirichlet(np.full(G, 5.0))
    inv_var = 1.0 / var
    gconst = np.log(wts) - 0.5 * (F * np.log(2 * np.pi) + np.log(var).sum(1) + (means ** 2 * inv_var).sum(1))
    dubm = dict(gconsts=gconst, means_invvars=means * inv_var, inv_vars=inv_var, weights=wts)
    prior_offset = 8.0
    uvar = xu.var(0) + 1e-3
    umean = xu[pick] + 0.3 * rng.standard_normal((G, F)) * xu.std(0)
    M = 0.25 * rng.standard_normal((G, F, D)) * np.sqrt(uvar)[None, :, None]
    M[:, :, 0] = umean / prior_offset
    S = rng.standard_normal((G, F, F)) * 0.03
    S = 0.5 * (S + S.transpose(0, 2, 1))
    dh = (1.0 / np.sqrt(uvar * 1.5))
    sigma_inv = dh[None, :, None] * (np.eye(F)[None] + S) * dh[None, None, :]
    ie = dict(M=M, sigma_inv=sigma_inv, w=wts, prior_offset=np.array([prior_offset]))
    return dict(lda=lda_mat, dubm=dubm, ie=ie, cmvn=stats)


# --------------------------------------------------------------------------- #
# Model directory
# --------------------------------------------------------------------------- #
MODEL_CONF = """--min-active=200
--max-active=7000
--beam=13.0
--lattice-beam=6.0
--acoustic-scale=1.0
--frame-subsampling-factor=3
--endpoint.silence-phones=
"""
MFCC_CONF = """--use-energy=false
--num-mel-bins=40
--num-ceps=40
--low-freq=20
--high-freq=-400
--allow-upsample=true
--allow-downsample=true
--dither=0
"""
IVECTOR_CONF = """--splice-config=model/ivector/splice.conf
--cmvn-config=model/ivector/online_cmvn.conf
--lda-matrix=model/ivector/final.mat
--global-cmvn-stats=model/ivector/global_cmvn.stats
--diag-ubm=model/ivector/final.dubm
--ivector-extractor=model/ivector/final.ie
--num-gselect=5
--min-post=0.025
--posterior-scale=0.1
--max-remembered-frames=1000
--max-count=100
--ivector-period=10
"""


def write_model_dir(root, arch_name="small", seed=0, graph_seed=1, overrides=None, model_conf_extra=""):
    """Writes <root>/model/... ; returns the path of the model directory."""
    arch = dict(ARCHS[arch_name])
    if overrides:
        arch.update(overrides)
    m = os.path.join(root, "model")
    for d in ("conf", "am", "graph/phones", "ivector"):
        os.makedirs(os.path.join(m, d), exist_ok=True)
    open(os.path.join(m, "conf/model.conf"), "w").write(MODEL_CONF + model_conf_extra)
    open(os.path.join(m, "conf/mfcc.conf"), "w").write(MFCC_CONF)
    open(os.path.join(m, "conf/ivector.conf"), "w").write(IVECTOR_CONF)
    open(os.path.join(m, "ivector/splice.conf"), "w").write("--left-context=3\n--right-context=3\n")
    open(os.path.join(m, "ivector/online_cmvn.conf"), "w").write(
        "--cmn-window=600\n--global-frames=200\n--speaker-frames=600\n--norm-vars=false\n")
    # sample features for BN0 / UBM placement
    feats = np.concatenate([mfcc_numpy(synth_audio(6.0, 9000 + i)) for i in range(4)])
    logmel = (feats.astype(np.float64) / lifter_coeffs()[None, :]) @ dct_matrix()
    nnet = make_nnet(arch, seed, (logmel.mean(0), logmel.std(0)), calib_feats=feats)
    g = make_graph(arch, graph_seed)
    nnet["tid2pdf"] = g["tid2pdf"]
    nnet["tid2phone"] = g["tid2phone"]
    tensors = {"config": np.frombuffer(nnet_config_text(arch).encode(), dtype=np.uint8)}
    tensors.update(nnet)
    write_vbt(os.path.join(m, "am/final.mdl"), tensors)
    write_const_fst(os.path.join(m, "graph/HCLG.fst"), g["start"], g["final"], g["src"],
                    g["ilabel"], g["olabel"], g["weight"], g["dst"])
    with open(os.path.join(m, "graph/words.txt"), "w") as f:
        f.write("<eps> 0\n")
        for i, wd in enumerate(g["words"], start=1):
            f.write(f"{wd} {i}\n")
    with open(os.path.join(m, "graph/phones/word_boundary.int"), "w") as f:
        kinds = {POS_B: "begin", POS_E: "end", POS_I: "internal", POS_S: "singleton"}
        f.write("1 nonword\n2 begin\n3 end\n4 internal\n5 singleton\n")
        for b in range(1, N_BASE_PHONES + 1):
            for pos in (POS_B, POS_E, POS_I, POS_S):
                f.write(f"{phone_id(b, pos)} {kinds[pos]}\n")
    iv = make_ivector_extractor(arch, seed + 17, feats)
    f32 = lambda d: {k: np.asarray(v, dtype=np.float32) for k, v in d.items()}
    write_vbt(os.path.join(m, "ivector/final.mat"), f32({"lda": iv["lda"]}))
    write_vbt(os.path.join(m, "ivector/final.dubm"), f32(iv["dubm"]))
    write_vbt(os.path.join(m, "ivector/final.ie"), f32(iv["ie"]))
    write_vbt(os.path.join(m, "ivector/global_cmvn.stats"), {"stats": iv["cmvn"].astype(np.float64)})
    return m


def parse_conf(path):
    out = {}
    if not os.path.exists(path):
        return out
    for line in open(path):
        line = line.split("#")[0].strip()
        if not line.startswith("--"):
            continue
        k, _, v = line[2:].partition("=")
        out[k.strip()] = v.strip()
    return out


def load_model_dir(m):
    """Loads everything the oracle needs as numpy arrays (test infrastructure)."""
    mdl = read_vbt(os.path.join(m, "am/final.mdl"))
    cfg = {}
    for line in vbt_text(mdl["config"]).splitlines():
        if line.strip():
            k, _, v = line.partition(" ")
            cfg[k] = v
    fst = read_fst(os.path.join(m, "graph/HCLG.fst"))
    graph = csr_canonical(fst, mdl["tid2pdf"])
    words = {}
    for line in open(os.path.join(m, "graph/words.txt")):
        wd, i = line.split()
        words[int(i)] = wd
    wb = {}
    for line in open(os.path.join(m, "graph/phones/word_boundary.int")):
        p, k = line.split()
        wb[int(p)] = k
    iv = dict(lda=read_vbt(os.path.join(m, "ivector/final.mat"))["lda"],
              dubm=read_vbt(os.path.join(m, "ivector/final.dubm")),
              ie=read_vbt(os.path.join(m, "ivector/final.ie")),
              cmvn=read_vbt(os.path.join(m, "ivector/global_cmvn.stats"))["stats"])
    conf = parse_conf(os.path.join(m, "conf/model.conf"))
    return dict(cfg=cfg, nnet=mdl, graph=graph, words=words, word_boundary=wb, ivector=iv, conf=conf)


if __name__ == "__main__":
    import argparse, time
    ap = argparse.ArgumentParser()
    ap.add_argument("root")
    ap.add_argument("--arch", default="small")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    t0 = time.time()
    p = write_model_dir(a.root, a.arch, a.seed)
    sz = os.path.getsize(os.path.join(p, "graph/HCLG.fst"))
    print(f"wrote {p} in {time.time() - t0:.1f}s, HCLG {sz / 1e6:.1f} MB")
