"""Generates the committed golden fixtures (run in the build container, where torchaudio and /root/reference exist).

  mfcc_torchaudio.npz : torchaudio.compliance.kaldi.mfcc (a third-party restatement of Kaldi compute-mfcc-feats,
                        options of [REF training/conf/mfcc.conf:1-7], dither 0) on a seeded synthetic waveform
  json_ref.json       : result texts produced by the REFERENCE's own src/json.h (oracle/_ref/libref_json.so)
                        for result objects assembled as in [REF src/batch_recognizer.cc:82-105]
"""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "vosk-api_b200", "tools"))
import vbmodel  # noqa: E402


def main():
    import torch
    import torchaudio
    wave = vbmodel.synth_audio(1.5, 4242)
    ref = torchaudio.compliance.kaldi.mfcc(torch.from_numpy(wave.astype(np.float32))[None], dither=0.0, num_ceps=40, num_mel_bins=40,
                                           low_freq=20, high_freq=-400, energy_floor=0.0, use_energy=False, sample_frequency=16000).numpy()
    np.savez_compressed(os.path.join(HERE, "mfcc_torchaudio.npz"), wave=wave, mfcc=ref.astype(np.float32))
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_json.so"))
    lib.ref_json_result.restype = ctypes.c_void_p
    cases = [
        dict(words=[], start=[], end=[], conf=[], text=""),
        dict(words=["one"], start=[0.84], end=[1.11], conf=[1.0], text="one"),
        dict(words=["one", "zero", "zero"], start=[0.0, 0.33, 1.02], end=[0.33, 1.02, 2.97], conf=[1.0, 0.5, 0.123456789], text="one zero zero"),
        dict(words=['qu"ote', "back\\slash", "tab\there"], start=[0.03, 0.06, 0.09], end=[0.06, 0.09, 0.12], conf=[1, 1, 1], text='qu"ote back\\slash tab\there'),
    ]
    out = []
    for c in cases:
        n = len(c["words"])
        arr = (ctypes.c_char_p * max(n, 1))(*[w.encode() for w in c["words"]])
        dbl = lambda v: (ctypes.c_double * max(n, 1))(*v)
        p = lib.ref_json_result(n, arr, dbl(c["start"]), dbl(c["end"]), dbl(c["conf"]), c["text"].encode())
        c = dict(c, dump=ctypes.string_at(p).decode())
        out.append(c)
    json.dump(out, open(os.path.join(HERE, "json_ref.json"), "w"), indent=1)
    print("wrote fixtures")


if __name__ == "__main__":
    main()
