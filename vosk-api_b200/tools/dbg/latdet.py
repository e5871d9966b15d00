"""Determinism probe: lattice-mode counters under different CTA variants / repeated runs."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path[:0] = [os.path.join(ROOT, "vosk-api_b200"), os.path.join(ROOT, "vosk-api_b200", "tools"), ROOT]
import numpy as np, bench, vosk
vosk.SetLogLevel(-1)
streams = 256
waves = bench.make_audio(streams, 0, 5.0, 8.0)
lengths = np.array([len(w) for w in waves], dtype=np.int32)
stride = int((lengths.max() + 7) // 8 * 8)
mat = np.zeros((streams, stride), dtype=np.int16)
for i, w in enumerate(waves):
    mat[i, :len(w)] = w
for opts in ("lattice=2,heavy-tokens=0", "lattice=2,heavy-tokens=1000000", "lattice=2", "lattice=2", "lattice=2,pipeline-slots=1"):
    m = vosk.BatchModel(bench.model_dir(), options="num-channels=%d,max-batch-size=%d,max-seconds=10,%s" % (streams, streams, opts))
    for rep in range(2):
        m.ResetStats()
        ms, texts = m.RunResident(mat, lengths)
        st = m.Stats()
        print(opts, rep, "tokens", int(st["tokens"]), "new", int(st["tokens_new"]), "staged", int(st["arcs_staged"]), "links", int(st["links"]), "lat_arcs", int(st["lattice_arcs"]), "eps", int(st["arcs_epsilon"]))
    del m
