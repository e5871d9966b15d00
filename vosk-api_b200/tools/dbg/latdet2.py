"""Find a stream whose pruned lattice differs between two identical runs (capture mode)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path[:0] = [os.path.join(ROOT, "vosk-api_b200"), os.path.join(ROOT, "vosk-api_b200", "tools"), ROOT, os.path.join(ROOT, "tests")]
import numpy as np, bench, vosk, helpers
vosk.SetLogLevel(0)
streams = 256
waves = bench.make_audio(streams, 0, 5.0, 8.0)
runs = []
for rep in range(12):
    got, st = helpers.run_engine(bench.model_dir(), waves, options="lattice=2,num-channels=%d,max-batch-size=%d,max-seconds=10" % (streams, streams), bytes_per_call=16320)
    print("rep", rep, "lat_arcs", int(st["lattice_arcs"]), "links", int(st["links"]), "errors", sum(1 for g in got if g["error"]))
    runs.append(got)
base = runs[0]
for rep in range(1, 12):
    for i, (a, b) in enumerate(zip(base, runs[rep])):
        if a["lat_hdr"][1] != b["lat_hdr"][1] or a["lat_hdr"][0] != b["lat_hdr"][0]:
            print("rep", rep, "stream", i, "hdr", a["lat_hdr"], b["lat_hdr"])
            ka = set(map(tuple, np.stack([a["lat_tok_frame"][a["lat_links"][:,0]], a["lat_tok_state"][a["lat_links"][:,0]], a["lat_links"][:,2]],1).tolist()))
            kb = set(map(tuple, np.stack([b["lat_tok_frame"][b["lat_links"][:,0]], b["lat_tok_state"][b["lat_links"][:,0]], b["lat_links"][:,2]],1).tolist()))
            print("  only in A:", sorted(ka-kb)[:10], " only in B:", sorted(kb-ka)[:10], "frames", a["lat_hdr"][5])
