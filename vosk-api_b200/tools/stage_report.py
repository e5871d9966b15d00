"""Per-stage report of the bench workload (device-resident): python tools/stage_report.py STREAMS LO HI "opts1" ["opts2" ...]
For every option set: one overlapped run and one run with the pipeline slots serialized; prints the stage times, the search
phase cycles and the host-side lattice figures as one JSON line per run."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.dirname(ROOT))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import vosk  # noqa: E402

streams, lo, hi = int(sys.argv[1]), float(sys.argv[2]), float(sys.argv[3])
arch = os.environ.get("VB_ARCH", "small")
vosk.SetLogLevel(-1)
waves = bench.make_audio(streams, 0, lo, hi)
lengths = np.array([len(w) for w in waves], dtype=np.int32)
stride = int((lengths.max() + 7) // 8 * 8)
mat = np.zeros((streams, stride), dtype=np.int16)
for i, w in enumerate(waves):
    mat[i, :len(w)] = w
audio = float(lengths.sum()) / 16000.0
for extra in sys.argv[4:]:
    model = vosk.BatchModel(bench.model_dir(arch), options="num-channels=%d,max-batch-size=%d,max-seconds=18%s" % (streams, min(streams, 1024), "," + extra if extra else ""))
    model.SetTiming(True)
    model.RunResident(mat, lengths)
    for slots in (64, 1):
        model.SetSlots(slots)
        model.ResetStats()
        ms, texts = model.RunResident(mat, lengths)
        st = model.Stats()
        cyc = {k[4:]: round(v / 1e6, 1) for k, v in st.items() if k.startswith("cyc_") and v}
        slow = {k[8:]: round(v / 1e3, 1) for k, v in st.items() if k.startswith("slowest_") and v}
        slow.update({k: v for k, v in st.items() if k.startswith("lane_launches_")})
        keep = ("steps", "launches", "ms_feat", "ms_ivector", "ms_nnet", "ms_search", "ms_prune", "host_launch_ms", "host_complete_ms", "host_fetch_ms",
                "post_ms", "post_jobs", "post_threads", "links", "lattice_arcs", "tokens", "arcs_emitting", "tokens_new", "max_tokens_per_frame",
                "lane_cycles_sum", "lane_cycles_max", "truncated", "lattice_fallbacks")
        print(json.dumps({"options": extra, "slots": slots, "ms": round(ms, 2), "rtfx": round(audio / ms * 1000.0), "audio_s": round(audio, 1),
                          **{k: round(st[k], 2) for k in keep if k in st}, "Mcycles": cyc, "slowest_lane_kcycles": slow,
                          "conf_below_1": sum(1 for t in texts if '"conf" : 0.' in t)}), flush=True)
    del model
