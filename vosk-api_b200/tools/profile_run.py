"""Short device-resident run for ncu / compute-sanitizer: python tools/profile_run.py [streams] [seconds] [options]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.dirname(ROOT))
import numpy as np  # noqa: E402

import bench  # noqa: E402
import vosk  # noqa: E402

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 512
seconds = float(sys.argv[2]) if len(sys.argv) > 2 else 3.0
extra = sys.argv[3] if len(sys.argv) > 3 else ""
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
vosk.SetLogLevel(-1)
model = vosk.BatchModel(bench.model_dir(), options="num-channels=%d,max-batch-size=%d,max-seconds=18%s" % (streams, min(streams, 1024), "," + extra if extra else ""))
waves = bench.make_audio(streams, 0, seconds, seconds + 0.5)
lengths = np.array([len(w) for w in waves], dtype=np.int32)
stride = int((lengths.max() + 7) // 8 * 8)
mat = np.zeros((streams, stride), dtype=np.int16)
for i, w in enumerate(waves):
    mat[i, :len(w)] = w
model.SetTiming(not os.environ.get("VB_NOTIMING"))  # (stage timing keeps the front end on one stream)
if os.environ.get("VB_SLOTS"):  # 1 = serialize the steps: per-stage device times free of overlap
    model.SetSlots(int(os.environ["VB_SLOTS"]))
for _ in range(reps):
    ms, texts = model.RunResident(mat, lengths)
st = model.Stats()
print("device ms %.2f  audio %.1f s  RTFx %.0f" % (ms, lengths.sum() / 16000.0, lengths.sum() / 16.0 / ms))
print({k: round(v, 2) for k, v in st.items()})
