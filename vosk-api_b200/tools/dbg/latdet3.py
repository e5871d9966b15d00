import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path[:0] = [os.path.join(ROOT, "vosk-api_b200"), os.path.join(ROOT, "vosk-api_b200", "tools"), ROOT, os.path.join(ROOT, "tests")]
import numpy as np, bench, vosk, helpers
vosk.SetLogLevel(-1)
streams = 256
waves = bench.make_audio(streams, 0, 5.0, 8.0)
for rep in range(12):
    got, st = helpers.run_engine(bench.model_dir(), waves, options="lattice=2,num-channels=%d,max-batch-size=%d,max-seconds=10" % (streams, streams), bytes_per_call=16320)
    print("rep", rep, "lat_arcs", int(st["lattice_arcs"]), "prune mismatches (same input, two launches)", int(st["cyc_light_x"]))
