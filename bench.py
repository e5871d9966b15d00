#!/usr/bin/env python3
"""bench.py — aggregate real-time factor (audio-seconds decoded per second) of the batch recognition path.

Workload = BASELINE.json configs[1]: small-en-us architecture (random-init TDNN-F, synthetic ~20 MB HCLG),
512 concurrent 16 kHz streams of U(8,16) s synthetic speech-like audio per GPU.  One "step" = decoding
the whole 512-stream batch once.
Result mode = the reference's: lattice -> phone-pruned determinization -> 0.9 LM scale -> word alignment -> MBR
[REF src/batch_recognizer.cc:43-107,138-149] (engine default lattice=1; the best-path mode is an extra leg).
  value : device-resident run (samples already in HBM; CUDA events) — whole-job audio-s / s, results (host lattice chain) included
  e2e   : the reference-facing C ABI (vosk_batch_recognizer_accept_waveform in 8000-byte calls, round robin
          as in [REF python/example/test_gpu_batch.py:27-51], vosk_batch_model_wait, front_result/pop) with
          host buffers, host<->device copies inside the timed region
  large : BASELINE.json configs[2] / [4] leg in the same line (assumed en-us-0.22 architecture, multi-GB HCLG, 1024 streams per GPU)
  --impl reference : the CPU restatement of the reference's recognizer path (oracle/, "port") on the host cores
  --inproc : one process, one BatchModel spanning --gpus devices, fed by the native multi-threaded feeder
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "vosk-api_b200"), os.path.join(ROOT, "vosk-api_b200", "tools"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "aggregate RTFx (audio-sec/sec)"
UNIT = "audio-seconds/second"
STREAMS = 512
MODEL_CACHE = os.environ.get("VB_MODEL_CACHE", "/tmp/vb_bench_models")


def model_dir(arch="small"):
    import vbmodel
    root = os.path.join(MODEL_CACHE, arch + "_0")
    if not os.path.exists(os.path.join(root, "model", "graph", "HCLG.fst")):
        vbmodel.write_model_dir(root, arch, 0)
    return os.path.join(root, "model")


def make_audio(n_streams, rank, lo=8.0, hi=16.0):
    """Synthetic speech-like streams, U(lo,hi) seconds, seed 1000+i."""
    import vbmodel
    waves = []
    # synthesising 512 x 12 s is slow in numpy: build 16 distinct voices and vary by circular shift / gain
    base = [vbmodel.synth_audio(hi, 1000 + k) for k in range(16)]
    rng = np.random.default_rng(12345 + rank)
    for i in range(n_streams):
        n = int(rng.uniform(lo, hi) * 16000)
        b = base[i % 16]
        sh = int(rng.integers(0, len(b)))
        w = np.roll(b, sh)[:n].astype(np.float32) * rng.uniform(0.7, 1.2)
        waves.append(np.clip(np.round(w), -32768, 32767).astype(np.int16))
    return waves


class ClockSampler:
    def __init__(self, device):
        self.rows = []
        self.proc = None
        self.device = device

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def e2e_pass(vosk, model, pieces, wait_each_round=False):
    """The reference's batch driver loop [REF python/example/test_gpu_batch.py:27-51]: 8000-byte reads fed round
    robin to one recognizer per stream, FinishStream at EOF, Wait, then drain the results."""
    recs = [vosk.BatchRecognizer(model, 16000.0) for _ in pieces]
    texts = [""] * len(pieces)
    rounds = max(len(p) for p in pieces)
    for r in range(rounds + 1):
        for i, rec in enumerate(recs):
            if r < len(pieces[i]):
                rec.AcceptWaveform(pieces[i][r])
            elif r == len(pieces[i]):
                rec.FinishStream()
        if wait_each_round:
            model.Wait()
    model.Wait()
    for i, rec in enumerate(recs):
        texts[i] = rec.Result()
    return texts


def latency_pass(vosk, mdir, n_streams, seconds, packet_ms=100, fpc=10):
    """BASELINE.json configs[3]: low-latency streaming with partial results.  n_streams recognizers are fed packet_ms
    packets in real time (one packet per stream per tick); the engine records, for every chunk, the time from the
    acceptance of its last sample to the moment the partial covering it is retrievable."""
    model = vosk.BatchModel(mdir, options="partials=1,frames-per-chunk=%d,num-channels=%d,max-batch-size=%d,max-seconds=%d"
                            % (fpc, n_streams, min(n_streams, 1024), int(seconds) + 4))
    waves = make_audio(64, 77, seconds, seconds + 0.01)
    step = packet_ms * 16
    recs = [vosk.BatchRecognizer(model, 16000.0) for _ in range(n_streams)]
    pk = [[w[i:i + step].tobytes() for i in range(0, len(w) - step + 1, step)] for w in waves]
    n_ticks = min(len(p) for p in pk)
    # warm-up tick (first launches, allocator) is not counted
    for i, r in enumerate(recs):
        r.AcceptWaveform(pk[i % 64][0])
    model.Wait()
    model.Latency(reset=True)
    t0 = time.perf_counter()
    late = 0
    for k in range(1, n_ticks):
        target = t0 + (k - 1) * packet_ms / 1000.0
        now = time.perf_counter()
        if now < target:
            time.sleep(target - now)
        elif now - target > packet_ms / 1000.0:
            late += 1
        for i, r in enumerate(recs):
            r.AcceptWaveform(pk[i % 64][k])
    model.Wait()
    wall = time.perf_counter() - t0
    lat = model.Latency()
    sample = recs[0].PartialResult()
    for r in recs:
        r.FinishStream()
    model.Wait()
    del recs
    del model
    return {"p50_ms": lat["p50"], "p90_ms": lat["p90"], "p99_ms": lat["p99"], "mean_ms": lat["mean"], "chunks": lat["count"],
            "streams": n_streams, "frames_per_chunk": fpc, "packet_ms": packet_ms, "feed": "real-time paced, one packet per stream per tick",
            "ticks": n_ticks - 1, "late_ticks": late, "wall_s": wall, "audio_s_per_stream": (n_ticks - 1) * packet_ms / 1000.0,
            "definition": "acceptance of a chunk's last sample -> partial result covering it retrievable (host clock, engine-side)",
            "sample_partial": sample[:80]}


def reduce_over_ranks(x, op="max", device="cpu"):
    """Max / sum of a python float over the ranks of the default process group (identity when not distributed)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def job_throughput(audio_seconds_local, seconds_local, device="cpu"):
    """Whole-job metric: streams are sharded by utterance with no exchange, so the job decodes the SUM of the ranks'
    audio in the MAX of the ranks' times."""
    return reduce_over_ranks(audio_seconds_local, "sum", device) / reduce_over_ranks(seconds_local, "max", device)


def run_oracle_sample(n_streams, threads):
    """CPU restatement of the reference recognizer path on a bounded sample; returns (audio_s, wall_s)."""
    import oracle
    import vbmodel
    model = vbmodel.load_model_dir(model_dir())
    waves = make_audio(n_streams, 999)
    rc = oracle.ResultCtx(model)
    oracle.recognize(model, waves[0][:16000], rc=rc)  # warm (table construction, page-in)
    audio = sum(len(w) for w in waves) / 16000.0
    t0 = time.perf_counter()
    if threads <= 1:
        for w in waves:
            oracle.recognize(model, w, rc=rc)
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL inside the oracle calls
            list(ex.map(lambda w: oracle.recognize(model, w, rc=rc), waves))
    return audio, time.perf_counter() - t0


def large_lattice(a):
    """BASELINE.json configs[2] as a single-GPU, device-resident measurement (not the headline line): large architecture
    (assumed), vocabulary 200 k / 64 successors -> ~1.4e8 arcs (~2.2 GB of arc records in HBM), lattice
    generation on the device (link log + lattice-beam pruning + compaction, raw lattices copied to the host)."""
    import vosk
    streams = a.streams if a.streams != STREAMS else 1024
    t0 = time.perf_counter()
    mdir = model_dir("large")
    t_gen = time.perf_counter() - t0
    vosk.SetLogLevel(0)
    opts = ("lattice=2,num-channels=%d,max-batch-size=%d,max-seconds=18,log-links-per-frame=4096,lat-link-cap=131072,lat-tok-cap=65536"
            % (streams, min(streams, 1024)))
    if a.options:
        opts += "," + a.options
    t0 = time.perf_counter()
    model = vosk.BatchModel(mdir, options=opts)
    t_load = time.perf_counter() - t0
    waves = make_audio(streams, 0)
    lengths = np.array([len(w) for w in waves], dtype=np.int32)
    stride = int((lengths.max() + 7) // 8 * 8)
    mat = np.zeros((streams, stride), dtype=np.int16)
    for i, w in enumerate(waves):
        mat[i, :len(w)] = w
    audio_s = float(lengths.sum()) / 16000.0
    for _ in range(max(1, a.warmup)):
        model.RunResident(mat, lengths)
    model.ResetStats()
    ms_total = 0.0
    for _ in range(a.steps):
        ms, texts = model.RunResident(mat, lengths)
        ms_total += ms
    st = model.Stats()
    model.ResetStats()
    model.SetTiming(True)
    model.SetSlots(1)
    model.RunResident(mat, lengths)
    ser = model.Stats()
    T, Ae, Aeps, N = ser["tokens"], ser["arcs_emitting"], ser["arcs_epsilon"], ser["tokens_new"]
    search_bytes = T * 16 + (Ae + Aeps) * 20 + Ae * 4 + N * 16
    out_frames = sum((int(1 + (n - 400) // 160) + 2) // 3 for n in lengths)
    print(json.dumps({"metric": METRIC, "value": audio_s * a.steps / (ms_total / 1000.0), "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
                      "ms_per_step": ms_total / a.steps, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": "large-lattice (BASELINE.json configs[2]): assumed en-us-0.22 architecture, synthetic HCLG vocab 200k x 64 successors, "
                                             "%d streams of U(8,16) s, lattice generation on the device (lattice=2)" % streams, "options": opts},
                      "kernel_ms_per_step": {"mfcc": ser["ms_feat"], "ivector": ser["ms_ivector"], "tdnnf": ser["ms_nnet"], "search": ser["ms_search"]},
                      "tdnnf_tflops_algorithmic": 47e6 * out_frames / (ser["ms_nnet"] / 1000.0) / 1e12,
                      "search_gbs_algorithmic": search_bytes / (ser["ms_search"] / 1000.0) / 1e9,
                      "links_logged_per_step": st["links"] / a.steps, "lattice_arcs_per_step": st["lattice_arcs"] / a.steps,
                      "nonempty_results": sum(1 for t in texts if '"text" : ""' not in t), "audio_seconds_per_step": audio_s,
                      "model_generation_s": t_gen, "model_load_s": t_load}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine")
    ap.add_argument("--streams", type=int, default=STREAMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--options", default="")
    ap.add_argument("--wait-each-round", action="store_true", help="call Wait() after every feeding round, as the reference example does")
    ap.add_argument("--no-extras", action="store_true", help="skip the lattice-mode and partial-latency legs")
    ap.add_argument("--workload", default="small", choices=["small", "large-lattice"],
                    help="small = BASELINE.json configs[1] (the headline line); large-lattice = configs[2]: assumed en-us-0.22 architecture, "
                         "synthetic multi-GB HCLG, lattice generation on the device, --streams (default 1024) streams, device-resident leg only")
    ap.add_argument("--latency-streams", type=int, default=2048)
    ap.add_argument("--latency-seconds", type=float, default=4.0)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "small-en-us arch (random-init TDNN-F, synthetic HCLG ~1.0M arcs), %d concurrent 16 kHz streams of U(8,16) s per GPU" % a.streams,
              "streams_per_gpu": a.streams, "frames_per_chunk": 51, "beam": 13.0, "lattice_beam": 6.0, "max_active": 7000,
              "l2": "inputs+state larger than L2 (audio ~190 MB, token logs GBs)", "sharding": "streams by utterance, no collective",
              "result_mode": "best path (lattice=0: word-aligned best path, conf 1); lattice generation and the host MBR chain are the separate lattice_mode legs",
              "tdnnf_arithmetic": "fp32 accumulate; operands split into fp16 hi + scaled fp16 lo, 3 f16 MMAs per product (log-likelihoods within 1e-3 of the fp64-accumulating oracle)"}

    if a.workload == "large-lattice":
        return large_lattice(a)
    if a.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        n = max(cores, 8)
        vals = []
        for _ in range(max(1, min(a.steps, 2))):
            audio, wall = run_oracle_sample(n, cores)
            vals.append(audio / wall)
        v = float(np.mean(vals))
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": config,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": "%d streams of U(8,16) s, one stream per thread, %d threads" % (n, cores)},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    os.environ["VOSK_BATCH_DEVICES"] = str(local_rank)
    if rank == 0:
        model_dir()  # generate once
    if world > 1:
        dist.barrier()
    import vosk
    mdir = model_dir()
    vosk.SetLogLevel(-1)
    opts = "num-channels=%d,max-batch-size=%d,max-seconds=18" % (a.streams, min(a.streams, 1024))
    if a.options:
        opts += "," + a.options
    model = vosk.BatchModel(mdir, options=opts)
    waves = make_audio(a.streams, rank)
    lengths = np.array([len(w) for w in waves], dtype=np.int32)
    stride = int((lengths.max() + 7) // 8 * 8)
    audio_mat = np.zeros((a.streams, stride), dtype=np.int16)
    for i, w in enumerate(waves):
        audio_mat[i, :len(w)] = w
    audio_s = float(lengths.sum()) / 16000.0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        return reduce_over_ranks(x, "max", "cuda")

    # ---------------- value: device-resident ----------------
    texts = None
    for _ in range(a.warmup):
        _, texts = model.RunResident(audio_mat, lengths)
    model.ResetStats()
    model.SetTiming(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ms, texts = model.RunResident(audio_mat, lengths)
        dev_ms += ms
    barrier()
    wall_resident = time.perf_counter() - t0
    clocks = sampler.stop()
    st = model.Stats()
    # one extra pass with the pipeline slots serialized: per-stage device times (CUDA events on the launching
    # stream) free of cross-slot overlap; these are the durations the roofline uses
    model.ResetStats()
    model.SetSlots(1)
    model.RunResident(audio_mat, lengths)
    ser = model.Stats()
    model.SetSlots(64)
    model.SetTiming(False)
    dev_s = max_over_ranks(dev_ms / 1000.0)
    audio_total = reduce_over_ranks(audio_s, "sum", "cuda")   # every rank takes part in every collective
    value = audio_total * a.steps / dev_s

    # ---------------- e2e: through the C ABI with host buffers ----------------
    pieces = [[w[i:i + 4000].tobytes() for i in range(0, len(w), 4000)] for w in waves]  # 8000-byte reads, as the reference driver
    for _ in range(min(a.warmup, 1)):
        e2e_pass(vosk, model, pieces, a.wait_each_round)
    barrier()
    t0 = time.perf_counter()
    e2e_texts = None
    for _ in range(a.steps):
        e2e_texts = e2e_pass(vosk, model, pieces, a.wait_each_round)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = audio_total * a.steps / e2e_s
    same = sum(1 for x, y in zip(texts, e2e_texts) if x == y)

    # ---------------- roofline of the dominant kernel ----------------
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks = {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops": pk["bf16_tflops_sustained"], "src": "measured"}
    except Exception:
        pass
    kernel_ms_overlapped = {"mfcc": st["ms_feat"] / a.steps, "ivector": st["ms_ivector"] / a.steps, "tdnnf": st["ms_nnet"] / a.steps, "search": st["ms_search"] / a.steps}
    kernel_ms = {"mfcc": ser["ms_feat"], "ivector": ser["ms_ivector"], "tdnnf": ser["ms_nnet"], "search": ser["ms_search"]}
    dominant = max(kernel_ms, key=kernel_ms.get)
    launches_timed = int(st["launches"])
    st = ser     # counters of the serialized pass (one step) feed the byte model
    roof_steps = 1
    out_frames = sum((int(1 + (n - 400) // 160) + 2) // 3 for n in lengths) * roof_steps
    flop = 7.95e6 * out_frames  # 7.95 MFLOP per output frame, small architecture
    # SURVEY.md §8(d) beam-search byte model from the in-kernel counters
    T, Ae, Aeps, N = st["tokens"], st["arcs_emitting"], st["arcs_epsilon"], st["tokens_new"]
    search_bytes = T * 16 + (Ae + Aeps) * 20 + Ae * 4 + N * 16
    feat_bytes = 480.0 * sum(int(1 + (n - 400) // 160) for n in lengths) * roof_steps
    if dominant == "tdnnf":
        ach = flop / (kernel_ms["tdnnf"] / 1000.0) / 1e12
        roof = {"kernel": "gemm_tc_kernel (TDNN-F chain, fp16 hi/lo operand split, fp32 accumulate)", "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": None, "peak_src": peaks["src"] + " bf16 sustained (f16 MMAs, 3 per algorithmic MAC)"}
    elif dominant == "search":
        ach = search_bytes / (kernel_ms["search"] / 1000.0) / 1e9
        roof = {"kernel": "decode_kernel", "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_src": peaks["src"]}
    else:
        ach = feat_bytes / (kernel_ms[dominant] / 1000.0) / 1e9
        roof = {"kernel": dominant, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": None, "peak_src": peaks["src"]}

    # ---------------- extras (rank 0): lattice generation on, and the low-latency / partial-result configuration ----------------
    del model
    lattice_mode = None
    partial_latency = None
    if rank == 0 and not a.no_extras:
        try:
            res = {}
            for mode, name in ((2, "device_lattice"), (1, "device_lattice_plus_host_mbr")):
                lm = vosk.BatchModel(mdir, options=opts + ",lattice=%d" % mode)
                lm.RunResident(audio_mat, lengths)
                lm.ResetStats()
                lm.SetTiming(True)
                ms, ltexts = lm.RunResident(audio_mat, lengths)
                lst = lm.Stats()
                res[name] = {"value": audio_s / (ms / 1000.0), "unit": UNIT, "ms_per_step": ms, "links_logged": lst["links"],
                             "lattice_arcs_after_pruning": lst["lattice_arcs"], "search_ms_overlapped": lst["ms_search"]}
                if mode == 1:
                    res[name]["results_with_confidence_below_1"] = sum(1 for t in ltexts if '"conf" : 0.' in t)
                    res[name]["host_threads"] = "hardware threads / 2 (lattice pool)"
                del lm
            lattice_mode = res
        except Exception as e:  # the extras never take the headline line down
            lattice_mode = {"error": str(e)[:200]}
        try:
            partial_latency = latency_pass(vosk, mdir, a.latency_streams, a.latency_seconds)
        except Exception as e:
            partial_latency = {"error": str(e)[:200]}

    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get({"search": "decode_kernel", "tdnnf": "gemm_tc_kernel", "ivector": "ivector_kernel", "mfcc": "mfcc_kernel"}[dominant])
        if ent:
            traffic = ent["value"]
            roof["traffic_note"] = ent["unit"] + "; from " + tj["_source"]
            if dominant == "search":  # the same unit for the algorithmic side: bytes of one search step (512 lanes x 17 frames)
                steps_run = max(1.0, st["steps"])
                roof["algorithmic_bytes_per_search_step"] = search_bytes / steps_run
    except Exception:
        pass
    roof["traffic"] = traffic
    # the same figures for every stage: algorithmic bytes or flops of one serialized step / its device time
    roof_all = {
        "mfcc": {"bound": "hbm", "achieved": feat_bytes / (kernel_ms["mfcc"] / 1000.0) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "algorithmic": "480 B per frame"},
        "tdnnf": {"bound": "tensor", "achieved": flop / (kernel_ms["tdnnf"] / 1000.0) / 1e12, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                  "algorithmic": "7.95 MFLOP per output frame (the fp16 hi/lo split issues 3x that on the tensor pipe)"},
        "search": {"bound": "hbm", "achieved": search_bytes / (kernel_ms["search"] / 1000.0) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                   "algorithmic": "T*16 + (Ae+Aeps)*20 + Ae*4 + N*16 bytes from the in-kernel counters"},
    }
    for v in roof_all.values():
        v["frac"] = v["achieved"] / v["peak"]

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        ca, cw = run_oracle_sample(12, 1)
        cpu = {"value": ca / cw, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "12 streams of U(8,16) s decoded one after another on one core (%.1f s of CPU work)" % cw}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dev_s * 1000.0 / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(lengths.sum()) * 2 + a.streams * 64,
                        "d2h_bytes_per_step": a.streams * (4 * (18 * 100 // 3 + 2) * 4 + 64 * 4 + 32),
                        "transcripts_equal_to_resident_run": "%d/%d" % (same, len(texts))},
                "gpu_launches": launches_timed, "clocks": clocks, "roofline": roof, "roofline_all_stages": roof_all, "cpu_baseline": cpu,
                "kernel_ms_per_step": kernel_ms, "kernel_ms_per_step_overlapped": kernel_ms_overlapped,
                "roofline_note": "stage durations from one extra pass with the pipeline slots serialized (no overlap); CUDA events on the launching stream",
                "host_wall_ms_per_step_resident": wall_resident * 1000.0 / a.steps,
                "search_counters_per_step": {"tokens": T, "arcs_emitting": Ae, "arcs_epsilon": Aeps, "tokens_new": N},
                "audio_seconds_per_step": audio_total,
                "lattice_mode": lattice_mode, "partial_latency": partial_latency}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
