#!/usr/bin/env python3
"""bench.py — aggregate real-time factor (audio-seconds decoded per second) of the batch recognition path.

Workload = BASELINE.json configs[1]: small-en-us architecture (random-init TDNN-F, synthetic ~20 MB HCLG),
512 concurrent 16 kHz streams of U(8,16) s synthetic speech-like audio per GPU.  One "step" = decoding the whole
512-stream batch once.  Steps run back to back as in continuous serving: the streams of step k+1 are queued while the
lattice chain of step k's results is still running; every result of every step is delivered inside the timed region.
Result mode = the reference's: lattice -> phone-pruned determinization -> 0.9 LM scale -> word alignment -> MBR
[REF src/batch_recognizer.cc:43-107,138-149] (engine default lattice=1; the best-path mode is an extra leg).
  value : device-resident run (samples already in HBM; CUDA events) — whole-job audio-s / s, results (host lattice chain) included
  e2e   : the reference-facing C ABI (vosk_batch_recognizer_accept_waveform in 8000-byte calls, round robin
          as in [REF python/example/test_gpu_batch.py:27-51], vosk_batch_model_wait, front_result/pop) with
          host buffers, host<->device copies inside the timed region
  large : BASELINE.json configs[2] / [4] leg in the same line (assumed en-us-0.22 architecture, multi-GB HCLG, 1024 streams per GPU)
  parity_sample : 12 of the bench's own streams checked against the CPU oracle (texts identical on the engine's log-likelihoods)
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




def e2e_run(vosk, model, pieces, steps):
    """The reference's batch driver loop [REF python/example/test_gpu_batch.py:27-51]: 8000-byte reads fed round robin to one
    recognizer per stream, FinishStream at EOF — `steps` times over with new recognizers, back to back (the accept calls never
    block), then Wait and the results of every recognizer of every step.  Returns the texts of the last step and the number
    of streams of earlier steps whose text differs from it."""
    rounds = max(len(p) for p in pieces)
    all_recs = []
    for _ in range(steps):
        recs = [vosk.BatchRecognizer(model, 16000.0) for _ in pieces]
        for r in range(rounds + 1):
            for i, rec in enumerate(recs):
                if r < len(pieces[i]):
                    rec.AcceptWaveform(pieces[i][r])
                elif r == len(pieces[i]):
                    rec.FinishStream()
        all_recs.append(recs)
    model.Wait()
    texts = [[rec.Result() for rec in recs] for recs in all_recs]
    bad = sum(1 for t in texts[:-1] for x, y in zip(t, texts[-1]) if x != y)
    return texts[-1], bad


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
    # caller-side sample: a second thread polls vosk_batch_recognizer_partial_frames of a few streams and takes, for every packet,
    # the time from the return of its accept call to the first poll that sees the partial grow (the poll period is part of it)
    watched = list(range(0, n_streams, max(1, n_streams // 8)))[:8]
    t_acc = {i: 0.0 for i in watched}
    seq = {i: 0 for i in watched}
    caller_lat, stop = [], threading.Event()

    def poll():
        last_f = {i: recs[i].PartialFrames() for i in watched}
        done = {i: 0 for i in watched}
        while not stop.is_set():
            for i in watched:
                f = recs[i].PartialFrames()
                if f > last_f[i]:
                    last_f[i] = f
                    sq, ta = seq[i], t_acc[i]
                    if sq > done[i]:
                        done[i] = sq
                        caller_lat.append(time.perf_counter() - ta)

    old_switch = sys.getswitchinterval()
    sys.setswitchinterval(1e-4)  # the poller must not wait 5 ms for the interpreter while the feeder loops
    poller = threading.Thread(target=poll, daemon=True)
    poller.start()
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
            if i in t_acc:
                t_acc[i] = time.perf_counter()
                seq[i] = k
    model.Wait()
    wall = time.perf_counter() - t0
    time.sleep(0.02)
    stop.set()
    poller.join()
    sys.setswitchinterval(old_switch)
    cl = np.sort(np.array(caller_lat)) * 1e3 if caller_lat else np.zeros(1)
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
            "caller_side": {"p50_ms": float(cl[len(cl) // 2]), "p90_ms": float(cl[int(len(cl) * 0.9)]), "max_ms": float(cl[-1]), "packets": int(len(caller_lat)),
                            "streams_polled": len(watched),
                            "definition": "return of vosk_batch_recognizer_accept_waveform for a packet -> a polling thread sees "
                                          "vosk_batch_recognizer_partial_frames grow (Python poller beside the Python feeder; its poll period is included)"},
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
    """CPU restatement of the reference recognizer path (lattice -> MBR result, as the reference) on a bounded sample of the
    bench workload's own streams (rank 0's first n); returns (audio_s, wall_s)."""
    import oracle
    import vbmodel
    model = vbmodel.load_model_dir(model_dir())
    waves = make_audio(n_streams, 0)
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


def parity_sample(vosk, arch, waves, bench_texts, n, extra_opts="", acoustic_scale=None):
    """n of the bench's own streams against the CPU oracle, and the oracle's single-core speed on them.  The streams run
    once more through a small engine instance with the test taps on: (1) its text equals the bench run's (a result does not
    depend on the batch), (2) the ORACLE's search + lattice chain on the engine's log-likelihoods gives the identical text,
    confidences included, (3) the engine's log-likelihoods are within 1e-3 of the oracle's own."""
    import oracle
    import vbmodel
    mdir = model_dir(arch)
    model_np = vbmodel.load_model_dir(mdir)
    rc = oracle.ResultCtx(model_np)
    oracle.recognize(model_np, waves[0][:16000], rc=rc)
    m = vosk.BatchModel(mdir, options="num-channels=%d,max-batch-size=%d,max-seconds=18,debug-capture=1%s" % (n, n, "," + extra_opts if extra_opts else ""))
    recs = [vosk.BatchRecognizer(m, 16000.0) for _ in range(n)]
    for r, w in zip(recs, waves[:n]):
        r.DebugCapture()
        r.AcceptWaveform(w.tobytes())
        r.FinishStream()
    m.Wait()
    P = int(model_np["cfg"]["num-pdfs"])
    lb = float(model_np["conf"].get("lattice-beam", 6.0))
    D = int(model_np["cfg"]["ivector-dim"])
    same_bench = same_oracle = same_pipeline = 0
    max_dll = max_dnet = max_div = 0.0
    over = 0
    cpu_s = 0.0
    for i, (r, w) in enumerate(zip(recs, waves[:n])):
        text = r.Result()
        ll = r.DebugGet("loglikes", np.float32).reshape(-1, P)
        t0 = time.perf_counter()
        ref = oracle.recognize(model_np, w, stages=True, rc=rc)   # the whole CPU path, timed: the cpu_baseline sample
        cpu_s += time.perf_counter() - t0
        dll = float(np.abs(ll - ref["loglikes"]).max())
        max_dll = max(max_dll, dll)
        over += dll >= 1e-3
        # the network alone: the oracle's forward pass on the ENGINE's features and i-vectors
        iv = r.DebugGet("ivectors", np.float32).reshape(-1, D)
        mf = r.DebugGet("mfcc", np.float32).reshape(-1, 40)
        max_div = max(max_div, float(np.abs(iv - ref["ivectors"]).max()))
        max_dnet = max(max_dnet, float(np.abs(ll - oracle.nnet_forward(model_np, mf, iv, ref["iv_index"])).max()))
        # (acoustic-scale option: the search sees scale * log-likelihood, one fp32 product)
        dec = oracle.decode(model_np, ll if acoustic_scale is None else (np.float32(acoustic_scale) * ll).astype(np.float32), lattice_beam=lb)
        want = oracle.lattice_result(model_np, dec, lb, rc=rc)
        same_oracle += text == want
        same_bench += bench_texts is None or text == bench_texts[i]
        same_pipeline += text == ref["text"]
    del recs, m
    audio = sum(len(w) for w in waves[:n]) / 16000.0
    return {"streams": n, "texts_identical_to_oracle_search_and_lattice_chain_on_engine_loglikes": "%d/%d" % (same_oracle, n),
            "texts_identical_to_bench_run": "%d/%d" % (same_bench, n), "loglike_tolerance": 1e-3,
            "max_abs_loglike_diff_network (oracle TDNN-F on the engine's features and i-vectors)": max_dnet,
            "max_abs_loglike_diff_whole_front_end (oracle's own features and i-vectors)": max_dll, "streams_over_tolerance_whole_front_end": int(over),
            "max_abs_ivector_diff": max_div,
            "note": "a stream goes over 1e-3 end to end only after a frame whose UBM posterior sits on the min_post / top-5 pruning threshold: the fp32 "
                    "front end and the fp64 oracle then prune differently (as Kaldi's own CPU and GPU feature code do), the i-vector moves by ~2e-4 and "
                    "decays back; the network itself stays within the tolerance",
            "texts_identical_to_whole_oracle_pipeline (its own log-likelihoods; informational)": "%d/%d" % (same_pipeline, n)}, audio, cpu_s


def peaks_measured():
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback (B200_PROFILING.md)"}
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peaks = {"hbm_gbs": pk["hbm_gbs"], "bf16_tflops": pk["bf16_tflops_sustained"], "src": "measured (MEASURED_PEAKS.json; sustained bf16)"}
    except Exception:
        pass
    return peaks


def stage_rooflines(ser, lengths, arch, peaks):
    """Algorithmic bytes / flops of one serialized pass over the batch divided by each stage's device time."""
    frames = sum(int(1 + (n - 400) // 160) for n in lengths if n >= 400)
    out_frames = sum((int(1 + (n - 400) // 160) + 2) // 3 for n in lengths if n >= 400)
    mflop = 7.95e6 if arch == "small" else 47e6
    D = 40 if arch == "small" else 100
    T, Ae, Aeps, N = ser["tokens"], ser["arcs_emitting"], ser["arcs_epsilon"], ser["tokens_new"]
    search_bytes = T * 16 + (Ae + Aeps) * 20 + Ae * 4 + N * 16
    chunks = ser["lanes"]
    iv_bytes = chunks * 2 * (D + D * (D + 1) // 2) * 4 + frames * 40 * 4
    prune_bytes = ser["links"] * 16 + ser["tokens"] * 8 + ser["lattice_arcs"] * 16   # link log read once, token costs + extras, survivors written
    ms = {"mfcc": ser["ms_feat"], "ivector": ser["ms_ivector"], "tdnnf": ser["ms_nnet"], "search": ser["ms_search"], "lattice_prune": ser["ms_prune"]}
    alg = {"mfcc": (480.0 * frames, "hbm", "480 B per frame"),
           "ivector": (iv_bytes, "hbm", "2 x (D + D(D+1)/2) x 4 B of statistics per chunk + 160 B per frame"),
           "tdnnf": (mflop * out_frames, "tensor", "%.2f MFLOP per output frame (the fp16 hi/lo operand split issues 3x that on the tensor pipe)" % (mflop / 1e6)),
           "search": (search_bytes, "hbm", "T*16 + (Ae+Aeps)*20 + Ae*4 + N*16 bytes from the in-kernel counters"),
           "lattice_prune": (prune_bytes, "hbm", "16 B per logged link + 8 B per logged token + 16 B per surviving link")}
    out = {}
    for k, (work, bound, note) in alg.items():
        if not ms[k]:
            continue
        if bound == "hbm":
            ach, peak, unit = work / (ms[k] / 1e3) / 1e9, peaks["hbm_gbs"], "GB/s"
        else:
            ach, peak, unit = work / (ms[k] / 1e3) / 1e12, peaks["bf16_tflops"], "TFLOP/s"
        out[k] = {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak, "ms": ms[k], "algorithmic": note}
    return out, search_bytes


def resident_leg(vosk, mdir, opts, audio_mat, lengths, warmup, steps):
    """One model, W warm-up + K timed back-to-back passes over the resident batch; returns a summary dict."""
    m = vosk.BatchModel(mdir, options=opts)
    m.RunResident(audio_mat, lengths, passes=max(1, warmup))
    m.ResetStats()
    ms, texts = m.RunResident(audio_mat, lengths, passes=steps)
    st = m.Stats()
    audio_s = float(lengths.sum()) / 16000.0
    out = {"value": audio_s * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "options": opts,
           "results_with_confidence_below_1": sum(1 for t in texts if '"conf" : 0.' in t), "nonempty_results": sum(1 for t in texts if '"text" : ""' not in t),
           "texts_differing_between_steps": m.resident_mismatches, "lattice_links_logged_per_step": st["links"] / steps,
           "lattice_arcs_after_pruning_per_step": st["lattice_arcs"] / steps, "host_lattice_chain_cpu_ms_per_step": st["post_ms"] / steps,
           "host_lattice_threads": st["post_threads"], "truncated": st["truncated"], "lattice_fallbacks": st["lattice_fallbacks"]}
    del m
    return out, texts


def large_leg(a, rank, world, dist, barrier, peaks):
    """BASELINE.json configs[2] (1 GPU) / configs[4] (8 GPUs, 8192 streams sharded by utterance): assumed en-us-0.22 architecture
    (SURVEY.md section 8), synthetic HCLG of vocabulary 200 k x 64 successors (~1.4e8 arcs, ~2.2 GB of arc records in HBM), 1024
    streams per GPU, the reference's lattice -> MBR result path.  Every rank runs it; rank 0 reports."""
    import vosk
    streams = a.large_streams
    local = {"ok": 0.0, "audio": 0.0, "ms": 1.0}
    info = {}
    t0 = time.perf_counter()
    if rank == 0:
        try:
            model_dir("large")   # generated once per box; the other ranks wait here
        except Exception as e:
            info = {"error": "model generation: " + str(e)[:200]}
    t_gen = time.perf_counter() - t0
    barrier()
    try:
        if "error" in info:
            raise RuntimeError(info["error"])
        mdir = model_dir("large")
        opts = "num-channels=%d,max-batch-size=%d,max-seconds=18,log-links-per-frame=4096,lat-link-cap=131072,lat-tok-cap=65536" % (streams, min(streams, 1024))
        if a.options:
            opts += "," + a.options
        t0 = time.perf_counter()
        model = vosk.BatchModel(mdir, options=opts)
        t_load = time.perf_counter() - t0
        waves = make_audio(streams, rank)
        lengths = np.array([len(w) for w in waves], dtype=np.int32)
        stride = int((lengths.max() + 7) // 8 * 8)
        mat = np.zeros((streams, stride), dtype=np.int16)
        for i, w in enumerate(waves):
            mat[i, :len(w)] = w
        audio_s = float(lengths.sum()) / 16000.0
        model.RunResident(mat, lengths, passes=1)
        model.ResetStats()
        ms, texts = model.RunResident(mat, lengths, passes=a.large_steps)   # ranks are independent: job time = max over ranks
        st = model.Stats()
        model.ResetStats()
        model.SetTiming(True)
        model.SetSlots(1)
        model.RunResident(mat, lengths)
        ser = model.Stats()
        roofs, _ = stage_rooflines(ser, lengths, "large", peaks)
        dom = max(("mfcc", "ivector", "tdnnf", "search"), key=lambda k: roofs[k]["ms"] if k in roofs else 0)
        local = {"ok": 1.0, "audio": audio_s * a.large_steps, "ms": ms}
        info = {"config": {"workload": "BASELINE.json configs[2] / [4]: assumed en-us-0.22 architecture (random-init, 6016 pdfs), synthetic HCLG vocab 200k x 64 successors "
                                       "(~1.4e8 arcs), %d streams of U(8,16) s per GPU, lattice -> determinization -> MBR results" % streams, "options": opts},
                "steps": a.large_steps, "ms_per_step": ms / a.large_steps, "audio_seconds_per_step_per_gpu": audio_s,
                "roofline": dict(roofs[dom], kernel=dom), "roofline_all_stages": roofs,
                "results_with_confidence_below_1": sum(1 for t in texts if '"conf" : 0.' in t), "nonempty_results": sum(1 for t in texts if '"text" : ""' not in t),
                "texts_differing_between_steps": model.resident_mismatches, "host_lattice_chain_cpu_ms_per_step": st["post_ms"] / a.large_steps,
                "truncated": st["truncated"], "lattice_fallbacks": st["lattice_fallbacks"], "model_generation_s": t_gen, "model_load_s": t_load}
        del model
        if rank == 0 and not a.no_large_parity:
            try:
                ps, _, _ = parity_sample(vosk, "large", waves, texts, 2, "log-links-per-frame=4096,lat-link-cap=131072,lat-tok-cap=65536")
                info["parity_sample"] = ps
            except Exception as e:
                info["parity_sample"] = {"error": str(e)[:200]}
    except Exception as e:  # never takes the headline line (or the other ranks' collectives) down
        info = {"error": str(e)[:300]}
    ok = reduce_over_ranks(local["ok"], "sum", "cuda") if world > 1 else local["ok"]
    audio = reduce_over_ranks(local["audio"], "sum", "cuda") if world > 1 else local["audio"]
    ms = reduce_over_ranks(local["ms"], "max", "cuda") if world > 1 else local["ms"]
    if ok == world and "error" not in info:
        info["value"] = audio / (ms / 1e3)
        info["unit"] = UNIT
        info["n_gpus"] = world
    elif "error" not in info:
        info["error"] = "the leg failed on another rank"
    return info


def inproc(a):
    """One process, one BatchModel spanning --gpus devices (the C-ABI user's multi-GPU path: one engine per device inside the
    model, streams assigned by id), fed through the reference ABI calls by the library's native multi-threaded feeder."""
    import vosk
    mdir = model_dir()
    vosk.SetLogLevel(-1)
    n = a.streams * a.gpus
    opts = "num-channels=%d,max-batch-size=%d,max-seconds=18,devices=%s" % (a.streams, min(a.streams, 1024), ":".join(str(i) for i in range(a.gpus)))
    if a.options:
        opts += "," + a.options
    model = vosk.BatchModel(mdir, options=opts)
    waves = []
    for r in range(a.gpus):
        waves += make_audio(a.streams, r)
    waves = [waves[(i % a.gpus) * a.streams + i // a.gpus] for i in range(n)]   # stream i lands on engine i % gpus: rank r's set per device
    audio_s = sum(len(w) for w in waves) / 16000.0
    threads = a.feeder_threads or min(32, os.cpu_count() or 8)
    for _ in range(max(1, min(a.warmup, 2))):
        model.FeedStreams(waves, 8000, threads, want_results=False)
    sampler = ClockSampler(0)
    sampler.start()
    model.ResetStats()
    t0 = time.perf_counter()
    texts = model.FeedStreams(waves, 8000, threads, passes=a.steps)   # steps back to back, one Wait: as the per-rank bench
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    st = model.Stats()
    v = audio_s * a.steps / wall
    print(json.dumps({"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": wall * 1e3 / a.steps,
                      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "mode": "inproc",
                      "config": {"workload": "small-en-us arch, %d streams of U(8,16) s per GPU, ONE process / ONE BatchModel over %d devices, native feeder (%d threads), "
                                             "reference ABI calls, lattice -> MBR results" % (a.streams, a.gpus, threads), "options": opts},
                      "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": st["h2d_bytes"] / a.steps, "d2h_bytes_per_step": st["d2h_bytes"] / a.steps,
                              "texts_differing_between_steps": model.feed_mismatches},
                      "gpu_launches": int(st["launches"]), "clocks": clocks, "nonempty_results": sum(1 for t in texts if '"text" : ""' not in t),
                      "results_with_confidence_below_1": sum(1 for t in texts if '"conf" : 0.' in t), "timing": "host wall clock around the feeder calls (one process)"}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine")
    ap.add_argument("--streams", type=int, default=STREAMS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the parity sample / CPU baseline leg")
    ap.add_argument("--options", default="")
    ap.add_argument("--no-extras", action="store_true", help="skip the best-path, device-lattice and partial-latency legs")
    ap.add_argument("--no-large", action="store_true", help="skip the large-architecture leg (BASELINE.json configs[2] / [4])")
    ap.add_argument("--no-large-parity", action="store_true")
    ap.add_argument("--large-streams", type=int, default=1024)
    ap.add_argument("--large-steps", type=int, default=2)
    ap.add_argument("--inproc", action="store_true", help="one process, one BatchModel over --gpus devices, native feeder; prints its own line")
    ap.add_argument("--feeder-threads", type=int, default=0)
    ap.add_argument("--latency-streams", type=int, default=2048)
    ap.add_argument("--latency-seconds", type=float, default=4.0)
    a = ap.parse_args()
    a.warmup = max(a.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = {"workload": "BASELINE.json configs[1]: small-en-us arch (random-init TDNN-F, synthetic HCLG ~1.0M arcs), %d concurrent 16 kHz streams of U(8,16) s per GPU" % a.streams,
              "streams_per_gpu": a.streams, "frames_per_chunk": 51, "beam": 13.0, "lattice_beam": 6.0, "max_active": 7000,
              "l2": "inputs+state larger than L2 (audio ~190 MB, token and link logs GBs)", "sharding": "streams by utterance, no collective",
              "result_mode": "lattice -> phone-pruned determinization -> 0.9 LM scale -> word alignment -> MBR (the reference's result path; engine default)",
              "steps": "back to back: the streams of step k+1 are queued while the lattice chain of step k's results runs; all results delivered inside the timed region",
              "tdnnf_arithmetic": "fp32 accumulate; operands split into fp16 hi + scaled fp16 lo, 3 f16 MMAs per product (log-likelihoods within 1e-3 of the fp64-accumulating oracle)"}

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
                                           "sample": "%d of the workload's streams (U(8,16) s), one stream per thread, %d threads, lattice -> MBR results" % (n, cores)},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    if a.inproc:
        return inproc(a)

    import torch
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=600))
    os.environ["VOSK_BATCH_DEVICES"] = str(local_rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        return reduce_over_ranks(x, "max", "cuda")

    if rank == 0:
        model_dir()  # generate once
    barrier()
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
    peaks = peaks_measured()

    # ---------------- value: device-resident, K steps back to back ----------------
    model.RunResident(audio_mat, lengths, passes=a.warmup)
    model.ResetStats()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, texts = model.RunResident(audio_mat, lengths, passes=a.steps)
    barrier()
    wall_resident = time.perf_counter() - t0
    clocks = sampler.stop()
    st = model.Stats()
    resident_bad = model.resident_mismatches
    # one extra pass with the pipeline slots serialized: per-stage device times (CUDA events on the launching
    # stream) free of cross-slot overlap; these are the durations the roofline uses
    model.ResetStats()
    model.SetTiming(True)
    model.SetSlots(1)
    model.RunResident(audio_mat, lengths)
    ser = model.Stats()
    model.SetSlots(64)
    model.SetTiming(False)
    dev_s = max_over_ranks(dev_ms / 1000.0)
    audio_total = reduce_over_ranks(audio_s, "sum", "cuda")   # every rank takes part in every collective
    value = audio_total * a.steps / dev_s

    # ---------------- e2e: through the C ABI with host buffers, K steps back to back ----------------
    pieces = [[w[i:i + 4000].tobytes() for i in range(0, len(w), 4000)] for w in waves]  # 8000-byte reads, as the reference driver
    e2e_run(vosk, model, pieces, 1)
    model.ResetStats()
    barrier()
    t0 = time.perf_counter()
    e2e_texts, e2e_bad = e2e_run(vosk, model, pieces, a.steps)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = audio_total * a.steps / e2e_s
    est = model.Stats()
    same = sum(1 for x, y in zip(texts, e2e_texts) if x == y)

    # ---------------- roofline of the dominant kernel (serialized pass: one step) ----------------
    roofs, search_bytes = stage_rooflines(ser, lengths, "small", peaks)
    kernel_ms = {k: v["ms"] for k, v in roofs.items()}
    dominant = max(kernel_ms, key=kernel_ms.get)
    kname = {"search": "decode_kernel", "tdnnf": "gemm_tc_kernel", "ivector": "ivector_kernel", "mfcc": "mfcc_kernel", "lattice_prune": "lattice_prune_kernel"}[dominant]
    roof = {"kernel": kname, "bound": roofs[dominant]["bound"], "achieved": roofs[dominant]["achieved"], "peak": roofs[dominant]["peak"],
            "unit": roofs[dominant]["unit"], "frac": roofs[dominant]["frac"], "traffic": None, "peak_src": peaks["src"],
            "algorithmic": roofs[dominant]["algorithmic"], "launch_ms": "sum of the kernel's launches of one serialized step: %.2f ms" % kernel_ms[dominant]}
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(kname)
        if ent:
            roof["traffic"] = ent["value"]
            roof["traffic_note"] = ent["unit"] + "; from " + tj["_source"]
            if dominant == "search":  # the same unit for the algorithmic side: bytes of one search step (512 lanes x 17 frames)
                roof["algorithmic_bytes_per_search_step"] = search_bytes / max(1.0, ser["steps"])
    except Exception:
        pass

    # ---------------- extras (rank 0) ----------------
    del model
    extras = {}
    partial_latency = None
    if rank == 0 and not a.no_extras:
        for name, o in (("best_path_mode (lattice=0: word-aligned best path, conf 1)", "lattice=0"),
                        ("device_lattice_only (lattice=2: link log + device pruning, no host chain; results are best-path texts)", "lattice=2")):
            try:
                extras[name], _ = resident_leg(vosk, mdir, opts + "," + o, audio_mat, lengths, 1, max(2, a.steps // 2))
            except Exception as e:  # the extras never take the headline line down
                extras[name] = {"error": str(e)[:200]}
        # a denser search frontier (SURVEY.md section 8d calls 3-7 k tokens per frame typical; the random-init network's posteriors are
        # peaky): the log-likelihoods are flattened by acoustic-scale 0.5, max-active 7000 binds on the heavy frames
        try:
            n_d = min(256, a.streams)
            dopts = ("num-channels=%d,max-batch-size=%d,max-seconds=18,acoustic-scale=0.5,log-tokens-per-frame=7168,log-links-per-frame=16384,"
                     "lat-link-cap=262144,lat-tok-cap=131072" % (n_d, n_d))
            # the search and the device lattice (lattice=2) on 256 streams; the whole result path (host chain: these lattices are ~4x the
            # default leg's and the pruned determinization grows faster than that) on the first 64 of them
            dleg, _ = resident_leg(vosk, mdir, dopts + ",lattice=2", audio_mat[:n_d], lengths[:n_d], 1, 2)
            dleg = {"value_device_lattice_only": dleg["value"], "ms_per_step_device_lattice_only": dleg["ms_per_step"], "unit": UNIT,
                    "lattice_links_logged_per_step": dleg["lattice_links_logged_per_step"], "lattice_arcs_after_pruning_per_step": dleg["lattice_arcs_after_pruning_per_step"],
                    "truncated": dleg["truncated"], "options": dopts}
            n_h = min(64, n_d)
            hleg, dtexts = resident_leg(vosk, mdir, dopts, audio_mat[:n_h], lengths[:n_h], 1, 1)
            dleg["whole_result_path_%d_streams" % n_h] = {k: hleg[k] for k in ("value", "ms_per_step", "results_with_confidence_below_1", "host_lattice_chain_cpu_ms_per_step",
                                                                                 "host_lattice_threads", "truncated", "lattice_fallbacks")}
            dm = vosk.BatchModel(mdir, options=dopts + ",lattice=2")
            dm.SetTiming(True)
            dm.SetSlots(1)
            dm.RunResident(audio_mat[:n_d], lengths[:n_d])
            dser = dm.Stats()
            del dm
            droofs, _ = stage_rooflines(dser, lengths[:n_d], "small", peaks)
            dframes = sum((int(1 + (n - 400) // 160) + 2) // 3 for n in lengths[:n_d])
            dleg.update({"streams": n_d, "tokens_per_frame": dser["tokens"] / dframes, "arcs_per_frame": (dser["arcs_emitting"] + dser["arcs_epsilon"]) / dframes,
                         "max_tokens_in_a_frame": dser["max_tokens_per_frame"], "roofline_search": droofs.get("search"), "roofline_lattice_prune": droofs.get("lattice_prune")})
            if not a.no_cpu_baseline:
                dps, _, _ = parity_sample(vosk, "small", waves, dtexts, 2, dopts.split("max-seconds=18,")[1], acoustic_scale=0.5)
                dleg["parity_sample"] = {k: v for k, v in dps.items() if "oracle_search" in k or "bench_run" in k or k == "streams"}
            extras["dense_frontier (acoustic-scale=0.5)"] = dleg
        except Exception as e:
            extras["dense_frontier (acoustic-scale=0.5)"] = {"error": str(e)[:300]}
        try:
            partial_latency = latency_pass(vosk, mdir, a.latency_streams, a.latency_seconds)
        except Exception as e:
            partial_latency = {"error": str(e)[:200]}

    cpu = None
    parity = None
    if rank == 0 and not a.no_cpu_baseline:
        try:
            parity, ca, cw = parity_sample(vosk, "small", waves, texts, 12)
            cpu = {"value": ca / cw, "unit": UNIT, "cores": 1, "kind": "port",
                   "sample": "12 of the workload's own streams (U(8,16) s) through the whole CPU restatement (features .. lattice -> MBR text) one after another "
                             "on one core (%.1f s of CPU work)" % cw}
        except Exception as e:
            parity = {"error": str(e)[:300]}

    large = None
    if not a.no_large:
        large = large_leg(a, rank, world, dist, barrier, peaks)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dev_s * 1000.0 / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(est["h2d_bytes"] / a.steps), "d2h_bytes_per_step": int(est["d2h_bytes"] / a.steps),
                        "transcripts_equal_to_resident_run": "%d/%d" % (same, len(texts)), "texts_differing_between_steps": e2e_bad,
                        "bytes": "counted by the engine at every host<->device copy of the timed steps (samples, lane descriptors, results, lattices)"},
                "gpu_launches": int(st["launches"]), "clocks": clocks, "roofline": roof, "roofline_all_stages": roofs, "cpu_baseline": cpu,
                "parity_sample": parity,
                "results": {"with_confidence_below_1": sum(1 for t in texts if '"conf" : 0.' in t), "nonempty": sum(1 for t in texts if '"text" : ""' not in t),
                            "texts_differing_between_steps": resident_bad, "truncated": st["truncated"], "lattice_fallbacks": st["lattice_fallbacks"]},
                "kernel_ms_per_step": kernel_ms,
                "roofline_note": "stage durations from one extra pass with the pipeline slots serialized (no overlap); CUDA events on the launching stream",
                "host_lattice_chain": {"cpu_ms_per_step": st["post_ms"] / a.steps, "threads": st["post_threads"],
                                       "note": "determinization / word alignment / MBR of the finished segments on the host lattice pool, beside the device work"},
                "host_wall_ms_per_step_resident": wall_resident * 1000.0 / a.steps,
                "search_counters_per_step": {"tokens": ser["tokens"], "arcs_emitting": ser["arcs_emitting"], "arcs_epsilon": ser["arcs_epsilon"],
                                             "tokens_new": ser["tokens_new"], "links_logged": ser["links"], "lattice_arcs_kept": ser["lattice_arcs"]},
                "audio_seconds_per_step": audio_total, "other_modes": extras, "partial_latency": partial_latency, "large": large}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
