/* vosk_b200.h — additive C surface of the B200 engine (nothing here exists in the reference).
 *
 * The reference's batch ABI has no model-path argument [REF src/vosk_api.cc:198-205], no partial
 * result (only a disabled callback [REF src/batch_recognizer.cc:120-137]) and no way to observe
 * intermediates.  These entry points add exactly that, for parity tests and measurement; existing
 * bindings never see them.
 */
#ifndef VOSK_B200_H
#define VOSK_B200_H
#include <stdint.h>

#include "vosk_api.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Like vosk_batch_model_new but with an explicit model directory and "key=value,key=value"
 * options: frames-per-chunk, max-batch-size, num-channels, beam, lattice-beam, max-active,
 * min-active, tok-cap, cand-cap, hash-size, max-seconds, log-tokens-per-frame, tensor-cores, lattice,
 * log-links-per-frame, lat-tok-cap, lat-link-cap, post-threads, partials, pipeline-slots, heavy-tokens, device-resample, model-conf,
 * fe-split (front-end chains of a full-width step, 1-4, default 2), debug-capture, devices (GPU indices separated by ':' or "all").  tensor-cores: 1 = fp16 hi/lo operand split (default),
 * 2 = TF32 hi/lo split, 0 = fp32 FFMA kernel.
 * Env VOSK_BATCH_OPTIONS / VOSK_BATCH_DEVICES are applied first.  NULL on failure. */
VoskBatchModel *vosk_batch_model_new_ex(const char *model_dir, const char *options);

/* Text of the last error raised while constructing a model on this thread ("" if none). */
const char *vosk_b200_last_error(void);

/* Samples per chunk (frames-per-chunk x 160), as BatchModel::samples_per_chunk_ [REF src/batch_model.cc:98]. */
int vosk_batch_model_samples_per_chunk(VoskBatchModel *model);

/* Cumulative counters since creation / the last reset, summed over the model's engines:
 * [0] audio seconds, [1] steps, [2] lanes, [3] kernel launches, [4] tokens expanded,
 * [5] emitting arcs, [6] epsilon arcs, [7] tokens created, [8..11] device ms features / i-vector /
 * network / search (only when timing is on), [12] GEMM launches, [13] sum and [14] max of the SM cycles
 * one lane spent in a search launch, [15] largest token count of a frame, [16] lane-launches, [17] host ms spent enqueueing steps,
 * [18] arcs parked below the running cutoff, [19] lattice links logged, [20] lattice arcs kept after pruning,
 * [21..28] / [29..36] SM cycles the 1024-thread / smaller search CTAs spent per phase (cutoff, rank, log, gather,
 * insert, closure, finalize, unused), [37] per-call segments resampled on the device, [38] results delivered although a device
 * capacity overflowed (logged; the result may be truncated), [39] lattice-mode results that fell back to the best path,
 * [40] host milliseconds spent by the lattice thread pool, [41] its jobs, [42] its threads, [43] device ms of the lattice
 * pruning launches (timing on), [44] batcher-thread ms spent completing steps, [45] of that, fetching lattices, [46] / [47] bytes the steps copied host-to-device /
 * device-to-host (samples, descriptors, results, lattices), [48..71] the per-phase SM cycles of the slowest lane-launch of each search
 * tier (1024 / 512 / 256 threads), [72..74] its cycles, [75..77] its largest token count, [78..80] lane-launches per tier.
 * Returns the number written. */
int vosk_batch_model_stats(VoskBatchModel *model, double *out, int n);
void vosk_batch_model_reset_stats(VoskBatchModel *model);
void vosk_batch_model_set_timing(VoskBatchModel *model, int on);
/* Number of pipeline slots (lane groups in flight) to use, 1..pipeline-slots; 1 serializes the steps so that
 * the per-stage device times of vosk_batch_model_stats are free of overlap. */
void vosk_batch_model_set_slots(VoskBatchModel *model, int n);

/* Device-resident run (kernel-level measurement): uploads the num_streams x samples_per_stream int16
 * matrix to HBM (untimed), then decodes all streams in lockstep with no host<->device sample traffic
 * and returns the elapsed device time in milliseconds (CUDA events), < 0 on error.  lengths (may be
 * NULL = every stream uses the full row) gives the number of valid samples of each row.
 * Result texts are kept until the next call; fetch with vosk_batch_model_resident_result. */
double vosk_batch_model_run_resident(VoskBatchModel *model, const int16_t *audio, int num_streams, int samples_per_stream,
                                     const int *lengths);
/* The same with the streams decoded `passes` times over, back to back, as new streams that are all queued at once: a stream
 * of pass p+1 starts as soon as a channel is free and the lattice chain of pass p's results runs beside the search of pass
 * p+1 (continuous serving); returns when every result of every pass has been delivered.  The texts kept are the last
 * pass's; *mismatches (may be NULL) = streams of earlier passes whose text differs from it (0 unless something is wrong). */
double vosk_batch_model_run_resident_passes(VoskBatchModel *model, const int16_t *audio, int num_streams, int samples_per_stream,
                                            const int *lengths, int passes, int *mismatches);
const char *vosk_batch_model_resident_result(VoskBatchModel *model, int stream);

/* Partial result (model option partials=1): the best path so far, without final costs, in the CPU API's text layout
 * {"partial" : "..."} [REF src/recognizer.cc:795-802]; the reference's batch path only has a disabled best-path callback
 * [REF src/batch_recognizer.cc:120-137].  The string is owned by the recognizer and valid until the next call.
 * partial_frames: decoder frames (30 ms) the latest partial covers. */
const char *vosk_batch_recognizer_partial_result(VoskBatchRecognizer *recognizer);
int vosk_batch_recognizer_partial_frames(VoskBatchRecognizer *recognizer);
/* Latency from the acceptance of a chunk's last sample to the moment its step's partial / final result became
 * retrievable, over the chunks completed since the last reset: out5 = {p50, p90, p99, mean (ms), count}. */
int vosk_batch_model_latency(VoskBatchModel *model, double *out5, int reset);

/* Test taps.  Enable before the first accept_waveform on a model created with debug-capture=1;
 * after finish_stream + wait, fetch "mfcc" [T][40], "ivectors" [chunks][D], "loglikes" [N][pdfs] (f32),
 * "frame_off" [N+2], "tok_state", "tok_arc", "tok_prev" (i32), "tok_cost" (f32), "error" (i32);
 * with lattice=1 also "lat_hdr" {states, links, finals, start, error, frames}, "lat_links" [n][4] {src, dst, arc, acoustic
 * cost bits}, "lat_final" [n][2] {state, final cost bits}, "lat_tok_frame", "lat_tok_state" (i32).
 * Returns the size in bytes (copies min(size, cap_bytes)), or -1 for an unknown name. */
void vosk_batch_recognizer_debug_capture(VoskBatchRecognizer *recognizer);
int64_t vosk_batch_recognizer_debug_get(VoskBatchRecognizer *recognizer, const char *what, void *out, int64_t cap_bytes);

/* Native feeder (measurement helper): n streams of int16 16 kHz samples go through vosk_batch_recognizer_new /
 * accept_waveform (bytes_per_call per call, round robin as the reference's driver [REF python/example/test_gpu_batch.py:27-51]) /
 * finish_stream from `threads` host threads, then vosk_batch_model_wait and front_result / pop.  results (may be NULL)
 * receives one malloc'd string per stream (a stream's segment texts concatenated; free with vosk_b200_free).  0 = ok. */
int vosk_b200_feed_streams(VoskBatchModel *model, const int16_t *const *samples, const int *lengths, int n, int bytes_per_call, int threads,
                           char **results);
/* The same with the streams fed `passes` times over (new recognizers every pass) before the one vosk_batch_model_wait; results =
 * the last pass's texts, *mismatches (may be NULL) = streams of earlier passes whose text differs from it. */
int vosk_b200_feed_streams_passes(VoskBatchModel *model, const int16_t *const *samples, const int *lengths, int n, int bytes_per_call, int threads,
                                  int passes, char **results, int *mismatches);
void vosk_b200_free(void *p);

/* Host-only hooks (no GPU needed; used by the CPU test suite).
 * device_for_stream: the utterance -> GPU sharding rule (stream id modulo the number of engines).
 * format_result: the engine's result-text writer on explicit words (frames are 30 ms decoder frames).
 * resample: the per-call resampler of accept_waveform (Kaldi LinearResample, flush=true) to 16 kHz.
 * model_check: loads a model directory with the engine's loaders and prints a one-line summary. */
int vosk_b200_device_for_stream(unsigned long long stream_id, int num_devices);
int vosk_b200_format_result(const char *const *words, const int *begin, const int *end, const float *conf, int n, float offset, char *out, int cap);
int vosk_b200_resample(const float *in, int n, float rate_in, float *out, int cap);
int vosk_b200_model_check(const char *model_dir, char *out, int cap);
/* model_tensor: loads a model directory (Kaldi formats or the generator's container, file by file) with the engine's
 * loaders and returns one loaded object as doubles: "meta" {ops, context, pdfs, ivector dim, bypass scale, prior offset,
 * gaussians}, "op<i>.meta" {in node, bypass node, uses ivector, relu+batchnorm, K, N, number of offsets, offsets...},
 * "op<i>.w" [N][K], "op<i>.b", "op<i>.bn_scale", "op<i>.bn_offset", "tid2pdf", "tid2phone", "iv.lda", "iv.gconsts",
 * "iv.weights", "iv.means_invvars", "iv.inv_vars", "iv.M", "iv.sigma_inv", "iv.cmvn".  Returns the element count (copies
 * min(count, cap)), 0 for an absent optional object, -1 on error (text in vosk_b200_last_error). */
int64_t vosk_b200_model_tensor(const char *model_dir, const char *name, double *out, int64_t cap);
/* lattice_result: the host lattice chain (phone-pruned determinization, graph scale 0.9, word alignment, MBR) on an
 * explicit raw lattice {states, start, links {src, dst, csr arc, acoustic cost}, finals}; stage 0 = result text, 3 = NLSML
 * text, 1 = determinized lattice, 2 = word-aligned lattice as text lines ("S start", "A src dst word graph acoustic tids",
 * "F state graph acoustic tids"), 11 / 12 = the same with the phone pass switched off, 4 = sizes and best-of-12 timings of
 * the chain's stages as JSON.  Returns the text length (copies at most cap-1 bytes), -1 on error. */
int vosk_b200_lattice_result(const char *model_dir, int n_states, int start, int n_links, const int *src, const int *dst, const int *arc,
                             const float *acoustic, int n_final, const int *final_state, const float *final_cost, float lattice_beam,
                             int stage, char *out, int cap);

#ifdef __cplusplus
}
#endif
#endif /* VOSK_B200_H */
