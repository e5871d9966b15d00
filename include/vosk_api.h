/* vosk_api.h — the reference-facing C ABI of the B200 batch recognition engine.
 *
 * This is the drop-in boundary: every function below has exactly the name, signature and
 * ownership rules of the reference's libvosk [REF src/vosk_api.h:287-346], so the reference's
 * cffi package (cdef generated from the header, dlopen of libvosk.so next to the package —
 * [REF python/vosk_builder.py:6-11], [REF python/vosk/__init__.py:17-32,185-235]) and any other
 * binding that calls the batch API load this library unchanged.  Only the batch path is
 * exported: the CPU recognizer, speaker model and grammar functions of the reference header
 * [REF src/vosk_api.h:58-285] are outside the accelerated path and are not defined
 * here — an ABI-mode binding resolves symbols lazily, so their absence only shows if called.
 *
 * All functions are extern "C"; handles are opaque; no exception crosses the boundary.
 */
#ifndef VOSK_API_H
#define VOSK_API_H

#ifdef __cplusplus
extern "C" {
#endif

/* Opaque handle: one loaded model plus one engine per selected GPU.
 * Replaces the object behind [REF src/vosk_api.h:45]. */
typedef struct VoskBatchModel VoskBatchModel;

/* Opaque handle: one audio stream.  Replaces the object behind [REF src/vosk_api.h:50]. */
typedef struct VoskBatchRecognizer VoskBatchRecognizer;

/* [REF src/vosk_api.h:294]  0 = info and errors, < 0 = errors only, > 0 = verbose. */
void vosk_set_log_level(int log_level);

/* [REF src/vosk_api.h:301]  One-time device initialisation; creates the CUDA context of the
 * selected device(s).  Safe to call more than once. */
void vosk_gpu_init();

/* [REF src/vosk_api.h:308]  Per-thread initialisation; a no-op here (the engine owns its threads). */
void vosk_gpu_thread_init();

/* [REF src/vosk_api.h:313]  Loads "model/..." relative to the current directory, exactly like
 * BatchModel::BatchModel() [REF src/batch_model.cc:28-37,76-77] (override: env VOSK_BATCH_MODEL_PATH).
 * Returns NULL on failure (missing files, no usable GPU); the reason is logged. */
VoskBatchModel *vosk_batch_model_new();

/* [REF src/vosk_api.h:316]  Free after every recognizer of the model has been freed. */
void vosk_batch_model_free(VoskBatchModel *model);

/* [REF src/vosk_api.h:319]  Blocks until every chunk pushed so far is decoded and its result is
 * visible to vosk_batch_recognizer_front_result. */
void vosk_batch_model_wait(VoskBatchModel *model);

/* [REF src/vosk_api.h:323]  New stream; audio of any sample rate is resampled to 16 kHz. */
VoskBatchRecognizer *vosk_batch_recognizer_new(VoskBatchModel *model, float sample_rate);

/* [REF src/vosk_api.h:326] */
void vosk_batch_recognizer_free(VoskBatchRecognizer *recognizer);

/* [REF src/vosk_api.h:329]  data = little-endian int16 mono PCM, length in BYTES; copied before
 * returning; never blocks on the GPU. */
void vosk_batch_recognizer_accept_waveform(VoskBatchRecognizer *recognizer, const char *data, int length);

/* [REF src/vosk_api.h:334]  Non-zero: results are NLSML XML instead of JSON. */
void vosk_batch_recognizer_set_nlsml(VoskBatchRecognizer *recognizer, int nlsml);

/* [REF src/vosk_api.h:337]  Flushes buffered audio and marks the end of the stream. */
void vosk_batch_recognizer_finish_stream(VoskBatchRecognizer *recognizer);

/* [REF src/vosk_api.h:340]  Oldest undelivered result, or "" (never NULL) when there is none.
 * The string belongs to the library and stays valid until pop / free. */
const char *vosk_batch_recognizer_front_result(VoskBatchRecognizer *recognizer);

/* [REF src/vosk_api.h:343]  Drops the oldest result; no-op when empty. */
void vosk_batch_recognizer_pop(VoskBatchRecognizer *recognizer);

/* [REF src/vosk_api.h:346]  Chunks of this stream queued but not yet decoded. */
int vosk_batch_recognizer_get_pending_chunks(VoskBatchRecognizer *recognizer);

#ifdef __cplusplus
}
#endif

#endif /* VOSK_API_H */
