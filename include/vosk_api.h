/* vosk_api.h — the reference-facing C ABI of the B200 batch recognition engine.
 *
 * This is the drop-in boundary: every function below has exactly the name, signature and
 * ownership rules of the reference's libvosk [REF src/vosk_api.h:287-346], so the reference's
 * cffi package (cdef generated from the header, dlopen of libvosk.so next to the package —
 * [REF python/vosk_builder.py:6-11], [REF python/vosk/__init__.py:17-32,185-235]) and any other
 * binding load this library unchanged.  The header declares the WHOLE surface of the reference header, so that
 * eager binders resolve every symbol (JNA's Native.register [REF java/lib/src/main/java/org/vosk/LibVosk.java:38-41],
 * cgo, P/Invoke): the batch / GPU / log functions are the accelerated path and are implemented; the CPU recognizer,
 * speaker-model and grammar functions [REF src/vosk_api.h:58-285] are outside it and are exported as
 * stubs that log "CPU API not built" and return NULL / -1 / "" — the mirror image of how the reference stubs its batch
 * half when it is built without CUDA [REF src/vosk_api.cc:198-282].
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

/* ---- CPU recognizer API of the reference: declared and exported, NOT implemented (stubs; see the header comment) ---- */
typedef struct VoskModel VoskModel;            /* [REF src/vosk_api.h:30] */
typedef struct VoskSpkModel VoskSpkModel;      /* [REF src/vosk_api.h:35] */
typedef struct VoskRecognizer VoskRecognizer;  /* [REF src/vosk_api.h:41] */

VoskModel *vosk_model_new(const char *model_path);                        /* [REF :58]  stub: NULL */
void vosk_model_free(VoskModel *model);                                   /* [REF :66]  stub: no-op */
int vosk_model_find_word(VoskModel *model, const char *word);             /* [REF :74]  stub: -1 */
VoskSpkModel *vosk_spk_model_new(const char *model_path);                 /* [REF :81]  stub: NULL */
void vosk_spk_model_free(VoskSpkModel *model);                            /* [REF :89]  stub: no-op */
VoskRecognizer *vosk_recognizer_new(VoskModel *model, float sample_rate); /* [REF :100] stub: NULL */
VoskRecognizer *vosk_recognizer_new_spk(VoskModel *model, float sample_rate, VoskSpkModel *spk_model);  /* [REF :115] stub: NULL */
VoskRecognizer *vosk_recognizer_new_grm(VoskModel *model, float sample_rate, const char *grammar);      /* [REF :137] stub: NULL */
void vosk_recognizer_set_spk_model(VoskRecognizer *recognizer, VoskSpkModel *spk_model);                /* [REF :146] stub */
void vosk_recognizer_set_max_alternatives(VoskRecognizer *recognizer, int max_alternatives);            /* [REF :162] stub */
void vosk_recognizer_set_words(VoskRecognizer *recognizer, int words);                                  /* [REF :198] stub */
void vosk_recognizer_set_partial_words(VoskRecognizer *recognizer, int partial_words);                  /* [REF :204] stub */
void vosk_recognizer_set_nlsml(VoskRecognizer *recognizer, int nlsml);                                  /* [REF :209] stub */
int vosk_recognizer_accept_waveform(VoskRecognizer *recognizer, const char *data, int length);          /* [REF :221] stub: -1 */
int vosk_recognizer_accept_waveform_s(VoskRecognizer *recognizer, const short *data, int length);       /* [REF :226] stub: -1 */
int vosk_recognizer_accept_waveform_f(VoskRecognizer *recognizer, const float *data, int length);       /* [REF :231] stub: -1 */
const char *vosk_recognizer_result(VoskRecognizer *recognizer);           /* [REF :250] stub: "" */
const char *vosk_recognizer_partial_result(VoskRecognizer *recognizer);   /* [REF :264] stub: "" */
const char *vosk_recognizer_final_result(VoskRecognizer *recognizer);     /* [REF :273] stub: "" */
void vosk_recognizer_reset(VoskRecognizer *recognizer);                   /* [REF :279] stub */
void vosk_recognizer_free(VoskRecognizer *recognizer);                    /* [REF :285] stub */

/* ---- the accelerated path ---- */

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
