// vb_common.h — shared host/device definitions of the B200 batch recognition engine.
//
// The engine replaces everything below the reference's BatchModel/BatchRecognizer
// [REF src/batch_model.cc], [REF src/batch_recognizer.cc] (Kaldi cudafeat / BatchedStaticNnet3 /
// CudaDecoder).  One Engine drives one GPU; streams occupy "channels" (persistent per-stream device
// state) and every engine step processes one chunk for each of up to max_lanes channels ("lanes").
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#define VB_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t e__ = (expr);                                                                        \
        if (e__ != cudaSuccess) {                                                                        \
            char b__[512];                                                                               \
            snprintf(b__, sizeof b__, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, \
                     cudaGetErrorString(e__));                                                           \
            throw std::runtime_error(b__);                                                               \
        }                                                                                                \
    } while (0)

namespace vb {

void log_msg(int level, const char *fmt, ...);  // level: -1 error, 0 info, >0 verbose (vosk_set_log_level semantics)
extern int g_log_level;

constexpr int kFrameLen = 400;    // 25 ms @ 16 kHz   [REF training/conf/mfcc.conf] + Kaldi defaults
constexpr int kFrameShift = 160;  // 10 ms
constexpr int kFftSize = 512;
constexpr int kNumMel = 40;
constexpr int kNumCeps = 40;
constexpr int kSubsample = 3;     // [REF src/batch_model.cc:82]
constexpr int kMaxNodes = 48;
constexpr int kMaxOffsets = 5;

// ---- engine configuration (defaults follow [REF src/batch_model.cc:69-88] where the reference has one) ----
struct Config {
    int frames_per_chunk = 51;  // [REF src/batch_model.cc:88]
    int max_lanes = 512;        // reference: max_batch_size = 32  [REF :70]; raised — a batch of 32 starves a B200
    int num_channels = 600;     // [REF :71]
    float beam = 13.0f;         // [REF :79]
    float lattice_beam = 6.0f;  // [REF :80]
    int max_active = 7000;      // [REF :78]
    int min_active = 200;       // Kaldi LatticeFasterDecoderConfig default
    float beam_delta = 0.5f;    // Kaldi default
    float acoustic_scale = 1.0f;  // [REF :81]
    int tok_cap = 65536;        // tokens per frame per channel
    int cand_cap = 262144;      // candidate records per frame (per resident CTA)
    int hash_size = 131072;     // open-addressing slots per resident CTA (power of two, >= 2*tok_cap)
    int max_seconds = 24;       // token-log capacity per channel (decoder frames = seconds * 100 / 3)
    int log_tokens_per_frame = 4096;  // average logged tokens per frame the token log is sized for
    int device = 0;
    int heavy_tokens = 2500;    // lanes above this many tokens per frame get the 1024-thread search CTAs,
    int mid_tokens = 900;       // ... lanes above this the 512-thread ones, the rest 256
    int heavy_threads = 1024, mid_threads = 512, light_threads = 256;
    int load_decay_percent = 80;  // a stream's load estimate = max(last chunk's token peak, this share of the previous estimate)
    float endpoint_rule5_seconds = 20.0f;  // reset_on_endpoint [REF src/batch_model.cc:72] with Kaldi's default endpoint rules and the
                                // reference's empty silence-phone list: only rule 5 (utterance length) can fire; 0 = never
    // Silence endpointing (Kaldi OnlineEndpointConfig rules 1-4), active when model.conf names --endpoint.silence-phones:
    // after every chunk the best path's trailing silence and the final relative cost are tested as kaldi::EndpointDetected does
    // [REF src/recognizer.cc:318], [REF src/model.cc:142-145]; defaults are Kaldi's.
    char endpoint_silence_phones[512] = "";
    int ep_must_contain_nonsilence[4] = {0, 1, 1, 1};
    float ep_min_trailing_silence[4] = {5.0f, 0.5f, 1.0f, 2.0f};
    float ep_max_relative_cost[4] = {INFINITY, 2.0f, 8.0f, INFINITY};  // inf = no limit (rules 1 and 4 also fire when no final state is active)
    float ep_min_utterance_length[4] = {0.f, 0.f, 0.f, 0.f};
    float mfcc_low_freq = 20.f, mfcc_high_freq = -400.f;  // conf/mfcc.conf [REF training/conf/mfcc.conf:4-5]; high <= 0: offset from Nyquist
    int fe_priority = 0;        // CUDA stream priority of the front-end pipe relative to the search pipe (1 / 0 / -1)
    int fe_split = 2;           // front-end chains of a full-width step (1-4): its launches are sub-wave, so the lanes are dealt to several chains side by side
    int device_resample = 1;    // resample non-16 kHz input on the GPU (0: on the host, in accept_waveform)
    int pipeline_slots = 4;     // lane groups in flight on separate CUDA streams
    int num_gselect = 5;        // ivector.conf
    float min_post = 0.025f, posterior_scale = 0.1f, max_count = 100.0f;  // [REF src/model.cc:257]
    int cmn_window = 600, global_frames = 200;
    int use_tensor_cores = 1;   // TDNN-F GEMMs on tcgen05, fp32 accumulate: 1 = fp16 hi/lo split (3 x f16 MMAs), 2 = TF32 hi/lo split; 0 = fp32 FFMA kernel
    int debug_capture = 0;      // allow per-stream capture of intermediates (tests)
    int lattice = 1;            // 1 (default, the reference's result path [REF src/batch_recognizer.cc:43-107,138-149]): link log +
                                // lattice_beam pruning on the device, raw lattice to the host chain (determinization, word
                                // alignment, MBR); 2: device lattice only, text from the best path; 0: best path only
    int model_conf = 0;         // 1: also take beam / lattice-beam / max-active / min-active / batch sizes and the silence
                                // endpointing rules from model.conf (the reference's batch path hard-codes them [REF src/batch_model.cc:69-88])
    int log_links_per_frame = 6144;   // average links per frame the link log is sized for
    int partials = 0;           // partial results: best path so far after every chunk (vosk_batch_recognizer_partial_result)
    int batcher_sleep = 1;      // the batcher thread waits for a step with short sleeps (not a spin) while the host lattice chain runs
    int post_threads = 0;       // host threads turning lattices into results (0 = hardware threads / (engines x local ranks))
    int num_engines = 1;        // engines sharing this host (set by BatchModel)
    int lat_tok_cap = 131072, lat_link_cap = 262144;  // pruned raw lattice of one stream (states / arcs)
};

// ---- acoustic model graph description (device-visible, passed by value to kernels) ----
struct NodeDesc {
    int dim;        // feature dimension of a row
    int step;       // time units between consecutive rows (1 or 3)
    int ring;       // ring size in rows (power of two)
    int t_start;    // first time index this node ever holds
    int cum_right;  // input frames (time units) of look-ahead this node needs beyond its own time
    float *buf;     // [num_channels][ring][dim]
};

struct OpDesc {
    int in_node, out_node, byp_node;  // byp_node = -1 if none
    int n_off;
    int offs[kMaxOffsets];
    int uses_ivec;  // append the lane's i-vector to the spliced input (tdnn1)
    int K, N;       // K = n_off*in_dim (+ ivec_dim)
    int relu, has_bn;
    const float *W;     // [N][K] fp32
    const void *W_hi;   // tensor-core operand split of W (vbk_split_weights): fp16 hi / TF32 hi
    const void *W_lo;   // fp16 (W - hi) * 2^11 / fp32 W - hi
    const float *bias;  // [N] or null
    const float *bn_scale, *bn_offset;
    float bypass_scale;
};

// per step, per lane (host -> device)
struct LaneDesc {
    int channel;
    int n_samples;     // new samples in this chunk
    int carry;         // samples carried from the previous chunk (device-resident)
    int frames_before; // MFCC frames that existed before this chunk
    int frames_after;  // ... after (F_k)
    int first, last;   // first / last chunk of the stream
    int iv_end_before; // spliced frames already accumulated into the i-vector stats
    int iv_end_after;
    int in_end_before; // node-0 timeline end (exclusive) before / after this chunk, incl. edge padding
    int in_end_after;
    int dec_frames_before;  // decoder frames consumed before this step
    int src_row;       // sample source: staging + src_row * src_stride + src_off   (int16 units)
    int src_off;
    int dec_first;     // first chunk of a SEGMENT: the search starts over (InitDecoding); features / i-vector carry on
    int dec_last;      // last chunk of a segment (endpoint or stream end): final pass, traceback, lattice, result
};

// per step, per node, per lane (device, written by the plan kernel)
struct NodeLane {
    int t_begin;  // first new time
    int n_rows;   // new rows this step
};

inline int num_frames_for(int64_t samples) {
    return samples < kFrameLen ? 0 : (int)(1 + (samples - kFrameLen) / kFrameShift);
}

}  // namespace vb
