// vb_engine.h — per-GPU batch engine: channels, chunk batching, pipeline launch, result hand-off.
//
// Replaces the objects BatchModel constructs from Kaldi — BatchedThreadedNnet3CudaOnlinePipeline and
// CudaOnlinePipelineDynamicBatcher [REF src/batch_model.cc:90-96] — and the calls BatchRecognizer makes on
// them: Push / SetLatticeCallback / GetNumPendingChunks / WaitForCompletion
// [REF src/batch_recognizer.cc:40,138-149,167,201], [REF src/batch_model.cc:118-121].
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "vb_common.h"
#include "vb_kernels.h"
#include "vb_model.h"

namespace vb {

// pruned raw lattice of one stream as the device hands it over (lattice=1): states are the surviving tokens in
// (frame, log) order, arcs refer to the decoding graph by csr arc id
struct RawLattice {
    int n_states = 0, start = -1, frames = 0, error = 0;
    std::vector<int> src, dst, arc;
    std::vector<float> acoustic;
    std::vector<int> final_state;
    std::vector<float> final_cost;
    std::vector<int> state_frame, state_graph;  // per lattice state: frame, graph state (graph state only with debug-capture)
};

// one finished lane's pruned lattice exactly as the device wrote it (copied off the pinned bounce buffer by the batcher
// thread; turned into a RawLattice where the result is made, on the lattice thread pool)
struct PackedLattice {
    int n_tok = 0, start = -1, frames = 0, error = 0;
    std::vector<int4> links;   // {src, dst, csr arc, acoustic cost bits}
    std::vector<int2> finals;  // {state, final cost bits}
    std::vector<int> tok_frame, tok_state;
    std::shared_ptr<RawLattice> unpack() const;
};

struct BestPath {
    std::shared_ptr<PackedLattice> packed;         // set when lattice generation is on
    mutable std::shared_ptr<RawLattice> lattice;   // unpacked on first use
    const RawLattice *raw_lattice() const {
        if (!lattice && packed) lattice = packed->unpack();
        return lattice.get();
    }
    std::vector<int> arcs;  // csr arc ids in path order
    float cost = 0.f;
    bool reached_final = false;
    int error = 0;
    int frames = 0;         // decoder frames
    float offset = 0.f;     // start of the segment in the stream, seconds (GetTimeOffsetSeconds)
    int seq = 0;            // index of the segment in its stream
};

// debug capture of one stream's intermediates (tests; enabled per stream before the first chunk)
struct Capture {
    std::vector<float> mfcc, ivectors, loglikes;
    std::vector<int> frame_off, tok_state, tok_arc, tok_prev;
    std::vector<float> tok_cost;
    int error = 0;
    std::shared_ptr<RawLattice> lattice;
    std::vector<int> lat_links;  // flattened [n][4] {src, dst, arc, acoustic bits}
    std::vector<int> lat_final;  // flattened [n][2] {state, final cost bits}
    std::vector<int> lat_hdr;    // {n_states, n_links, n_final, start, error, frames}
};

struct Stream {
    uint64_t id = 0;
    int channel = -1;
    bool started = false, finished = false, queued = false;
    bool in_flight = false;  // the stream's next chunk waits until the step of this one has completed (silence endpointing: the
                             // endpoint decision; a segment closed in mid-stream: its result still lives in the channel state)
    int64_t samples = 0;     // samples handed to the GPU so far
    int frames = 0;          // MFCC frames computed so far
    int iv_end = 0, in_end = 0, dec_frames = 0, carry = 0;
    int load = 0;            // tokens per frame seen in the stream's last step (batching key)
    int seg_start = 0;       // decoder frames before the current segment
    int seg_index = 0;       // segments closed so far
    bool seg_open = false;   // the search of the current segment has been initialised
    int last_tier = -1;      // search tier (pipe) that ran the stream's previous chunk
    bool resident = false;   // samples are read from a device-resident matrix
    int resident_row = 0;    // its row there
    struct Chunk {
        std::vector<int16_t> samples;
        int n_resident = 0;  // chunk length when the samples live in a device-resident matrix
        // device-side resampling (stream opened at another rate): the chunk's n_out 16 kHz samples are produced on the GPU
        // from the raw input-rate samples of the accept_waveform calls that overlap it (vb_kernels.h ResampleSeg)
        struct Seg { int raw_off, in_base, n_in, out_first, out_pos, n_out; };
        int rate = 0;        // input rate; 0 = samples[] already is 16 kHz
        int n_out = 0;
        std::vector<int16_t> raw;
        std::vector<Seg> segs;
        bool last;
        bool close_segment = false;  // an empty chunk injected after a silence endpoint: final pass, traceback and result of the segment
        std::chrono::steady_clock::time_point t_push;  // when the chunk's last sample was accepted
    };
    // partial result (partials=1): output labels of the best path so far, oldest first
    std::mutex partial_mu;
    std::vector<int> partial_words;
    int partial_frames = 0;  // decoder frames the partial covers
    std::deque<Chunk> pending;            // guarded by Engine::mu_
    std::atomic<int> pending_chunks{0};
    std::function<void(const BestPath &)> on_result;  // called on the engine worker thread
    std::unique_ptr<Capture> capture;
};

struct StepStats {
    double audio_seconds = 0;
    long long steps = 0, lanes = 0, launches = 0;
    unsigned long long tok = 0, arc_e = 0, arc_eps = 0, tok_new = 0;
    unsigned long long lane_cycles_sum = 0, lane_cycles_max = 0, max_tokens = 0, lane_launches = 0;
    unsigned long long arcs_staged = 0, links = 0, lat_arcs = 0;
    unsigned long long phase_slowest[24] = {};  // the same phases for the slowest lane-launch of each search tier (1024 / 512 / 256 threads)
    unsigned long long tier_slowest_cycles[3] = {}, tier_slowest_tokens[3] = {}, tier_lane_launches[3] = {};
    unsigned long long phase[16] = {};  // search cycles per phase (cutoff, rank, log, gather, insert, closure, finalize, -) of the heavy / light CTAs  // arcs parked below the running cutoff, links logged, lattice arcs kept
    double t_feat = 0, t_ivec = 0, t_nnet = 0, t_dec = 0, t_total = 0;  // device ms (only when timing enabled)
    long long dec_launches = 0, gemm_launches = 0, resample_segments = 0;
    double host_launch_ms = 0;  // host time spent enqueueing steps
    long long truncated = 0;          // results delivered although a device capacity (tokens / candidates / log / links / lattice) overflowed
    long long lattice_fallbacks = 0;  // lattice-mode results that fell back to the best path (no lattice, capacity error, chain failure)
    unsigned long long h2d_bytes = 0, d2h_bytes = 0;  // bytes of the host<->device copies of the steps (samples, descriptors, results, lattices)
    double t_prune = 0;               // device ms of the lattice pruning launches (only when timing enabled)
    double host_complete_ms = 0, host_fetch_ms = 0;  // batcher-thread milliseconds spent completing steps / of that, fetching lattices
    double post_ms = 0;               // host milliseconds the lattice pool spent on results (summed over its threads)
    long long post_jobs = 0;
};

class Engine {
   public:
    Engine(const Model &model, const Config &cfg);
    ~Engine();
    Engine(const Engine &) = delete;

    int samples_per_chunk() const { return cfg_.frames_per_chunk * kFrameShift; }
    const Config &config() const { return cfg_; }
    std::shared_ptr<Stream> open_stream();
    // copies the samples; n <= samples_per_chunk unless it is the (possibly empty) last chunk
    void push(const std::shared_ptr<Stream> &s, const int16_t *samples, int n, bool last);
    void push_chunk(const std::shared_ptr<Stream> &s, Stream::Chunk &&ch);  // a prepared chunk (16 kHz samples, or raw segments to resample)
    // limits of one device-resampled chunk (a chunk beyond them is resampled on the host by the recognizer)
    static constexpr int kMaxResampleSegs = 64;
    int max_resample_raw() const { return samples_per_chunk() * 3 + 4096; }
    void wait();  // until every chunk pushed so far is decoded and its result delivered
    StepStats stats();
    // latency from a chunk's acceptance to its step's results (partial / final) being available: {p50, p90, p99, mean, count} in ms
    void latency(double *out5, bool reset);
    void reset_stats();
    void set_timing(bool on) { timing_ = on; }
    void count_fallback() { lattice_fallbacks_.fetch_add(1); }
    int post_thread_count() const { return (int)post_threads_.size(); }
    // limits the number of pipeline slots in use (1 = fully serialized steps: per-kernel timings without overlap)
    void set_active_slots(int n) { active_slots_ = std::max(1, std::min(n, (int)slots_.size())); }

    // Device-resident run for kernel-level benchmarking: `audio` holds num_streams x samples int16 already in
    // HBM; processes every stream chunk by chunk with no host<->device sample traffic.  Returns device ms.
    // passes > 1: the streams are decoded that many times over, back to back (see the definition); out = results of the last pass.
    double run_resident(const int16_t *d_audio, int num_streams, int stride, const int *lengths, std::vector<BestPath> *out, int passes = 1);
    // called for every finished segment of a resident run, on the thread that delivers results (pass, stream index, result)
    std::function<void(int, int, const BestPath &)> resident_hook;

   private:
    struct Lane {
        std::shared_ptr<Stream> s;
        Stream::Chunk chunk;
        int dec_frames_after = 0;
        bool seg_end = false;    // this chunk closes a segment: a result is due
        int seg_start = 0;       // decoder frames before the segment this chunk belongs to
        bool endpoint = false;   // set on completion: a silence endpoint was detected after this chunk
        bool holds = false;      // this lane holds its stream back (Stream::in_flight) until the step has completed
        float seg_offset = 0.f;
        int seg_index = 0;
    };
    // One pipeline slot = the buffers of one engine step (one chunk for each of up to max_lanes streams).  Steps flow
    // through two in-order pipes: the front end (H2D, features, i-vector, TDNN-F) on fe_stream_ and the search
    // (beam search, lattice pruning, D2H of results) on dec_stream_.  The search of step s only waits for the front end
    // of step s, so the front end of step s+1 runs beside it; a stream may have one chunk in every slot, and a slot is
    // reused only after its step has completed, which also bounds how far the front end runs ahead of the search in
    // the log-likelihood rings.
    struct Slot {
        cudaStream_t stream = nullptr;  // utility stream (debug taps, lattice fetch) of the slot
        cudaEvent_t fe_done = nullptr, dec_done = nullptr, fork = nullptr, join[3] = {}, tier_done[3] = {};
        int *d_queue = nullptr;  // [2] lane queues of the two search launches
        cudaEvent_t ev[8] = {};  // [6], [7]: around the lattice pruning launch
        bool pruned = false;
        cudaEvent_t done = nullptr;
        int16_t *d_staging = nullptr, *h_staging = nullptr;
        int16_t *d_raw = nullptr, *h_raw = nullptr;          // raw input-rate samples of the step's resampled lanes (allocated on first use)
        ResampleSeg *d_segs = nullptr, *h_segs = nullptr;
        LaneDesc *d_lanes = nullptr, *h_lanes = nullptr;
        NodeLane *d_table = nullptr;
        int *d_rowoff = nullptr;
        int2 *d_rows = nullptr;  // [nodes][rows_cap_] packed row table of the step
        int *d_rowoff2[3] = {};  // the same two for the other shares of the lanes when the front end runs as several chains
        int2 *d_rows2[3] = {};
        DecArgs dec{};
        DecChannelState *h_cs = nullptr;
        int *h_path = nullptr;
        int *d_load = nullptr, *h_load = nullptr;
        int *h_partial = nullptr;  // [L][kPartialCap] words + [L] counts
        int *h_endp = nullptr;     // [L] trailing silence frames + [L] final relative cost (float bits)
        LatHeader *h_lat_hdr = nullptr;
        char *h_lat_pool = nullptr;  // pinned bounce buffer the finished lanes' lattices are copied through, many per synchronize
        bool failed = false;         // the launch threw: nothing of this step can be read back
        std::vector<Lane> lanes;
        bool busy = false, timed = false;
        double audio = 0;
        long long launches = 0, gemms = 0, resample_segs = 0;
        unsigned long long h2d = 0, d2h = 0;  // bytes copied for this step
    };
    int resample_table(int rate);  // index of the device phase table of an input rate (built on first use)
    void worker();
    void launch_step(Slot &sl, const int16_t *d_resident, int resident_stride);
    void complete_step(Slot &sl);
    void upload_model();
    void alloc_state();
    void finish_lane(Slot &sl, Lane &ln, int lane_idx, std::shared_ptr<PackedLattice> lat);
    void fetch_lattices(Slot &sl, std::vector<std::shared_ptr<PackedLattice>> *out);  // all finished lanes of the step, indexed by lane position

    const Model &model_;
    Config cfg_;
    cudaStream_t stream_ = nullptr;  // setup / utility stream
    cudaStream_t fe_stream2_[3] = {};     // the other front-end chains of a step (fe-split)
    float post_backlog_ema_ = 0.f;        // share of the recent launches that found the host lattice pool behind
    cudaStream_t fe_stream_ = nullptr, dec_stream_ = nullptr, dec_stream2_ = nullptr, dec_stream3_ = nullptr, post_stream_ = nullptr;  // the two pipes (+ the light-lane search launch)
    std::vector<Slot> slots_;
    bool timing_ = false;
    std::atomic<int> active_slots_{1};
    // model on device
    FeatTables feat_tab_{};
    IvecModel iv_model_{};
    std::vector<NodeDesc> nodes_;
    std::vector<OpDesc> ops_;
    struct MapPair { alignas(64) unsigned char hi[128]; alignas(64) unsigned char lo[128]; };
    std::vector<MapPair> maps_;
    NodeDesc *d_nodes_ = nullptr;
    GraphDev graph_{};
    std::vector<void *> allocs_;
    bool endpointing_ = false;  // silence rules 1-4 active (model.conf names silence phones)
    std::map<int, int> resample_ids_;
    std::vector<ResampleTable> resample_tables_;
    ResampleTable *d_resample_tables_ = nullptr;
    // per channel / per step state
    IvecState iv_state_{};
    int *d_iv_sel_g_ = nullptr;    // scratch between the i-vector frame kernel and its statistics kernel (front-end stream order)
    float *d_iv_sel_w_ = nullptr, *d_iv_fu_ = nullptr;
    int iv_frames_cap_ = 0;
    int16_t *d_carry_ = nullptr;
    int *d_node_end_ = nullptr;
    DecArgs dec_{};  // template: graph, options and per-channel arrays; each slot adds its own scratch
    float *d_capture_ = nullptr, *h_capture_ = nullptr;
    size_t capture_floats_ = 0;
    int tier_scratch_[3] = {0, 0, 0};            // first scratch index of each search tier
    cudaEvent_t last_tier_done_[3] = {};         // most recent completion event recorded on each tier pipe
    int max_in_rows_ = 0, log_cap_ = 0, max_frames_ = 0, path_cap_ = 0, slot_lanes_ = 0, link_cap_ = 0, rows_cap_ = 0;
    std::vector<int> free_channels_;
    // batching
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::deque<std::shared_ptr<Stream>> ready_;
    const int16_t *resident_audio_ = nullptr;  // set for the duration of run_resident
    int resident_stride_ = 0;
    long long outstanding_ = 0;  // chunks pushed and not yet completed
    bool stop_ = false;
    std::thread thread_;
    // lattice post-processing pool (lattice=1): determinization / alignment / MBR of finished segments run here, off the
    // batcher thread, as the reference pipeline does with its lattice thread pool
    std::vector<std::thread> post_threads_;
    std::deque<std::function<void()>> post_queue_;
    std::mutex post_mu_;
    std::condition_variable post_cv_;
    long long post_outstanding_ = 0;  // guarded by mu_ (wait() watches it)
    bool post_stop_ = false;
    void post_worker();
    uint64_t next_id_ = 0;
    StepStats stats_;
    std::vector<float> latencies_ms_;  // guarded by stats_mu_
    std::atomic<long long> lattice_fallbacks_{0};
    static constexpr size_t kLatPoolBytes = (size_t)96 << 20;
    std::mutex stats_mu_;
};

}  // namespace vb
