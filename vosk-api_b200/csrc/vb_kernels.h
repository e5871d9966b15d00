// vb_kernels.h — the C-ABI CUDA layer: plain-pointer launch entry points of the sm_100a kernels.
// Host code (vb_engine.cc) talks to the GPU only through these functions and the CUDA runtime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vb_common.h"

namespace vb {

// ---------------- K1: fused framing / FFT / mel / DCT / lifter ----------------
struct FeatTables {          // device pointers, built once per model
    const float *window;     // [400]  Povey window
    const float *twiddle;    // [256][2] cos,sin(-2*pi*k/512)
    const int *mel_start;    // [40]
    const int *mel_len;      // [40]
    const float *mel_w;      // [40][kMelMaxLen]
    const float *dct_t;      // [40 mel][40 cep]  (transposed DCT-II)
    const float *lifter;     // [40]
};
constexpr int kMelMaxLen = 64;
constexpr int kCarryMax = 400;

struct FeatArgs {
    const LaneDesc *lanes;   // device
    int num_lanes;
    const int16_t *staging;  // sample source (pinned-copy staging rows, or a device-resident audio matrix)
    long long src_stride;    // int16 elements between source rows
    int samples_per_chunk;
    int16_t *carry;          // [num_channels][kCarryMax]
    NodeDesc in_node;        // node 0 (MFCC ring, padded timeline)
    int context;
    FeatTables tab;
};

// ---------------- K0: input resampling to 16 kHz (streams opened at another rate) ----------------
// Kaldi LinearResample with flush=true per accept_waveform call [REF src/batch_recognizer.cc:27-29,157-158]: every call is
// filtered on its own (zeros outside the call).  A chunk of 16 kHz samples is therefore a concatenation of segments, each
// the share one call contributes; the host stages the raw input-rate samples the segment's taps can reach.
struct ResampleTable {       // device pointers; one per distinct input rate
    int in_unit, out_unit, max_taps;
    const int *first_index;  // [out_unit]
    const int *n_taps;       // [out_unit]
    const float *weights;    // [out_unit][max_taps]
};
struct ResampleSeg {
    int lane;       // staging row the samples go to
    int raw_off;    // first staged raw sample of the segment in the step's raw buffer
    int in_base;    // call-relative index of that sample
    int n_in;       // samples of the whole call (taps beyond it read zero)
    int out_first;  // call-relative index of the segment's first output sample
    int out_pos;    // position of that sample in the lane's chunk
    int n_out;
    int table;
};
struct ResampleArgs {
    const ResampleSeg *segs;  // device
    int num_segs;
    const int16_t *raw;       // device: raw input-rate samples of the step
    const ResampleTable *tables;
    int16_t *staging;         // [lanes][samples_per_chunk]
    int samples_per_chunk;
};

// ---------------- K1c: online CMN + i-vector ----------------
struct IvecModel {           // device pointers
    int feat_dim, ivec_dim, num_gauss, splice_dim;
    const float *lda_t;      // [splice_dim+1][feat_dim]  (transposed; last row = offset)
    const float *gconsts;    // [G]
    const float2 *ubm_t;     // [F][G] {mean * inv_var, -0.5 * inv_var}, transposed
    const float *sim;        // [G][F][D]   Sigma_i^{-1} M_i
    const float *U_tri;      // [G][D(D+1)/2]  M_i^T Sigma_i^{-1} M_i, packed lower triangle (row i: columns 0..i)
    const double *gcmvn_sum; // [F] global cmvn sums
    double gcmvn_count;
    float prior_offset;
    int num_gselect;
    float min_post, posterior_scale, max_count;
    int cmn_window, global_frames;
};
struct IvecState {           // per channel, device
    double *cmvn_sum;        // [C][F]
    float *norm_ring;        // [C][kNormRing][F]
    double *lin;             // [C][D]
    double *quad;            // [C][D(D+1)/2]  packed lower triangle
    double *num_frames;      // [C]
    float *ivec;             // [C][D]   current i-vector (prior offset removed)
};
constexpr int kNormRing = 128;

struct IvecArgs {
    const LaneDesc *lanes;
    int num_lanes;
    NodeDesc in_node;
    int context;
    IvecModel m;
    IvecState st;
    // per-step scratch between the frame kernel and the statistics kernel: [max_lanes][frames_cap] rows
    int *sel_g;              // [..][kMaxGselect] selected Gaussians of a frame
    float *sel_w;            // [..][kMaxGselect] their scaled posteriors (0 = pruned)
    float *fu;               // [..][F] LDA of the raw splice
    int frames_cap;          // vbk_ivector_frames_cap(samples_per_chunk)
};

// ---------------- K2: TDNN-F ----------------
struct NnetPlanArgs {
    const LaneDesc *lanes;
    int num_lanes;
    int num_nodes;
    const NodeDesc *nodes;   // device [num_nodes]
    int *node_end;           // [num_channels][kMaxNodes] next time to compute per node
    NodeLane *table;         // [num_nodes][max_lanes]
    int *rowoff;             // [num_nodes][max_lanes+1]
    int max_lanes;
    int2 *rows;              // [num_nodes][rows_cap] {channel, time} of every packed row (saves the GEMM tiles a search)
    int rows_cap;
};
struct GemmArgs {
    OpDesc op;
    NodeDesc in, out, byp;
    const LaneDesc *lanes;
    int num_lanes;
    const NodeLane *table;   // out node's table [max_lanes]
    const int *rowoff;       // out node's prefix [max_lanes+1]
    const int2 *rows;        // out node's packed rows {channel, time}
    const float *ivec;       // [C][ivec_dim]
    int ivec_dim;
    int max_rows;            // upper bound of total rows (grid sizing)
    int tc_mode;             // tensor-core operand split: 1 = fp16 hi/lo (kind::f16), 2 = TF32 hi/lo (kind::tf32)
    const void *map_hi;      // HOST pointers to the 128-byte TMA descriptors of W_hi / W_lo (tensor-core path)
    const void *map_lo;
};

// ---------------- K3: beam search ----------------
struct GraphDev {
    int num_states, num_arcs, start;
    const float *final_cost;
    const int *e_begin;      // [S+1]
    const int *eps_begin;    // [S]
    const int4 *arcs;        // {weight bits, nextstate, pdf, olabel | kNextHasEps if the next state has epsilon arcs}
    const int2 *state_arcs;  // [S+1] {first emitting arc, first epsilon arc}; epsilon arcs of s end at state_arcs[s+1].x
};
struct DecChannelState {     // one per channel
    int n_cur, parity, frame, error;
    int log_count, path_len, reached_final;
    float best_cost;
    int link_count, pad0, pad1, pad2;
    // of the current token list (what GetCutoff needs, gathered while the list was written): smallest state among the cheapest
    // tokens, #{cost < best + beam}, #{cost <= best + beam}, ordered bits of the minimum cost
    int nx_state, nx_lt, nx_le;
    unsigned nx_best;
};
// pruned raw lattice of one finished stream (lattice=1): written by lattice_prune_kernel, read by the host
struct LatHeader {
    int n_tok, n_links, n_final, start;  // start = lattice state of the initial token (-1: empty lattice)
    int error, frames, pad0, pad1;
};
constexpr int kPartialCap = 256;
constexpr int kNextHasEps = 0x40000000;  // flag in the olabel field of a device arc record (labels and state ids stay below 2^29)
constexpr int kArcSilence = 0x20000000;  // same field: the arc's transition-id belongs to an endpointing silence phone
constexpr int kArcFlagMask = kNextHasEps | kArcSilence;
constexpr int kEpsLinkFlag = 0x40000000;  // set in the destination field of an epsilon link
struct DecArgs {
    const LaneDesc *lanes;
    int num_lanes;
    GraphDev g;
    float beam, beam_delta, acoustic_scale;
    int max_active, min_active;
    int tok_cap, cand_cap, hash_size, log_cap, max_frames, path_cap;
    // loglikes
    NodeDesc out_node;
    const NodeLane *out_table;
    // per channel
    DecChannelState *cs;
    int *tok_state;          // [C][2][tok_cap]
    float *tok_cost;
    int *tok_arc;
    int *tok_prev;
    int *log_prev;           // [C][log_cap]
    int *log_arc;
    float *log_cost;
    int *log_state;          // optional (debug capture) or null
    int *log_frame_off;      // [C][max_frames+2]
    int *path;               // [C][path_cap]  best-path arcs, last arc first
    // per resident CTA scratch
    int *hash_key;           // [G][hash_size]
    unsigned long long *hash_val;
    int *hash_tok;
    int4 *cand;              // [G][cand_cap] {packed lo, packed hi, table slot, source}
    int *rank;               // [G][tok_cap]
    int *sv_pref, *sv_a0, *sv_src;    // [G][tok_cap] expandable survivors: arc prefix, first arc, log index
    float *sv_cost;
    int *win_owner;          // [G][cand_cap/32+2] survivor owning the first arc of each 32-arc window
    // lattice generation (lattice != 0): every arc below the frame's final cutoff whose cost is within lattice_beam
    // of the best cost of its destination state becomes a link {src token, dst token, arc, acoustic cost}
    int lattice;
    float lattice_beam;
    int4 *links;             // [C][link_cap] {src log index, dst log index | kEpsLinkFlag (-1 = dropped), arc, acoustic cost bits}
    int *link_off;           // [C][max_frames+3] segment s = links whose destination token belongs to frame s
    int link_cap;
    float *frame_offset;     // [C][max_frames+2] cost offset (-best token cost) of every frame: taken back out of the lattice arcs
    int *cand_next;          // [G][cand_cap] destination state of a candidate
    int *eps_work;           // [G][cand_cap] work list of the epsilon closure: candidates that became a state's best word and whose
                             // state has epsilon arcs
    unsigned *lat_extra;     // [C][log_cap] extra-cost scratch of the lattice pruning, or null: the pruning then
                             // reuses log_prev (whose content is dead once the best path has been traced)
    // per slot: outputs of the pruning for the lanes that finished in this step (index = lane)
    LatHeader *lat_hdr;      // [L]
    int4 *lat_links;         // [L][lat_link_cap] {src state, dst state, arc, acoustic cost bits}
    int2 *lat_final;         // [L][lat_final_cap] {state, final cost bits}
    int *lat_tok_frame;      // [L][lat_tok_cap] frame of each lattice state
    int *lat_tok_state;      // [L][lat_tok_cap] graph state (only when log_state is kept), else null
    int lat_link_cap, lat_final_cap, lat_tok_cap;
    // partial results (partials != 0): output labels of the current best path of every unfinished lane, newest first
    int *partial_words;      // [L][kPartialCap]
    int *partial_count;      // [L] words on the path (may exceed kPartialCap: the oldest are cut)
    int *endp_silence;       // [L] endpointing: trailing silence frames of the best path so far (TrailingSilenceLength) ...
    float *endp_relcost;     // [L] ... and FinalRelativeCost of the current frame (inf: no final token); null = off
    unsigned long long *counters;     // [32] profiling counters (tokens, arcs, ...)
    int *lane_load;          // [lanes] largest token count a lane saw in this launch (load feedback)
    int grid;                // CTAs' worth of scratch allocated per slot
    // one launch serves lanes [lane_begin, lane_end): CTAs pull the next lane from *queue (zeroed before the launch; lanes
    // are sorted by descending load, so this is longest-processing-time-first list scheduling) and use scratch
    // [scratch_base + blockIdx.x]; a step issues up to two launches (1024-thread CTAs for its heavy lanes, 256 for the rest)
    int lane_begin, lane_end, scratch_base;
    int *queue;
    int l1_slots;            // level-1 (shared memory) table entries of this launch (set by vbk_decode)
};

extern "C" {
int vbk_feat_smem_bytes(int samples_per_chunk);
cudaError_t vbk_mfcc(const FeatArgs *a, cudaStream_t s);
cudaError_t vbk_resample(const ResampleArgs *a, cudaStream_t s);
cudaError_t vbk_ivector(const IvecArgs *a, cudaStream_t s);  // three launches
int vbk_ivector_frames_cap(int samples_per_chunk);
cudaError_t vbk_nnet_plan(const NnetPlanArgs *a, cudaStream_t s);
cudaError_t vbk_gemm_fp32(const GemmArgs *a, cudaStream_t s);
cudaError_t vbk_gemm_tc(const GemmArgs *a, cudaStream_t s);
// tensor-core operand form of a weight matrix (tc_mode 1: fp16 hi + 2^11-scaled fp16 lo, [N][K rounded up to 8]; 2: TF32 hi / lo,
// fp32 [N][K]); hi and lo must hold N * K * 4 bytes each
cudaError_t vbk_split_weights(const float *w, int N, int K, int tc_mode, void *hi, void *lo, cudaStream_t s);
// encodes the TMA descriptor (box = one 128-byte swizzle row x the N tile) of such an operand into out128
cudaError_t vbk_make_weight_map(const void *w, int N, int K, int tc_mode, void *out128);
cudaError_t vbk_decode(const DecArgs *a, int heavy, cudaStream_t s);
// backward extra-cost pruning (lattice_beam) + compaction of the link log of the lanes whose stream ended in this step
cudaError_t vbk_lattice_prune(const DecArgs *a, cudaStream_t s);
// best path so far (no final costs) of every unfinished lane -> its output labels (GetBestPath(use_final_probs = false))
cudaError_t vbk_partial(const DecArgs *a, cudaStream_t s);
int vbk_decode_max_grid(int device);
int vbk_decode_blocks_per_sm(int threads);  // resident search CTAs per SM of a tier's CTA size
// copies rows [t_begin, t_begin+n) of a node ring for one channel into dst (debug capture / tests)
cudaError_t vbk_copy_rows(NodeDesc node, int channel, int t_begin, int n_rows, float *dst, cudaStream_t s);
cudaError_t vbk_split_tf32(const float *w, float *hi, float *lo, long long n, cudaStream_t s);
}

}  // namespace vb
