/* oracle.h — CPU restatement of the reference's recognition path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  Nothing under vosk-api_b200/ links, includes or calls it.
 *
 * PARITY UNPINNED: the reference keeps all arithmetic of this path in Kaldi/OpenFst
 * (alphacep/kaldi branch "vosk", floating, no pin — [REF travis/Dockerfile.manylinux:16,24]),
 * which is absent from /root/reference and cannot be built offline, and the reference holds no
 * golden vector / known-answer test for the path.  Each function below
 * restates the published Kaldi algorithm that the cited reference call site invokes; the
 * independent pins available offline are torchaudio.compliance.kaldi.mfcc (MFCC), fp64 numpy
 * (network, i-vector solve), a brute-force dense Viterbi (search) and the reference's own
 * src/json.h compiled from where it lies (result text; oracle/_ref).
 */
#ifndef VB_ORACLE_H
#define VB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- features: MFCC as configured by [REF training/conf/mfcc.conf:1-7] + --dither=0,
 *      called through OnlineNnet2FeaturePipeline [REF src/recognizer.cc:28,305-311]. ---- */
int orc_num_frames(int64_t num_samples);
int orc_mfcc(const int16_t *wave, int64_t n, float *out /* [T][40] */);

/* ---- online i-vector (OnlineIvectorFeature, options [REF src/model.cc:250-260]) ---- */
typedef struct {
    int feat_dim, ivec_dim, num_gauss;
    int splice_left, splice_right;
    int num_gselect;
    float min_post, posterior_scale, max_count;
    int cmn_window, global_frames;
    const float *lda;           /* [feat_dim][7*feat_dim+1] last column = offset */
    const float *gconsts;       /* [G] */
    const float *means_invvars; /* [G][F] */
    const float *inv_vars;      /* [G][F] */
    const float *M;             /* [G][F][D] */
    const float *sigma_inv;     /* [G][F][F] */
    float prior_offset;
    const double *global_cmvn;  /* [2][F+1] */
} OrcIvectorParams;
/* ends[c] = number of (spliced) frames whose statistics are included in i-vector c.
 * avail[c] = number of base frames available when i-vector c is computed (edge replication limit). */
int orc_ivectors(const OrcIvectorParams *p, const float *mfcc, int T, const int *ends, const int *avail,
                 int nchunks, float *out /* [nchunks][D] */);

/* ---- acoustic model: collapsed TDNN-F of [REF training/local/chain/run_tdnn.sh:98-129] ---- */
typedef struct {
    int feat_dim, ivec_dim, hidden, bottleneck;
    int num_tdnnf;
    const int *strides;
    int prefinal_small, prefinal_big, num_pdfs;
    float bypass_scale;
    const float *const *tensors; /* order documented in oracle/orc_nnet.cc */
} OrcNnetParams;
/* iv_index[t + ctx-2] selects the i-vector row used by tdnn1 frame t, t in [-(ctx-2), T-1+ctx-2]. */
int orc_nnet_forward(const OrcNnetParams *p, const float *mfcc, int T, const float *ivecs, const int *iv_index,
                     float *loglikes /* [(T+2)/3][num_pdfs] */);

/* ---- search: LatticeFasterDecoder token passing (config [REF src/batch_model.cc:78-80],
 *      [REF src/model.cc:135-137]) in the canonical order-independent form of DESIGN.md ---- */
typedef struct {
    int num_states, num_arcs, start;
    const float *final_cost;  /* [S] (+inf = not final) */
    const int *e_begin;       /* [S+1] first emitting arc of state s; e_begin[S] = num_arcs */
    const int *eps_begin;     /* [S]   first epsilon arc of state s (arcs [eps_begin[s], e_begin[s+1])) */
    const float *arc_w;
    const int *arc_next;
    const int *arc_pdf;
    const int *arc_src;
} OrcGraph;
typedef struct {
    float beam;
    int max_active, min_active;
    float beam_delta;
} OrcDecodeOpts;
typedef struct OrcDecoder OrcDecoder;
OrcDecoder *orc_decode(const OrcGraph *g, const OrcDecodeOpts *o, const float *loglikes, int num_frames, int num_pdfs);
void orc_decoder_free(OrcDecoder *d);
int orc_decoder_num_frames(const OrcDecoder *d);          /* frames actually decoded */
int64_t orc_decoder_num_tokens(const OrcDecoder *d);
/* offsets[f] .. offsets[f+1] = tokens logged for frame f (f = 0..num_frames), sorted by state id */
void orc_decoder_tokens(const OrcDecoder *d, int64_t *offsets, int *state, float *cost, int *arc, int64_t *prev);
int orc_decoder_best_path(const OrcDecoder *d, int *arcs, int cap, float *total_cost, int *reached_final);
/* raw lattice after FinalizeDecoding's lattice_beam pruning (GetRawLattice): lattice states = surviving tokens in log
 * order (tok_index maps them to token indices of orc_decoder_tokens), links {src, dst, csr arc, acoustic cost}, final
 * states with their final costs.  Returns the number of links (writes at most cap_links). */
int64_t orc_decoder_num_links(const OrcDecoder *d);
void orc_decoder_cost_offsets(const OrcDecoder *d, float *out /* [frames decoded] -best token cost of each frame */);
int64_t orc_decoder_lattice(const OrcDecoder *d, const OrcGraph *g, float lattice_beam, int64_t *n_states, int64_t *tok_index,
                            int64_t *lsrc, int64_t *ldst, int *larc, float *lac, int64_t cap_links,
                            int64_t *fin_state, float *fin_cost, int64_t *n_final);

/* ---- result text (PushLattice for a linear lattice [REF src/batch_recognizer.cc:43-107]) ---- */
typedef struct {
    const int *arc_ilabel, *arc_olabel;
    const int *tid2phone;
    const int *phone_type; /* [num_phones+1]: 0 none,1 nonword,2 begin,3 end,4 internal,5 singleton */
    int num_phones;
    const char *const *words;
    int num_words;
} OrcResultCtx;
/* words/begin/end (frames) of the best path; returns number of words */
int orc_align_words(const OrcResultCtx *c, const int *arcs, int n_arcs, int *word_ids, int *begin, int *end, int cap);
/* returns malloc'd string (caller frees with orc_free) */
char *orc_result_json(const OrcResultCtx *c, const int *arcs, int n_arcs, float offset_seconds, int nlsml);
void orc_free(void *p);

/* ---- lattice result chain (orc_lattice.cc): DeterminizeLatticePhonePrunedWrapper -> ScaleLattice(GraphLatticeScale(lm_scale)) ->
 *      WordAlignLattice -> MinimumBayesRisk -> text, as the pipeline + PushLattice do [REF src/batch_recognizer.cc:43-107,138-149].
 *      Input = the raw lattice of orc_decoder_lattice.  tid_flags (may be NULL = the generator's chain topology): bit 0 self-loop,
 *      bit 1 final transition, bit 2 leaves HMM state 0.  stage 0 = result text (JSON / NLSML), 1 = determinized + scaled
 *      CompactLattice, 2 = word-aligned CompactLattice as text lines ("S start", "A src dst word graph acoustic tids",
 *      "F state graph acoustic tids"), 5 = the raw MBR one-best ("word-id begin end confidence" per line, frames).  phone_pass 0 skips the phone-level first determinization pass.
 *      Returns a malloc'd string (orc_free). ---- */
char *orc_lattice_result(const OrcResultCtx *c, const float *arc_w, const unsigned char *tid_flags, int num_tids, int n_states, int start,
                         int64_t n_links, const int64_t *src, const int64_t *dst, const int *arc, const float *acoustic, int64_t n_final,
                         const int64_t *final_state, const float *final_cost, float lattice_beam, double lm_scale, float offset_seconds,
                         int nlsml, int stage, int phone_pass);

#ifdef __cplusplus
}
#endif
#endif
