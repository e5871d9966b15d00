// vb_lattice.h — host lattice pipeline of one finished segment (the reference's result path):
//   raw lattice (device, lattice_beam-pruned)  ->  phone-pruned determinization (phone pass, then word pass)  ->
//   graph scale 0.9  ->  word alignment  ->  MinimumBayesRisk one-best words / times / confidences.
// Replaces what the reference gets from Kaldi for RESULT_TYPE_LATTICE + BatchRecognizer::PushLattice
// [REF src/batch_recognizer.cc:43-56, 138-149]: DeterminizeLatticePhonePrunedWrapper (inside the cudadecoder
// pipeline's lattice post-processing), fst::ScaleLattice(GraphLatticeScale(0.9)), WordAlignLattice,
// MinimumBayesRisk::{GetOneBest, GetOneBestConfidences, GetOneBestTimes}.
// Flat-array implementation with per-thread scratch that is reused from segment to segment (no allocation in steady
// state); the CPU oracle (oracle/orc_lattice.cc) restates the same algorithms in Kaldi's own object structure.
#pragma once
#include <string>
#include <vector>

#include "vb_engine.h"
#include "vb_mbr.h"
#include "vb_model.h"
#include "vb_result.h"

namespace vb {

struct LatticeStats {
    int raw_states = 0, raw_arcs = 0;
    int det1_states = 0, det1_arcs = 0;  // phone pass (expanded form)
    int det_states = 0, det_arcs = 0;    // word pass (compact form, trimmed)
    int ali_states = 0, ali_arcs = 0;    // word-aligned lattice
    int mbr_iters = 0, mbr_q = 0;
    double ms_det = 0, ms_align = 0, ms_mbr = 0;
};

// One word-aligned lattice as MinimumBayesRisk sees it (flat arrays; PrepareLatticeAndInitStats + the arc posteriors): the input of
// mbr_solve (vb_mbr.h), on the host or on the device.  N = 0: nothing to decode (no complete path).
struct MbrJob {
    int N = 0;
    std::vector<MbrArc> arcs;        // grouped by end node, in start-node order
    std::vector<int> pre_off;        // [N + 2]
    std::vector<int> state_times;    // [N + 1]
    std::vector<double> post;        // [arcs]
    std::vector<int> R0;             // words of the best path
};
// determinization + graph scale + word alignment + MBR preparation (false: the lattice could not be determinized)
bool lattice_to_mbr_job(const RawLattice &raw, const Model &m, float lattice_beam, double lm_scale, MbrJob *job, LatticeStats *stats = nullptr);
// the MBR decision loop on the host
std::vector<WordSpan> mbr_solve_host(const MbrJob &job, LatticeStats *stats = nullptr);

// the whole chain, as the pipeline + PushLattice run it.  Returns the MBR one-best; *ok = false when the lattice could not be
// determinized (no complete path, not a lattice), true when the chain ran — an empty result then is a lattice without words.
std::vector<WordSpan> lattice_to_words(const RawLattice &raw, const Model &m, float lattice_beam, double lm_scale = 0.9,
                                       LatticeStats *stats = nullptr, bool *ok = nullptr);

// test hook: stage 1 = determinized (and graph-scaled) lattice, 2 = word-aligned lattice, as text lines
// "S start" / "A src dst word graph acoustic tids" / "F state graph acoustic tids"
std::string lattice_debug_text(const RawLattice &raw, const Model &m, float lattice_beam, int stage, bool phone_pass = true,
                               double lm_scale = 0.9);

}  // namespace vb
