// vb_lattice.h — host lattice pipeline of one finished segment (lattice=1):
//   raw lattice (device, lattice_beam-pruned)  ->  pruned word determinization  ->  graph scale 0.9  ->
//   word alignment  ->  MinimumBayesRisk one-best words / times / confidences.
// Replaces what the reference gets from Kaldi for RESULT_TYPE_LATTICE + BatchRecognizer::PushLattice
// [REF src/batch_recognizer.cc:43-56, 138-149]: DeterminizeLatticePhonePrunedWrapper (inside the cudadecoder
// pipeline's lattice post-processing), fst::ScaleLattice(GraphLatticeScale(0.9)), WordAlignLattice,
// MinimumBayesRisk::{GetOneBest, GetOneBestConfidences, GetOneBestTimes}.
#pragma once
#include <vector>

#include "vb_engine.h"
#include "vb_model.h"
#include "vb_result.h"

namespace vb {

struct LatWeight {  // Kaldi LatticeWeight: (graph cost, acoustic cost); natural order = total cost, then graph cost
    float g = 0.f, a = 0.f;
    float cost() const { return g + a; }
};

// CompactLattice: acceptor on words, weights carry the transition-id string
struct CLatArc {
    int dst;
    int word;  // 0 = epsilon / silence
    LatWeight w;
    std::vector<int> tids;
};
struct CLat {
    int start = -1;
    std::vector<std::vector<CLatArc>> arcs;
    std::vector<char> is_final;
    std::vector<LatWeight> final_w;
    std::vector<std::vector<int>> final_tids;
    int add_state() {
        arcs.emplace_back();
        is_final.push_back(0);
        final_w.emplace_back();
        final_tids.emplace_back();
        return (int)arcs.size() - 1;
    }
    size_t num_states() const { return arcs.size(); }
    size_t num_arcs() const {
        size_t n = 0;
        for (auto &a : arcs) n += a.size();
        return n;
    }
};

// what the model contributes to the lattice stages
struct LatticeCtx {
    const Graph *graph;                   // csr arc id -> ilabel (transition id), olabel (word), graph cost
    const std::vector<int32_t> *tid2phone;
    const std::vector<int> *phone_type;   // 0 none, 1 nonword, 2 begin, 3 end, 4 internal, 5 singleton
};

// DeterminizeLatticePruned on the word labels (see vb_lattice.cc for the stated differences to Kaldi's two-pass
// phone/word variant); beam = lattice_beam.  Returns false if the raw lattice has no complete path.
bool determinize_lattice(const RawLattice &raw, const LatticeCtx &ctx, float beam, CLat *out);
void scale_graph_costs(CLat *lat, float scale);
// WordAlignLattice (reorder = true, silence / partial-word label 0)
void word_align_lattice(const CLat &in, const LatticeCtx &ctx, CLat *out);
// MinimumBayesRisk (decode_mbr = true, print_silence = false): one-best words with begin/end (frames) and confidence
std::vector<WordSpan> mbr_one_best(const CLat &aligned);

// the whole chain, as PushLattice sees it
std::vector<WordSpan> lattice_to_words(const RawLattice &raw, const Model &m, float lattice_beam, float lm_scale = 0.9f);

}  // namespace vb
