// vb_result.h — result text for one finished segment: what BatchRecognizer::PushLattice produces
// [REF src/batch_recognizer.cc:43-107] for the best path (a linear lattice): word-aligned spans,
// MinimumBayesRisk one-best (conf = 1), json.h text layout [REF src/json.h:343-384] or NLSML.
#pragma once
#include <string>
#include <vector>

#include "vb_engine.h"
#include "vb_model.h"

namespace vb {

struct WordSpan {
    int word;
    float begin, end;  // decoder frames (30 ms each); fractional for MBR time averages
    float conf;
};

std::vector<WordSpan> align_words(const Model &m, const std::vector<int> &arcs);
std::string result_json(const Model &m, const std::vector<WordSpan> &words, float offset_seconds);
// same text from explicit word strings (host-only test hook of include/vosk_b200.h)
std::string result_json_words(const std::vector<std::string> &words, const std::vector<WordSpan> &spans, float offset_seconds);
std::string result_nlsml(const Model &m, const std::vector<WordSpan> &words);
std::string partial_json(const Model &m, const std::vector<WordSpan> &words);

// Kaldi LinearResample restated; the reference calls it with flush=true on every AcceptWaveform
// [REF src/batch_recognizer.cc:27-29,157-158], so every call is filtered independently.
class LinearResampler {
   public:
    LinearResampler(float rate_in, float rate_out, float cutoff, int num_zeros);
    void resample_flush(const std::vector<float> &in, std::vector<float> *out) const;
    bool identity() const { return identity_; }
    // number of output samples one flushed call of n_in input samples produces (LinearResample::GetNumOutputSamples)
    long long num_output(long long n_in) const;
    // input indices [*lo, *hi] (call coordinates, not yet clipped to the call) that outputs t0..t1 of one call read
    void input_range(long long t0, long long t1, long long *lo, long long *hi) const;
    int in_rate() const { return in_rate_; }
    int in_unit() const { return in_unit_; }
    int out_unit() const { return out_unit_; }
    const std::vector<int> &first_index() const { return first_index_; }
    const std::vector<std::vector<float>> &weights() const { return weights_; }

   private:
    int in_rate_, out_rate_, in_unit_, out_unit_;
    double cutoff_;
    int num_zeros_;
    bool identity_;
    std::vector<int> first_index_;
    std::vector<std::vector<float>> weights_;
};

}  // namespace vb
