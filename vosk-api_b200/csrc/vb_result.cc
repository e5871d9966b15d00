// vb_result.cc — see vb_result.h
#include "vb_result.h"

#include <cmath>
#include <numeric>
#include <sstream>

namespace vb {

std::vector<WordSpan> align_words(const Model &m, const std::vector<int> &arcs) {
    // WordAlignLattice on a linear lattice: a phone instance begins at every forward
    // transition-id and is extended by its self-loop ids; word = begin..end phones (or a singleton);
    // nonword (silence) phones yield no word; the k-th word span takes the k-th output label of the path.
    const Graph &g = m.graph;
    struct Seg { int phone, b, e; };
    std::vector<Seg> segs;
    std::vector<int> labels;
    int t = 0;
    for (int a : arcs) {
        if (g.arc_olabel[a] != 0) labels.push_back(g.arc_olabel[a]);
        const int tid = g.arc_ilabel[a];
        if (tid == 0) continue;
        // a phone instance begins at every transition-id that is not a self-loop (reorder = true: the self-loops follow)
        const bool forward = !(tid > 0 && tid < (int)m.tid_flags.size() && (m.tid_flags[tid] & 1));
        if (forward || segs.empty()) segs.push_back({m.tid2phone[tid], t, t + 1});
        else segs.back().e = t + 1;
        t++;
    }
    std::vector<WordSpan> out;
    size_t li = 0;
    int wb = -1;
    for (size_t s = 0; s < segs.size(); s++) {
        const int ph = segs[s].phone;
        const int ty = ph >= 0 && ph < (int)m.phone_type.size() ? m.phone_type[ph] : 0;
        bool emit = false;
        if (ty == 5) { wb = segs[s].b; emit = true; }
        else if (ty == 2) wb = segs[s].b;
        else if (ty == 3) { if (wb < 0) wb = segs[s].b; emit = true; }
        else if (ty == 4) { if (wb < 0) wb = segs[s].b; }
        const bool last = s + 1 == segs.size();
        if (!emit && last && wb >= 0 && li < labels.size()) emit = true;  // partial word forced out at the end
        if (emit) {
            if (li < labels.size()) out.push_back({labels[li++], (float)wb, (float)segs[s].e, 1.0f});
            wb = -1;
        }
    }
    return out;
}

static std::string escape(const std::string &s) {
    std::string o;
    for (char c : s) switch (c) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default: o += c;
        }
    return o;
}

static std::string word_str(const Model &m, int id) { return id >= 0 && id < (int)m.words.size() ? m.words[id] : std::string(); }

static std::string joined(const Model &m, const std::vector<WordSpan> &w) {
    std::string text;
    for (size_t i = 0; i < w.size(); i++) {
        if (i) text += ' ';
        text += word_str(m, w[i].word);
    }
    return text;
}

std::string result_json_words(const std::vector<std::string> &ws, const std::vector<WordSpan> &w, float offset) {
    // keys come out alphabetically (std::map in json.h), floats as "%f", arrays inline, 2-space pad per depth
    std::string s = "{\n", text;
    if (!w.empty()) {
        s += "  \"result\" : [";
        for (size_t i = 0; i < w.size(); i++) {
            // [REF src/batch_recognizer.cc:91-93]: round(frame) * 0.03 + offset, evaluated in double
            const double st = (double)std::round(w[i].begin) * 0.03 + (double)offset;
            const double en = (double)std::round(w[i].end) * 0.03 + (double)offset;
            if (i) s += ", ";
            s += "{\n      \"conf\" : " + std::to_string((double)w[i].conf);
            s += ",\n      \"end\" : " + std::to_string(en);
            s += ",\n      \"start\" : " + std::to_string(st);
            s += ",\n      \"word\" : \"" + escape(ws[i]) + "\"\n    }";
            if (i) text += ' ';
            text += ws[i];
        }
        s += "],\n";
    }
    s += "  \"text\" : \"" + escape(text) + "\"\n}";
    return s;
}

std::string result_json(const Model &m, const std::vector<WordSpan> &w, float offset) {
    std::vector<std::string> ws;
    for (auto &x : w) ws.push_back(word_str(m, x.word));
    return result_json_words(ws, w, offset);
}

std::string partial_json(const Model &m, const std::vector<WordSpan> &w) {
    // CPU API format [REF src/recognizer.cc:795-802]
    return "{\n  \"partial\" : \"" + escape(joined(m, w)) + "\"\n}";
}

std::string result_nlsml(const Model &m, const std::vector<WordSpan> &w) {
    // [REF src/batch_recognizer.cc:58-80]
    std::stringstream ss;
    float confidence = 0.f;
    for (auto &x : w) confidence += x.conf;
    confidence /= (float)w.size();
    const std::string text = joined(m, w);
    ss << "<?xml version=\"1.0\"?>\n<result grammar=\"default\">\n";
    ss << "<interpretation grammar=\"default\" confidence=\"" << confidence << "\">\n";
    ss << "<input mode=\"speech\">" << text << "</input>\n";
    ss << "<instance>" << text << "</instance>\n</interpretation>\n</result>\n";
    return ss.str();
}

// ------------------------------------------------------------------------------------------------
LinearResampler::LinearResampler(float rate_in, float rate_out, float cutoff, int num_zeros)
    : in_rate_((int)rate_in), out_rate_((int)rate_out), cutoff_(cutoff), num_zeros_(num_zeros) {
    identity_ = in_rate_ == out_rate_;  // taps are delta[0] up to ~1e-7: treated as identity
    const int base = std::gcd(in_rate_, out_rate_);
    in_unit_ = in_rate_ / base;
    out_unit_ = out_rate_ / base;
    const double window_width = num_zeros_ / (2.0 * cutoff_);
    first_index_.resize(out_unit_);
    weights_.resize(out_unit_);
    auto filter = [&](double t) {
        double window = std::fabs(t) < window_width ? 0.5 * (1 + std::cos(2 * M_PI * cutoff_ / num_zeros_ * t)) : 0.0;
        double f = t != 0 ? std::sin(2 * M_PI * cutoff_ * t) / (M_PI * t) : 2 * cutoff_;
        return f * window;
    };
    for (int i = 0; i < out_unit_; i++) {
        const double out_t = i / (double)out_rate_, min_t = out_t - window_width, max_t = out_t + window_width;
        const int lo = (int)std::ceil(min_t * in_rate_), hi = (int)std::floor(max_t * in_rate_);
        first_index_[i] = lo;
        weights_[i].resize(hi - lo + 1);
        for (int j = 0; j <= hi - lo; j++) {
            const double dt = (lo + j) / (double)in_rate_ - out_t;
            weights_[i][j] = (float)(filter(dt) / in_rate_);
        }
    }
}

long long LinearResampler::num_output(long long n_in) const {
    if (identity_) return n_in;
    const long long tick = std::lcm((long long)in_rate_, (long long)out_rate_);
    const long long ticks_in = tick / in_rate_, ticks_out = tick / out_rate_;
    const long long interval = n_in * ticks_in;
    if (interval <= 0) return 0;
    long long last = interval / ticks_out;
    if (last * ticks_out == interval) last--;
    return last + 1;
}

void LinearResampler::input_range(long long t0, long long t1, long long *lo, long long *hi) const {
    const long long u0 = t0 / out_unit_, u1 = t1 / out_unit_;
    const int w0 = (int)(t0 - u0 * out_unit_), w1 = (int)(t1 - u1 * out_unit_);
    *lo = first_index_[w0] + u0 * in_unit_;
    *hi = first_index_[w1] + u1 * in_unit_ + (long long)weights_[w1].size() - 1;
}

void LinearResampler::resample_flush(const std::vector<float> &in, std::vector<float> *out) const {
    if (identity_) {
        *out = in;
        return;
    }
    const long long n_out = num_output((long long)in.size());
    out->assign((size_t)n_out, 0.f);
    const long long n_in = (long long)in.size();
    for (long long so = 0; so < n_out; so++) {
        const long long unit = so / out_unit_;
        const int wrapped = (int)(so - unit * out_unit_);
        const long long first = first_index_[wrapped] + unit * in_unit_;
        const std::vector<float> &w = weights_[wrapped];
        float acc = 0.f;
        for (size_t i = 0; i < w.size(); i++) {
            const long long idx = first + (long long)i;
            if (idx >= 0 && idx < n_in) acc += w[i] * in[(size_t)idx];
        }
        (*out)[(size_t)so] = acc;
    }
}

}  // namespace vb
