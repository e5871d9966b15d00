// orc_decode.cc — oracle (TEST INFRASTRUCTURE, see oracle.h): WFST token passing, best path, result text.
//
// Search restates Kaldi's LatticeFasterDecoder (the core SingleUtteranceNnet3IncrementalDecoder runs
// for the reference's CPU recognizer [REF src/recognizer.cc:39-43,313-318]) with the decoder options of
// [REF src/batch_model.cc:78-80] / [REF src/model.cc:135-137], in the canonical ORDER-INDEPENDENT form
// defined in DESIGN.md:
//   * GetCutoff: best+beam, exact max_active-th / min_active-th order statistic, adaptive beam (+beam_delta)
//   * emitting step: tot = (tok + (cost_offset - loglike)) + arc.w ; next_cutoff = min(tot) + adaptive_beam
//     evaluated over ALL candidate arcs before filtering (Kaldi's running next_cutoff converges to the
//     same value; filtering with the final value makes the kept set independent of visiting order)
//   * one token per state = minimum of the 64-bit word (ordered-float(cost) << 32 | csr arc id): equal-cost
//     arrivals are resolved by the smaller canonical arc id, never by arrival order
//   * epsilon closure to the fixed point under the same rule, strictly below next_cutoff
//   * tokens that fail the next frame's cutoff are dropped from the log (they have no successors)
// Result text restates BatchRecognizer::PushLattice [REF src/batch_recognizer.cc:43-107] for a linear
// (best-path) lattice: WordAlignLattice spans, MinimumBayesRisk one-best (conf = 1), json.h layout
// [REF src/json.h:343-384].
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <sstream>
#include <string>
#include <vector>

namespace {
const float kInf = std::numeric_limits<float>::infinity();

inline uint32_t ord(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
inline float unord(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline uint64_t pack(float c, int arc) { return ((uint64_t)ord(c) << 32) | (uint32_t)arc; }
}  // namespace

struct OrcDecoder {
    std::vector<int64_t> offsets;  // per frame
    std::vector<int> state, arc;
    std::vector<float> cost;
    std::vector<int64_t> prev;
    int frames_decoded = 0;
    std::vector<int> best_arcs;
    float best_cost = kInf;
    int reached_final = 0;
    // ForwardLinks of LatticeFasterDecoder: every arc taken below the frame's cutoff, endpoints as log indices
    struct Link { int64_t src, dst; int arc; float ac; };
    std::vector<Link> links;
    std::vector<int64_t> link_off;  // segment f = links whose destination token is in frame f
    std::vector<float> cost_offset; // per decoded frame
};

namespace {
struct Frontier {
    // dense per-state best word + touched list
    std::vector<uint64_t> best;
    std::vector<int> touched;
    explicit Frontier(int S) : best(S, ~0ull) {}
    void clear() {
        for (int s : touched) best[s] = ~0ull;
        touched.clear();
    }
    bool relax(int s, uint64_t w) {
        if (w < best[s]) {
            if (best[s] == ~0ull) touched.push_back(s);
            best[s] = w;
            return true;
        }
        return false;
    }
};

void closure(const OrcGraph *g, Frontier &fr, float cutoff) {
    std::vector<int> queue(fr.touched);
    while (!queue.empty()) {
        int s = queue.back();
        queue.pop_back();
        float c = unord((uint32_t)(fr.best[s] >> 32));
        if (c >= cutoff) continue;
        for (int a = g->eps_begin[s]; a < g->e_begin[s + 1]; a++) {
            float tot = c + g->arc_w[a];
            if (tot < cutoff && fr.relax(g->arc_next[a], pack(tot, a))) queue.push_back(g->arc_next[a]);
        }
    }
}

float get_cutoff(const std::vector<float> &costs_in, const OrcDecodeOpts *o, float *adaptive_beam, float *best_out) {
    std::vector<float> tmp(costs_in);
    float best = kInf;
    for (float c : tmp) best = std::min(best, c);
    *best_out = best;
    float beam_cutoff = best + o->beam, min_active_cutoff = kInf, max_active_cutoff = kInf;
    if ((int)tmp.size() > o->max_active) {
        std::nth_element(tmp.begin(), tmp.begin() + o->max_active, tmp.end());
        max_active_cutoff = tmp[o->max_active];
    }
    if (max_active_cutoff < beam_cutoff) {
        *adaptive_beam = max_active_cutoff - best + o->beam_delta;
        return max_active_cutoff;
    }
    if ((int)tmp.size() > o->min_active) {
        if (o->min_active == 0)
            min_active_cutoff = best;
        else {
            std::nth_element(tmp.begin(), tmp.begin() + o->min_active,
                             (int)tmp.size() > o->max_active ? tmp.begin() + o->max_active : tmp.end());
            min_active_cutoff = tmp[o->min_active];
        }
    }
    if (min_active_cutoff > beam_cutoff) {
        *adaptive_beam = min_active_cutoff - best + o->beam_delta;
        return min_active_cutoff;
    }
    *adaptive_beam = o->beam;
    return beam_cutoff;
}
}  // namespace

extern "C" OrcDecoder *orc_decode(const OrcGraph *g, const OrcDecodeOpts *o, const float *loglikes, int N, int P) {
    OrcDecoder *d = new OrcDecoder;
    const int S = g->num_states;
    Frontier cur(S), nxt(S);
    std::vector<int64_t> log_index_prev(S, -1), log_index_cur(S, -1);
    std::vector<int> prev_states, cur_states;
    cur.relax(g->start, pack(0.f, -1));
    closure(g, cur, o->beam);
    d->offsets.push_back(0);
    struct Pending { int64_t src_log; int src_state; int dst_state; int arc; float ac; };
    std::vector<Pending> pending;  // links into the frontier that is logged next (src_state >= 0: epsilon link inside it)
    auto eps_links = [&](Frontier &fr, float cutoff) {
        // ProcessNonemitting: every token below the cutoff links to the target of each of its epsilon arcs below the cutoff
        for (int s : fr.touched) {
            float c = unord((uint32_t)(fr.best[s] >> 32));
            if (c >= cutoff) continue;
            for (int a = g->eps_begin[s]; a < g->e_begin[s + 1]; a++)
                if (c + g->arc_w[a] < cutoff) pending.push_back({-1, s, g->arc_next[a], a, 0.f});
        }
    };
    auto log_frame = [&](Frontier &fr, float cutoff, bool all) {
        // survivors sorted by state id; resolves prev via the previous/current frame's maps
        std::vector<int> sv;
        for (int s : fr.touched) {
            float c = unord((uint32_t)(fr.best[s] >> 32));
            if (all || c <= cutoff) sv.push_back(s);
        }
        std::sort(sv.begin(), sv.end());
        int64_t base = (int64_t)d->state.size();
        for (int s : prev_states) log_index_prev[s] = -1;
        prev_states.swap(cur_states);
        log_index_prev.swap(log_index_cur);
        cur_states = sv;
        for (size_t i = 0; i < sv.size(); i++) log_index_cur[sv[i]] = base + (int64_t)i;
        for (int s : sv) {
            uint64_t w = fr.best[s];
            int a = (int)(uint32_t)w;
            d->state.push_back(s);
            d->cost.push_back(unord((uint32_t)(w >> 32)));
            d->arc.push_back(a);
            int64_t pv = -1;
            if (a >= 0) {
                bool eps = g->arc_pdf[a] < 0;
                pv = eps ? log_index_cur[g->arc_src[a]] : log_index_prev[g->arc_src[a]];
                if (pv < 0) pv = -2;  // predecessor pruned (negative epsilon weight): not supported
            }
            d->prev.push_back(pv);
        }
        d->offsets.push_back((int64_t)d->state.size());
        // links into this frame: tokens that were not logged (they fail this frame's cutoff) take their links with them
        d->link_off.push_back((int64_t)d->links.size());
        for (const Pending &p : pending) {
            int64_t src = p.src_state >= 0 ? log_index_cur[p.src_state] : p.src_log;
            int64_t dst = log_index_cur[p.dst_state];
            if (src >= 0 && dst >= 0) d->links.push_back({src, dst, p.arc, p.ac});
        }
        pending.clear();
        return sv;
    };
    eps_links(cur, o->beam);
    int f = 0;
    for (; f < N; f++) {
        if (cur.touched.empty()) break;
        std::vector<float> costs;
        costs.reserve(cur.touched.size());
        for (int s : cur.touched) costs.push_back(unord((uint32_t)(cur.best[s] >> 32)));
        float adaptive_beam, best;
        float cur_cutoff = get_cutoff(costs, o, &adaptive_beam, &best);
        std::vector<int> sv = log_frame(cur, cur_cutoff, false);
        const float cost_offset = -best;
        d->cost_offset.push_back(cost_offset);
        const float *ll = loglikes + (size_t)f * P;
        float min_tot = kInf;
        for (int s : sv) {
            float c = unord((uint32_t)(cur.best[s] >> 32));
            for (int a = g->e_begin[s]; a < g->eps_begin[s]; a++) {
                float ac = cost_offset - ll[g->arc_pdf[a]];
                float tot = c + ac + g->arc_w[a];
                min_tot = std::min(min_tot, tot);
            }
        }
        float next_cutoff = min_tot + adaptive_beam;
        nxt.clear();
        for (int s : sv) {
            float c = unord((uint32_t)(cur.best[s] >> 32));
            for (int a = g->e_begin[s]; a < g->eps_begin[s]; a++) {
                float ac = cost_offset - ll[g->arc_pdf[a]];
                float tot = c + ac + g->arc_w[a];
                if (tot < next_cutoff) {
                    nxt.relax(g->arc_next[a], pack(tot, a));
                    pending.push_back({log_index_cur[s], -1, g->arc_next[a], a, ac});
                }
            }
        }
        closure(g, nxt, next_cutoff);
        eps_links(nxt, next_cutoff);
        std::swap(cur, nxt);
    }
    d->frames_decoded = f;
    log_frame(cur, kInf, true);
    d->link_off.push_back((int64_t)d->links.size());
    // best path: minimum of (cost + final, state); if no final state is active, minimum of (cost, state)
    int64_t lo = d->offsets[d->offsets.size() - 2], hi = d->offsets.back();
    int64_t bi = -1;
    uint64_t bw = ~0ull;
    for (int pass = 0; pass < 2 && bi < 0; pass++) {
        for (int64_t i = lo; i < hi; i++) {
            float fc = g->final_cost[d->state[i]];
            if (pass == 0 && fc == kInf) continue;
            float tot = pass == 0 ? d->cost[i] + fc : d->cost[i];
            uint64_t w = pack(tot, d->state[i]);
            if (w < bw) {
                bw = w;
                bi = i;
            }
        }
        if (bi >= 0) {
            d->reached_final = pass == 0;
            d->best_cost = unord((uint32_t)(bw >> 32));
        }
    }
    for (int64_t i = bi; i >= 0 && d->arc[i] >= 0; i = d->prev[i]) d->best_arcs.push_back(d->arc[i]);
    std::reverse(d->best_arcs.begin(), d->best_arcs.end());
    return d;
}

extern "C" void orc_decoder_free(OrcDecoder *d) { delete d; }
extern "C" int orc_decoder_num_frames(const OrcDecoder *d) { return d->frames_decoded; }
extern "C" int64_t orc_decoder_num_tokens(const OrcDecoder *d) { return (int64_t)d->state.size(); }
extern "C" void orc_decoder_tokens(const OrcDecoder *d, int64_t *offsets, int *state, float *cost, int *arc, int64_t *prev) {
    memcpy(offsets, d->offsets.data(), d->offsets.size() * sizeof(int64_t));
    size_t n = d->state.size();
    memcpy(state, d->state.data(), n * sizeof(int));
    memcpy(cost, d->cost.data(), n * sizeof(float));
    memcpy(arc, d->arc.data(), n * sizeof(int));
    memcpy(prev, d->prev.data(), n * sizeof(int64_t));
}
extern "C" int orc_decoder_best_path(const OrcDecoder *d, int *arcs, int cap, float *total_cost, int *reached_final) {
    int n = (int)d->best_arcs.size();
    for (int i = 0; i < n && i < cap; i++) arcs[i] = d->best_arcs[i];
    if (total_cost) *total_cost = d->best_cost;
    if (reached_final) *reached_final = d->reached_final;
    return n;
}

// --------------------------------------------------------------------------------------------
// raw lattice: LatticeFasterDecoder::FinalizeDecoding (PruneForwardLinksFinal, then PruneForwardLinks(delta = 0) and
// PruneTokensForFrame for every earlier frame) + GetRawLattice, on the link log above.  Token extra costs start at 0
// (Token constructor) and every frame is swept in list order until nothing changes, as Kaldi does; the final frame is
// iterated to the exact fixed point (Kaldi stops it at a relative change of 1e-5).
// --------------------------------------------------------------------------------------------
extern "C" int64_t orc_decoder_lattice(const OrcDecoder *d, const OrcGraph *g, float lattice_beam, int64_t *n_states_out,
                                       int64_t *tok_index /* [n_tokens] lattice state -> token index */,
                                       int64_t *lsrc, int64_t *ldst, int *larc, float *lac, int64_t cap_links,
                                       int64_t *fin_state, float *fin_cost, int64_t *n_final_out) {
    const int64_t nt = (int64_t)d->state.size();
    const int F = d->frames_decoded;
    std::vector<float> extra(nt, 0.f);
    std::vector<char> alive(d->links.size(), 1);
    // links grouped by source token
    std::vector<std::vector<int64_t>> out(nt);
    for (size_t k = 0; k < d->links.size(); k++) out[d->links[k].src].push_back((int64_t)k);
    auto link_extra = [&](const OrcDecoder::Link &l) {
        float graph = g->arc_w[l.arc];
        return extra[l.dst] + ((d->cost[l.src] + l.ac + graph) - d->cost[l.dst]);
    };
    for (int f = F; f >= 0; f--) {
        bool changed = true;
        while (changed) {
            changed = false;
            for (int64_t i = d->offsets[f]; i < d->offsets[f + 1]; i++) {
                float tok_extra = kInf;
                if (f == F) {
                    float fc = d->reached_final ? g->final_cost[d->state[i]] : 0.f;
                    tok_extra = d->cost[i] + fc - d->best_cost;
                }
                for (int64_t k : out[i]) {
                    if (!alive[k]) continue;
                    float le = link_extra(d->links[k]);
                    if (le > lattice_beam) {
                        alive[k] = 0;
                    } else {
                        if (le < 0.f) le = 0.f;
                        if (le < tok_extra) tok_extra = le;
                    }
                }
                if (f == F && tok_extra > lattice_beam) tok_extra = kInf;
                if (tok_extra != extra[i]) changed = true;
                extra[i] = tok_extra;
            }
        }
    }
    std::vector<int64_t> newidx(nt, -1);
    int64_t ns = 0;
    for (int64_t i = 0; i < nt; i++)
        if (extra[i] != kInf) {
            tok_index[ns] = i;
            newidx[i] = ns++;
        }
    *n_states_out = ns;
    int64_t nl = 0;
    for (size_t k = 0; k < d->links.size(); k++) {
        const OrcDecoder::Link &l = d->links[k];
        if (!alive[k] || newidx[l.src] < 0 || newidx[l.dst] < 0) continue;
        if (nl < cap_links) {
            lsrc[nl] = newidx[l.src];
            ldst[nl] = newidx[l.dst];
            larc[nl] = l.arc;
            // GetRawLattice: Weight(graph_cost, acoustic_cost - cost_offsets_[frame]) for emitting arcs
            bool eps = g->arc_pdf[l.arc] < 0;
            int64_t fsrc = std::upper_bound(d->offsets.begin(), d->offsets.end(), l.src) - d->offsets.begin() - 1;
            lac[nl] = eps ? 0.f : l.ac - d->cost_offset[fsrc];
        }
        nl++;
    }
    int64_t nf = 0;
    for (int64_t i = d->offsets[F]; i < d->offsets[F + 1]; i++) {
        if (newidx[i] < 0) continue;
        float fc = d->reached_final ? g->final_cost[d->state[i]] : 0.f;
        if (fc == kInf) continue;
        fin_state[nf] = newidx[i];
        fin_cost[nf] = fc;
        nf++;
    }
    *n_final_out = nf;
    return nl;
}
extern "C" int64_t orc_decoder_num_links(const OrcDecoder *d) { return (int64_t)d->links.size(); }
extern "C" void orc_decoder_cost_offsets(const OrcDecoder *d, float *out) { memcpy(out, d->cost_offset.data(), d->cost_offset.size() * sizeof(float)); }

// --------------------------------------------------------------------------------------------
// word alignment of a linear path and result text
// --------------------------------------------------------------------------------------------
extern "C" int orc_align_words(const OrcResultCtx *c, const int *arcs, int n, int *word_ids, int *begin, int *end, int cap) {
    // frame t = index among emitting arcs.  A phone instance starts at every forward transition id
    // (even tid in the chain topology of the model generator: tid = 2*tstate+2); self-loop tids extend it.
    std::vector<int> labels;  // olabels in path order
    struct Seg { int phone, b, e; };
    std::vector<Seg> segs;
    int t = 0;
    for (int i = 0; i < n; i++) {
        int a = arcs[i];
        if (c->arc_olabel[a] != 0) labels.push_back(c->arc_olabel[a]);
        int tid = c->arc_ilabel[a];
        if (tid == 0) continue;
        bool fwd = (tid % 2) == 0;
        if (fwd || segs.empty()) segs.push_back({c->tid2phone[tid], t, t + 1});
        else segs.back().e = t + 1;
        t++;
    }
    int nw = 0;
    size_t li = 0;
    int wb = -1;
    for (size_t s = 0; s < segs.size(); s++) {
        int ty = (segs[s].phone >= 0 && segs[s].phone <= c->num_phones) ? c->phone_type[segs[s].phone] : 0;
        bool emit = false;
        if (ty == 5) { wb = segs[s].b; emit = true; }
        else if (ty == 2) wb = segs[s].b;
        else if (ty == 3) { if (wb < 0) wb = segs[s].b; emit = true; }
        else if (ty == 4) { if (wb < 0) wb = segs[s].b; }
        bool last = s + 1 == segs.size();
        if (!emit && last && wb >= 0 && li < labels.size()) emit = true;  // partial word forced out at the end
        if (emit) {
            if (li < labels.size() && nw < cap) {
                word_ids[nw] = labels[li++];
                begin[nw] = wb;
                end[nw] = segs[s].e;
                nw++;
            }
            wb = -1;
        }
    }
    return nw;
}

namespace {
std::string json_escape(const std::string &s) {
    std::string o;
    for (char ch : s) switch (ch) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default: o += ch;
        }
    return o;
}
}  // namespace

extern "C" char *orc_result_json(const OrcResultCtx *c, const int *arcs, int n, float offset, int nlsml) {
    std::vector<int> w(n + 1), b(n + 1), e(n + 1);
    int nw = orc_align_words(c, arcs, n, w.data(), b.data(), e.data(), n + 1);
    auto word = [&](int id) { return std::string(id >= 0 && id < c->num_words ? c->words[id] : ""); };
    std::string text;
    for (int i = 0; i < nw; i++) text += (i ? " " : "") + word(w[i]);
    std::string out;
    if (nlsml) {
        std::stringstream ss;
        float conf = 0.f;
        for (int i = 0; i < nw; i++) conf += 1.0f;
        conf /= nw;
        ss << "<?xml version=\"1.0\"?>\n<result grammar=\"default\">\n";
        char cbuf[64];
        snprintf(cbuf, sizeof cbuf, "%g", (double)conf);  // = ostream << float (see orc_lattice.cc)
        ss << "<interpretation grammar=\"default\" confidence=\"" << cbuf << "\">\n";
        ss << "<input mode=\"speech\">" << text << "</input>\n";
        ss << "<instance>" << text << "</instance>\n</interpretation>\n</result>\n";
        out = ss.str();
    } else {
        out = "{\n";
        if (nw > 0) {
            out += "  \"result\" : [";
            for (int i = 0; i < nw; i++) {
                double st = (double)std::round((float)b[i]) * 0.03 + (double)offset;
                double en = (double)std::round((float)e[i]) * 0.03 + (double)offset;
                if (i) out += ", ";
                out += "{\n      \"conf\" : " + std::to_string((double)1.0f) + ",\n      \"end\" : " + std::to_string(en) +
                       ",\n      \"start\" : " + std::to_string(st) + ",\n      \"word\" : \"" + json_escape(word(w[i])) + "\"\n    }";
            }
            out += "],\n";
        }
        out += "  \"text\" : \"" + json_escape(text) + "\"\n}";
    }
    char *r = (char *)malloc(out.size() + 1);
    memcpy(r, out.c_str(), out.size() + 1);
    return r;
}

extern "C" void orc_free(void *p) { free(p); }
