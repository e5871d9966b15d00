// vb_lattice.cc — see vb_lattice.h.
//
// Everything here restates published Kaldi algorithms whose sources are NOT in /root/reference (they live in
// alphacep/kaldi); the reference only names them at its call sites
// [REF src/batch_recognizer.cc:45-54].  Stated differences:
//   * determinization runs once on the word labels (Kaldi's wrapper first determinizes with phone boundaries
//     inserted, then on words; both yield a word-deterministic lattice holding the best alignment per word sequence);
//     tasks are expanded best-first on (forward cost + arc + exact backward cost), arcs beyond best + beam are dropped;
//   * subsets are matched on exact states and strings and on weights within 1/1024 (Kaldi's kDelta);
//   * there is no max_mem retry loop: if a lattice grows beyond kMaxDetStates the beam is halved and the
//     determinization redone.
#include "vb_lattice.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <queue>
#include <unordered_map>

namespace vb {

namespace {
constexpr float kInfF = std::numeric_limits<float>::infinity();
constexpr double kInfD = std::numeric_limits<double>::infinity();
constexpr int kMaxDetStates = 200000;
constexpr float kDelta = 1.0f / 1024.0f;

// fst::Compare(LatticeWeight): 1 if a is better (lower total cost, then lower graph cost)
inline int compare_w(const LatWeight &a, const LatWeight &b) {
    const float fa = a.g + a.a, fb = b.g + b.a;
    if (fa < fb) return 1;
    if (fa > fb) return -1;
    if (a.g < b.g) return 1;
    if (a.g > b.g) return -1;
    return 0;
}
inline LatWeight times_w(const LatWeight &a, const LatWeight &b) { return LatWeight{a.g + b.g, a.a + b.a}; }
inline LatWeight divide_w(const LatWeight &a, const LatWeight &b) { return LatWeight{a.g - b.g, a.a - b.a}; }
// LatticeDeterminizerPruned::Compare on strings: the shorter one is better, then the lexicographically larger (sic)
inline int compare_str(const std::vector<int> &a, const std::vector<int> &b) {
    if (a.size() > b.size()) return -1;
    if (a.size() < b.size()) return 1;
    for (size_t i = 0; i < a.size(); i++) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}

struct RArc {
    int dst, word, tid;
    LatWeight w;
};
struct RLat {
    int start = -1;
    std::vector<std::vector<RArc>> arcs;
    std::vector<float> final_cost;  // inf = not final
};

struct Elem {
    int state;
    LatWeight w;
    std::vector<int> str;
};
inline int compare_elem(const LatWeight &aw, const std::vector<int> &as, const LatWeight &bw, const std::vector<int> &bs) {
    const int c = compare_w(aw, bw);
    return c ? c : compare_str(as, bs);
}

RLat build_raw(const RawLattice &raw, const LatticeCtx &ctx) {
    RLat r;
    r.start = raw.start;
    r.arcs.resize(raw.n_states);
    r.final_cost.assign(raw.n_states, kInfF);
    const Graph &g = *ctx.graph;
    for (size_t k = 0; k < raw.src.size(); k++) {
        const int a = raw.arc[k];
        if (raw.src[k] < 0 || raw.dst[k] < 0 || a < 0 || a >= g.num_arcs) continue;
        r.arcs[raw.src[k]].push_back(RArc{raw.dst[k], g.arc_olabel[a], g.arc_ilabel[a], LatWeight{g.arc_w[a], raw.acoustic[k]}});
    }
    for (size_t k = 0; k < raw.final_state.size(); k++)
        if (raw.final_state[k] >= 0 && raw.final_state[k] < raw.n_states) r.final_cost[raw.final_state[k]] = raw.final_cost[k];
    return r;
}

// states reachable from the start, in topological order (the lattice is acyclic)
template <class ArcsOf>
std::vector<int> topo_order(int n, int start, ArcsOf arcs_of) {
    std::vector<int> indeg(n, 0), order;
    std::vector<char> seen(n, 0);
    std::vector<int> stack{start};
    seen[start] = 1;
    while (!stack.empty()) {
        int s = stack.back();
        stack.pop_back();
        arcs_of(s, [&](int d) {
            indeg[d]++;
            if (!seen[d]) {
                seen[d] = 1;
                stack.push_back(d);
            }
        });
    }
    stack.push_back(start);
    while (!stack.empty()) {
        int s = stack.back();
        stack.pop_back();
        order.push_back(s);
        arcs_of(s, [&](int d) {
            if (--indeg[d] == 0) stack.push_back(d);
        });
    }
    return order;
}

class Determinizer {
   public:
    Determinizer(const RLat &in, float beam) : in_(in), beam_(beam) {}

    bool run(CLat *out) {
        out_ = out;
        *out = CLat();
        if (in_.start < 0) return false;
        backward_costs();
        if (beta_[in_.start] == kInfD) return false;
        cutoff_ = beta_[in_.start] + beam_;
        std::vector<Elem> sub{Elem{in_.start, LatWeight{}, {}}};
        closure(&sub);
        to_minimal(&sub);
        LatWeight w0;
        std::vector<int> s0;
        normalize(&sub, &w0, &s0);
        const int first = new_state(sub, (double)w0.cost());
        if (w0.g != 0.f || w0.a != 0.f || !s0.empty()) {
            const int st = out_->add_state();  // extra start state carrying the initial weight / string
            states_.push_back(OutState{{}, 0.0});  // keeps states_ aligned with the output state ids
            out_->arcs[st].push_back(CLatArc{first, 0, w0, s0});
            out_->start = st;
        } else {
            out_->start = first;
        }
        process_state(first);
        while (!tasks_.empty()) {
            if ((int)out_->num_states() > kMaxDetStates) return false;
            Task t = tasks_.top();
            tasks_.pop();
            bool is_new = false;
            const int dst = find_or_add(*t.subset, states_[t.src].fwd + (double)t.w.cost(), &is_new);
            out_->arcs[t.src].push_back(CLatArc{dst, t.label, t.w, t.str});
            if (is_new) process_state(dst);
        }
        return true;
    }

   private:
    struct OutState {
        std::vector<Elem> subset;
        double fwd;
    };
    struct Task {
        double prio;
        int src, label;
        std::shared_ptr<std::vector<Elem>> subset;
        LatWeight w;
        std::vector<int> str;
        bool operator<(const Task &o) const { return prio > o.prio; }  // min-heap
    };

    void backward_costs() {
        const int n = (int)in_.arcs.size();
        beta_.assign(n, kInfD);
        auto arcs_of = [&](int s, auto f) {
            for (const RArc &a : in_.arcs[s]) f(a.dst);
        };
        std::vector<int> order = topo_order(n, in_.start, arcs_of);
        topo_.assign(n, n);
        for (size_t i = 0; i < order.size(); i++) topo_[order[i]] = (int)i;
        stamp_.assign(n, 0);
        slot_.assign(n, 0);
        keepable_.assign(n, 0);
        for (int s2 = 0; s2 < n; s2++) {
            bool k = in_.final_cost[s2] != kInfF;
            for (const RArc &a : in_.arcs[s2]) k = k || a.word != 0;
            keepable_[s2] = k;
        }
        for (auto it = order.rbegin(); it != order.rend(); ++it) {
            const int s = *it;
            double b = in_.final_cost[s] == kInfF ? kInfD : (double)in_.final_cost[s];
            for (const RArc &a : in_.arcs[s])
                if (beta_[a.dst] != kInfD) b = std::min(b, (double)a.w.cost() + beta_[a.dst]);
            beta_[s] = b;
        }
    }

    // Follow arcs without a word label; per state keep the best (weight, then string) element; then drop the elements
    // whose state neither is final nor has an arc with a word label (ConvertToMinimal).  The raw lattice is acyclic, so
    // states are expanded in topological order (each once, after all its predecessors inside the closure); strings grow
    // as parent-linked nodes in an arena and are materialised only for the elements that survive.
    struct SNode { int parent, tid; };
    struct Work { int state; LatWeight w; int base, node; };  // string = input element `base`'s string + arena path `node`
    std::vector<int> materialize(const std::vector<Elem> &in, const Work &x, const std::vector<SNode> &arena) const {
        std::vector<int> tail;
        for (int n = x.node; n >= 0; n = arena[n].parent) tail.push_back(arena[n].tid);
        std::vector<int> out(in[x.base].str);
        out.insert(out.end(), tail.rbegin(), tail.rend());
        return out;
    }
    void closure_minimal(std::vector<Elem> *sub) {
        const std::vector<Elem> &in = *sub;
        std::vector<Work> cur;
        std::vector<SNode> arena;
        epoch_++;
        typedef std::pair<int, int> QE;  // (topological index, work index)
        std::priority_queue<QE, std::vector<QE>, std::greater<QE>> heap;
        for (size_t i = 0; i < in.size(); i++) {
            cur.push_back(Work{in[i].state, in[i].w, (int)i, -1});
            stamp_[in[i].state] = epoch_;
            slot_[in[i].state] = (int)i;
            heap.push(QE(topo_[in[i].state], (int)i));
        }
        std::vector<char> done;
        while (!heap.empty()) {
            const int i = heap.top().second;
            heap.pop();
            if ((int)done.size() < (int)cur.size()) done.resize(cur.size(), 0);
            if (done[i]) continue;
            done[i] = 1;
            const Work e = cur[i];
            for (const RArc &a : in_.arcs[e.state]) {
                if (a.word != 0) continue;
                int node = e.node;
                if (a.tid != 0) {
                    arena.push_back(SNode{e.node, a.tid});
                    node = (int)arena.size() - 1;
                }
                Work nx{a.dst, times_w(e.w, a.w), e.base, node};
                if (stamp_[a.dst] != epoch_) {
                    stamp_[a.dst] = epoch_;
                    slot_[a.dst] = (int)cur.size();
                    cur.push_back(nx);
                    heap.push(QE(topo_[a.dst], (int)cur.size() - 1));
                } else {
                    Work &old = cur[slot_[a.dst]];
                    int c = compare_w(nx.w, old.w);
                    if (c == 0) c = compare_str(materialize(in, nx, arena), materialize(in, old, arena));
                    if (c == 1) old = nx;  // (its successors have not been expanded yet: topological order)
                }
            }
        }
        std::vector<Elem> keep;
        for (const Work &x : cur)
            if (keepable_[x.state]) keep.push_back(Elem{x.state, x.w, materialize(in, x, arena)});
        std::sort(keep.begin(), keep.end(), [](const Elem &x, const Elem &y) { return x.state < y.state; });
        sub->swap(keep);
    }
    void closure(std::vector<Elem> *sub) { closure_minimal(sub); }
    void to_minimal(std::vector<Elem> *) {}  // (folded into closure_minimal)

    // take the best weight and the longest common string prefix out of the subset
    void normalize(std::vector<Elem> *sub, LatWeight *tot, std::vector<int> *common) {
        *tot = LatWeight{};
        common->clear();
        if (sub->empty()) return;
        LatWeight best = (*sub)[0].w;
        size_t pre = (*sub)[0].str.size();
        for (const Elem &e : *sub) {
            if (compare_w(e.w, best) == 1) best = e.w;
            size_t k = 0;
            while (k < pre && k < e.str.size() && e.str[k] == (*sub)[0].str[k]) k++;
            pre = k;
        }
        common->assign((*sub)[0].str.begin(), (*sub)[0].str.begin() + pre);
        for (Elem &e : *sub) {
            e.w = divide_w(e.w, best);
            e.str.erase(e.str.begin(), e.str.begin() + pre);
        }
        *tot = best;
    }

    static std::string key_of(const std::vector<Elem> &sub) {
        std::string k;
        for (const Elem &e : sub) {
            k.append(reinterpret_cast<const char *>(&e.state), 4);
            const int n = (int)e.str.size();
            k.append(reinterpret_cast<const char *>(&n), 4);
            k.append(reinterpret_cast<const char *>(e.str.data()), (size_t)n * 4);
        }
        return k;
    }
    int new_state(const std::vector<Elem> &sub, double fwd) {
        const int id = out_->add_state();
        states_.push_back(OutState{sub, fwd});
        index_[key_of(sub)].push_back(id);
        return id;
    }
    int find_or_add(const std::vector<Elem> &sub, double fwd, bool *is_new) {
        auto it = index_.find(key_of(sub));
        if (it != index_.end())
            for (int id : it->second) {
                const std::vector<Elem> &o = states_[id].subset;
                bool same = true;
                for (size_t i = 0; i < sub.size() && same; i++)
                    same = std::fabs(sub[i].w.g - o[i].w.g) <= kDelta && std::fabs(sub[i].w.a - o[i].w.a) <= kDelta;
                if (same) {
                    *is_new = false;
                    return id;
                }
            }
        *is_new = true;
        return new_state(sub, fwd);
    }

    void process_state(int s) {
        const std::vector<Elem> subset = states_[s].subset;  // copy: states_ may grow
        const double fwd = states_[s].fwd;
        // final weight: best of element weight x final weight of its state
        bool have = false;
        LatWeight fw;
        std::vector<int> fs;
        for (const Elem &e : subset) {
            if (in_.final_cost[e.state] == kInfF) continue;
            const LatWeight w = times_w(e.w, LatWeight{in_.final_cost[e.state], 0.f});
            if (!have || compare_elem(w, e.str, fw, fs) == 1) {
                fw = w;
                fs = e.str;
                have = true;
            }
        }
        if (have) {
            out_->is_final[s] = 1;
            out_->final_w[s] = fw;
            out_->final_tids[s] = fs;
        }
        std::map<int, std::vector<Elem>> by_label;
        for (const Elem &e : subset)
            for (const RArc &a : in_.arcs[e.state]) {
                if (a.word == 0) continue;
                Elem nx{a.dst, times_w(e.w, a.w), e.str};
                if (a.tid != 0) nx.str.push_back(a.tid);
                by_label[a.word].push_back(std::move(nx));
            }
        for (auto &kv : by_label) {
            std::vector<Elem> &v = kv.second;
            std::stable_sort(v.begin(), v.end(), [](const Elem &x, const Elem &y) { return x.state < y.state; });
            std::vector<Elem> merged;
            for (Elem &e : v) {
                if (!merged.empty() && merged.back().state == e.state) {
                    if (compare_elem(e.w, e.str, merged.back().w, merged.back().str) == 1) merged.back() = std::move(e);
                } else {
                    merged.push_back(std::move(e));
                }
            }
            closure(&merged);
            to_minimal(&merged);
            if (merged.empty()) continue;
            Task t;
            normalize(&merged, &t.w, &t.str);
            double best_back = kInfD;
            for (const Elem &e : merged) best_back = std::min(best_back, (double)e.w.cost() + beta_[e.state]);
            t.prio = fwd + (double)t.w.cost() + best_back;
            if (!(t.prio <= cutoff_ + 1e-4)) continue;  // beyond the beam (or no way to a final state)
            t.src = s;
            t.label = kv.first;
            t.subset = std::make_shared<std::vector<Elem>>(std::move(merged));
            tasks_.push(std::move(t));
        }
    }

    const RLat &in_;
    float beam_;
    CLat *out_ = nullptr;
    std::vector<double> beta_;
    std::vector<int> topo_, stamp_, slot_;
    std::vector<char> keepable_;
    int epoch_ = 0;
    double cutoff_ = 0;
    std::vector<OutState> states_;  // index = output state id of the subset states (the optional extra start state is added last)
    std::unordered_map<std::string, std::vector<int>> index_;
    std::priority_queue<Task> tasks_;
};
}  // namespace

bool determinize_lattice(const RawLattice &raw, const LatticeCtx &ctx, float beam, CLat *out) {
    if (raw.n_states <= 0 || raw.start < 0) return false;
    const RLat r = build_raw(raw, ctx);
    for (int attempt = 0; attempt < 4; attempt++, beam *= 0.5f) {
        Determinizer d(r, beam);
        if (d.run(out)) return true;
        if ((int)out->num_states() <= kMaxDetStates) return false;  // no complete path: shrinking the beam cannot help
    }
    return false;
}

void scale_graph_costs(CLat *lat, float scale) {
    for (auto &as : lat->arcs)
        for (CLatArc &a : as) a.w.g *= scale;
    for (size_t s = 0; s < lat->num_states(); s++)
        if (lat->is_final[s]) lat->final_w[s].g *= scale;
}

// ------------------------------------------------------------------------------------------------------------
// WordAlignLattice
// ------------------------------------------------------------------------------------------------------------
namespace {
constexpr int kSilTmp = -1, kPartialTmp = -2;  // temporary labels so that epsilon removal leaves these arcs alone

struct CompState {  // LatticeWordAligner::ComputationState (its weight is always One after Advance)
    std::vector<int> tids, words;
    bool operator<(const CompState &o) const { return tids != o.tids ? tids < o.tids : words < o.words; }
    bool empty() const { return tids.empty() && words.empty(); }
};

class WordAligner {
   public:
    WordAligner(const CLat &in, const LatticeCtx &ctx) : ctx_(ctx) {
        lat_ = in;
        // CreateSuperFinal: the only final weight left is One on one extra state
        const int sf = lat_.add_state();
        for (int s = 0; s < sf; s++)
            if (lat_.is_final[s]) {
                lat_.arcs[s].push_back(CLatArc{sf, 0, lat_.final_w[s], lat_.final_tids[s]});
                lat_.is_final[s] = 0;
            }
        lat_.is_final[sf] = 1;
        lat_.final_w[sf] = LatWeight{};
        lat_.final_tids[sf].clear();
    }

    void run(CLat *out) {
        out_ = out;
        *out = CLat();
        if (lat_.start < 0) return;
        out->start = state_for(lat_.start, CompState());
        while (!queue_.empty()) {
            auto item = queue_.back();
            queue_.pop_back();
            process(item.first, item.second);
        }
        remove_eps_local();
        for (auto &as : out->arcs)
            for (CLatArc &a : as)
                if (a.word < 0) a.word = 0;
    }

   private:
    int phone_of(int tid) const { return tid >= 0 && tid < (int)ctx_.tid2phone->size() ? (*ctx_.tid2phone)[tid] : -1; }
    int type_of(int phone) const { return phone >= 0 && phone < (int)ctx_.phone_type->size() ? (*ctx_.phone_type)[phone] : 0; }
    // chain topology of the model container: tid = 2*tstate+1 self-loop, 2*tstate+2 forward (= final) transition
    static bool is_final_tid(int tid) { return tid > 0 && (tid % 2) == 0; }
    static bool is_self_loop(int tid) { return tid > 0 && (tid % 2) == 1; }

    // number of leading tids that make up one complete phone, or 0 if its end cannot be decided yet (reorder = true:
    // the self-loops follow the final transition, so a phone is only known to be over when the next tid is in sight)
    static size_t phone_span(const std::vector<int> &t, size_t from) {
        size_t i = from;
        const size_t len = t.size();
        for (; i < len; i++)
            if (is_final_tid(t[i])) break;
        if (i == len) return 0;
        i++;
        while (i < len && is_self_loop(t[i])) i++;
        if (i == len) return 0;
        return i;
    }

    bool output_arc(CompState *c, CLatArc *arc) const {
        if (c->tids.empty()) return false;
        const int ty = type_of(phone_of(c->tids[0]));
        if (ty == 1) {  // OutputSilenceArc
            const size_t i = phone_span(c->tids, 0);
            if (!i) return false;
            *arc = CLatArc{-1, kSilTmp, LatWeight{}, std::vector<int>(c->tids.begin(), c->tids.begin() + i)};
            c->tids.erase(c->tids.begin(), c->tids.begin() + i);
            return true;
        }
        if (c->words.empty()) return false;
        size_t i = 0;
        if (ty == 5) {  // OutputOnePhoneWordArc
            i = phone_span(c->tids, 0);
            if (!i) return false;
        } else if (ty == 2) {  // OutputNormalWordArc: begin phone, word-internal phones, end phone
            i = phone_span(c->tids, 0);
            if (!i) return false;
            const size_t len = c->tids.size();
            while (i < len) {
                const int t2 = type_of(phone_of(c->tids[i]));
                if (t2 == 3 || t2 == 5) break;
                i++;
            }
            if (i == len) return false;
            i = phone_span(c->tids, i);
            if (!i) return false;
        } else {
            return false;
        }
        *arc = CLatArc{-1, c->words[0], LatWeight{}, std::vector<int>(c->tids.begin(), c->tids.begin() + i)};
        c->tids.erase(c->tids.begin(), c->tids.begin() + i);
        c->words.erase(c->words.begin());
        return true;
    }

    // OutputArcForce: at a final state whatever is pending goes out as one arc (silence, word, or partial word)
    void output_arc_force(CompState *c, CLatArc *arc) const {
        if (!c->tids.empty()) {
            const int ty = type_of(phone_of(c->tids[0]));
            int label;
            if (ty == 1) {
                label = kSilTmp;
            } else if (!c->words.empty()) {
                label = c->words[0];
                c->words.erase(c->words.begin());
            } else {
                label = kPartialTmp;
            }
            *arc = CLatArc{-1, label, LatWeight{}, c->tids};
            c->tids.clear();
        } else {
            *arc = CLatArc{-1, c->words[0], LatWeight{}, {}};
            c->words.erase(c->words.begin());
        }
    }

    int state_for(int in_state, const CompState &c) {
        auto key = std::make_pair(in_state, c);
        auto it = map_.find(key);
        if (it != map_.end()) return it->second;
        const int id = out_->add_state();
        map_.emplace(key, id);
        queue_.emplace_back(key, id);
        return id;
    }

    void process(const std::pair<int, CompState> &tuple, int out_state) {
        CompState c = tuple.second;
        CLatArc arc;
        if (output_arc(&c, &arc)) {
            arc.dst = state_for(tuple.first, c);
            out_->arcs[out_state].push_back(std::move(arc));
            return;
        }
        // ProcessFinal
        if (lat_.is_final[tuple.first]) {
            if (tuple.second.empty()) {
                out_->is_final[out_state] = 1;
                out_->final_w[out_state] = lat_.final_w[tuple.first];
            } else {
                CompState c2 = tuple.second;
                output_arc_force(&c2, &arc);
                arc.dst = state_for(tuple.first, c2);
                out_->arcs[out_state].push_back(std::move(arc));
            }
        }
        for (const CLatArc &ain : lat_.arcs[tuple.first]) {  // Advance: consume one input arc; its weight goes out on an epsilon arc
            CompState nx = tuple.second;
            nx.tids.insert(nx.tids.end(), ain.tids.begin(), ain.tids.end());
            if (ain.word != 0) nx.words.push_back(ain.word);
            const int d = state_for(ain.dst, nx);
            out_->arcs[out_state].push_back(CLatArc{d, 0, ain.w, {}});
        }
    }

    // RemoveEpsLocal: merge an arc with the arc(s) leaving its destination when one of the two is an epsilon and the
    // destination has a single arc in (pattern 1) or a single arc out (pattern 2); never increases the arc count.
    // (Kaldi also pushes weight for stochasticity; path weights are unchanged by that and it is omitted.)
    static bool can_combine(const CLatArc &a, const CLatArc &b, CLatArc *c) {
        if (a.word != 0 && b.word != 0) return false;
        c->word = a.word != 0 ? a.word : b.word;
        c->w = times_w(a.w, b.w);
        c->tids = a.tids;
        c->tids.insert(c->tids.end(), b.tids.begin(), b.tids.end());
        c->dst = b.dst;
        return true;
    }
    void remove_eps_local() {
        CLat &f = *out_;
        const int n = (int)f.num_states();
        const int dead = -7;
        std::vector<int> nin(n, 0), nout(n, 0);
        for (int s = 0; s < n; s++) {
            for (const CLatArc &a : f.arcs[s]) {
                nin[a.dst]++;
                nout[s]++;
            }
            if (f.is_final[s]) nout[s]++;
        }
        if (f.start >= 0) nin[f.start]++;
        for (int s = 0; s < n; s++) {
            for (size_t pos = 0; pos < f.arcs[s].size(); pos++) {
                CLatArc arc = f.arcs[s][pos];
                const int t = arc.dst;
                if (t == dead || t == s) continue;
                if (nin[t] == 1 && nout[t] > 1) {  // pattern 1
                    bool removed = false, kept = false;
                    std::vector<CLatArc> add;
                    for (CLatArc &nx : f.arcs[t]) {
                        if (nx.dst == dead) continue;
                        CLatArc c;
                        if (can_combine(arc, nx, &c)) {
                            removed = true;
                            nout[t]--;
                            nin[nx.dst]--;
                            nx.dst = dead;
                            add.push_back(std::move(c));
                        } else {
                            kept = true;
                        }
                    }
                    if (f.is_final[t]) {
                        if (arc.word == 0 && arc.tids.empty()) {  // CanCombineFinal: an epsilon arc into a final state
                            removed = true;
                            nout[t]--;
                            const LatWeight fw = times_w(arc.w, f.final_w[t]);
                            if (!f.is_final[s]) {
                                nout[s]++;
                                f.is_final[s] = 1;
                                f.final_w[s] = fw;
                            } else if (compare_w(fw, f.final_w[s]) == 1) {
                                f.final_w[s] = fw;
                            }
                            f.is_final[t] = 0;
                        } else {
                            kept = true;
                        }
                    }
                    if (removed && !kept) {
                        nout[s]--;
                        nin[t]--;
                        f.arcs[s][pos].dst = dead;
                    }
                    for (CLatArc &c : add) {
                        nout[s]++;
                        nin[c.dst]++;
                        f.arcs[s].push_back(std::move(c));
                    }
                } else if (nout[t] == 1) {  // pattern 2
                    bool del = false;
                    if (f.is_final[t]) {
                        if (arc.word == 0 && arc.tids.empty()) {
                            const LatWeight fw = times_w(arc.w, f.final_w[t]);
                            if (nin[t] == 1) f.is_final[t] = 0;
                            if (!f.is_final[s]) {
                                nout[s]++;
                                f.is_final[s] = 1;
                                f.final_w[s] = fw;
                            } else if (compare_w(fw, f.final_w[s]) == 1) {
                                f.final_w[s] = fw;
                            }
                            del = true;
                        }
                    } else {
                        CLatArc *nx = nullptr;
                        for (CLatArc &x : f.arcs[t])
                            if (x.dst != dead) {
                                nx = &x;
                                break;
                            }
                        CLatArc c;
                        if (nx && can_combine(arc, *nx, &c)) {
                            del = true;
                            if (nin[t] == 1) {
                                nout[t]--;
                                nin[nx->dst]--;
                                nx->dst = dead;
                            }
                            nout[s]++;
                            nin[c.dst]++;
                            f.arcs[s].push_back(std::move(c));
                        }
                    }
                    if (del) {
                        nout[s]--;
                        nin[t]--;
                        f.arcs[s][pos].dst = dead;
                    }
                }
            }
        }
        // drop the deleted arcs and the states that became unreachable
        std::vector<int> remap(n, -1);
        std::vector<int> stack;
        if (f.start >= 0) {
            remap[f.start] = 0;
            stack.push_back(f.start);
        }
        int cnt = f.start >= 0 ? 1 : 0;
        while (!stack.empty()) {
            const int s = stack.back();
            stack.pop_back();
            for (const CLatArc &a : f.arcs[s])
                if (a.dst != dead && remap[a.dst] < 0) {
                    remap[a.dst] = cnt++;
                    stack.push_back(a.dst);
                }
        }
        CLat g;
        for (int i = 0; i < cnt; i++) g.add_state();
        g.start = f.start >= 0 ? 0 : -1;
        for (int s = 0; s < n; s++) {
            if (remap[s] < 0) continue;
            const int ns = remap[s];
            g.is_final[ns] = f.is_final[s];
            g.final_w[ns] = f.final_w[s];
            for (CLatArc &a : f.arcs[s])
                if (a.dst != dead) {
                    a.dst = remap[a.dst];
                    g.arcs[ns].push_back(std::move(a));
                }
        }
        f = std::move(g);
    }

    LatticeCtx ctx_;
    CLat lat_;
    CLat *out_ = nullptr;
    std::map<std::pair<int, CompState>, int> map_;
    std::vector<std::pair<std::pair<int, CompState>, int>> queue_;
};
}  // namespace

void word_align_lattice(const CLat &in, const LatticeCtx &ctx, CLat *out) {
    WordAligner w(in, ctx);
    w.run(out);
}

// ------------------------------------------------------------------------------------------------------------
// MinimumBayesRisk (Xu, Povey, Mangu, Zhu: "Minimum Bayes Risk decoding and system combination based on a
// recursion for edit distance", as implemented by Kaldi lat/sausages.cc)
// ------------------------------------------------------------------------------------------------------------
namespace {
struct MbrArc {
    int word, start, end;  // 1-based nodes
    double loglike;
};
inline double log_add(double a, double b) {
    if (a == -kInfD) return b;
    if (b == -kInfD) return a;
    const double m = std::max(a, b);
    return m + std::log1p(std::exp(-std::fabs(a - b)));
}
inline double edit_l(int a, int b, bool penalize = false) {
    if (a == b) return 0.0;
    return penalize ? 1.0 + 1.0e-05 : 1.0;
}
}  // namespace

std::vector<WordSpan> mbr_one_best(const CLat &aligned) {
    std::vector<WordSpan> result;
    if (aligned.start < 0 || aligned.num_states() == 0) return result;
    // CreateSuperFinal + TopSort
    CLat lat = aligned;
    const int sf = lat.add_state();
    bool any_final = false;
    for (int s = 0; s < sf; s++)
        if (lat.is_final[s]) {
            lat.arcs[s].push_back(CLatArc{sf, 0, lat.final_w[s], lat.final_tids[s]});
            any_final = true;
        }
    if (!any_final) return result;
    const int n_all = (int)lat.num_states();
    auto arcs_of = [&](int s, auto f) {
        for (const CLatArc &a : lat.arcs[s]) f(a.dst);
    };
    std::vector<int> order = topo_order(n_all, lat.start, arcs_of);
    // the super-final state must come last
    {
        auto it = std::find(order.begin(), order.end(), sf);
        if (it == order.end()) return result;
        order.erase(it);
        order.push_back(sf);
    }
    const int N = (int)order.size();
    std::vector<int> node_of(n_all, 0);
    for (int i = 0; i < N; i++) node_of[order[i]] = i + 1;
    std::vector<double> state_times(N + 1, 0.0);
    std::vector<MbrArc> arcs;
    std::vector<std::vector<int>> pre(N + 1);
    for (int i = 0; i < N; i++) {
        const int s = order[i];
        for (const CLatArc &a : lat.arcs[s]) {
            const int e = node_of[a.dst];
            if (!e) continue;
            state_times[e] = state_times[i + 1] + (double)a.tids.size();
            pre[e].push_back((int)arcs.size());
            arcs.push_back(MbrArc{a.word, i + 1, e, -((double)a.w.g + (double)a.w.a)});
        }
    }
    // initial R: words of the best path
    std::vector<int> R;
    {
        std::vector<double> best(N + 1, kInfD);
        std::vector<int> back(N + 1, -1);
        best[1] = 0;
        for (int nn = 2; nn <= N; nn++)
            for (int ai : pre[nn]) {
                const double c = best[arcs[ai].start] - arcs[ai].loglike;
                if (c < best[nn]) {
                    best[nn] = c;
                    back[nn] = ai;
                }
            }
        for (int nn = N; nn > 1 && back[nn] >= 0; nn = arcs[back[nn]].start)
            if (arcs[back[nn]].word != 0) R.push_back(arcs[back[nn]].word);
        std::reverse(R.begin(), R.end());
    }
    std::vector<std::vector<std::pair<int, float>>> gamma_out;
    std::vector<std::vector<std::pair<float, float>>> times_out;
    std::vector<std::pair<float, float>> one_best_times;
    std::vector<float> one_best_conf;
    for (int counter = 0;; counter++) {
        {  // NormalizeEps: epsilons between all words and at both ends
            std::vector<int> r2;
            r2.push_back(0);
            for (int w : R)
                if (w != 0) {
                    r2.push_back(w);
                    r2.push_back(0);
                }
            R.swap(r2);
        }
        const int Q = (int)R.size();
        auto r = [&](int q) { return R[q - 1]; };
        // ---- AccStats ----
        std::vector<double> alpha(N + 1, 0.0);
        std::vector<std::vector<double>> alpha_dash(N + 1, std::vector<double>(Q + 1, 0.0)), beta_dash(N + 1, std::vector<double>(Q + 1, 0.0));
        std::vector<double> alpha_dash_arc(Q + 1), beta_dash_arc(Q + 1);
        std::vector<char> b_arc(Q + 1);
        // per position q: (word, gamma, tau_b, tau_e) of the few words aligned there; linear search beats a tree here
        struct Acc { int word; double g, tb, te; };
        std::vector<std::vector<Acc>> acc(Q + 1);
        auto add = [&](int q, int word, double g, double tb, double te) {
            for (Acc &x : acc[q])
                if (x.word == word) {
                    x.g += g;
                    x.tb += tb;
                    x.te += te;
                    return;
                }
            acc[q].push_back(Acc{word, g, tb, te});
        };
        alpha[1] = 0.0;
        alpha_dash[1][0] = 0.0;
        for (int q = 1; q <= Q; q++) alpha_dash[1][q] = alpha_dash[1][q - 1] + edit_l(0, r(q));
        for (int nn = 2; nn <= N; nn++) {
            double alpha_n = -kInfD;
            for (int ai : pre[nn]) alpha_n = log_add(alpha_n, alpha[arcs[ai].start] + arcs[ai].loglike);
            alpha[nn] = alpha_n;
            for (int ai : pre[nn]) {
                const MbrArc &arc = arcs[ai];
                const int s_a = arc.start, w_a = arc.word;
                const double p_a = arc.loglike;
                const double arc_post_f = std::exp(alpha[s_a] + p_a - alpha[nn]);
                for (int q = 0; q <= Q; q++) {
                    if (q == 0) {
                        alpha_dash_arc[q] = alpha_dash[s_a][q] + edit_l(w_a, 0, true);
                    } else {
                        const double a1 = alpha_dash[s_a][q - 1] + edit_l(w_a, r(q)), a2 = alpha_dash[s_a][q] + edit_l(w_a, 0, true),
                                     a3 = alpha_dash_arc[q - 1] + edit_l(0, r(q));
                        alpha_dash_arc[q] = std::min(a1, std::min(a2, a3));
                    }
                    alpha_dash[nn][q] += arc_post_f * alpha_dash_arc[q];
                }
            }
        }
        beta_dash[N][Q] = 1.0;
        for (int nn = N; nn >= 2; nn--) {
            for (int ai : pre[nn]) {
                const MbrArc &arc = arcs[ai];
                const int s_a = arc.start, w_a = arc.word;
                const double p_a = arc.loglike;
                alpha_dash_arc[0] = alpha_dash[s_a][0] + edit_l(w_a, 0, true);
                for (int q = 1; q <= Q; q++) {
                    const double a1 = alpha_dash[s_a][q - 1] + edit_l(w_a, r(q)), a2 = alpha_dash[s_a][q] + edit_l(w_a, 0, true),
                                 a3 = alpha_dash_arc[q - 1] + edit_l(0, r(q));
                    if (a1 <= a2) {
                        if (a1 <= a3) b_arc[q] = 1; else b_arc[q] = 3;
                    } else {
                        if (a2 <= a3) b_arc[q] = 2; else b_arc[q] = 3;
                    }
                    alpha_dash_arc[q] = std::min(a1, std::min(a2, a3));
                }
                std::fill(beta_dash_arc.begin(), beta_dash_arc.end(), 0.0);
                const double arc_post = std::exp(alpha[s_a] + p_a - alpha[nn]);
                for (int q = Q; q >= 1; q--) {
                    beta_dash_arc[q] += arc_post * beta_dash[nn][q];
                    switch (b_arc[q]) {
                        case 1:
                            beta_dash[s_a][q - 1] += beta_dash_arc[q];
                            add(q, w_a, beta_dash_arc[q], state_times[s_a] * beta_dash_arc[q], state_times[nn] * beta_dash_arc[q]);
                            break;
                        case 2:
                            beta_dash[s_a][q] += beta_dash_arc[q];
                            break;
                        case 3:
                            beta_dash_arc[q - 1] += beta_dash_arc[q];
                            add(q, 0, beta_dash_arc[q], state_times[s_a] * beta_dash_arc[q], state_times[s_a] * beta_dash_arc[q]);
                            break;
                    }
                }
                beta_dash_arc[0] += arc_post * beta_dash[nn][0];
                beta_dash[s_a][0] += beta_dash_arc[0];
            }
        }
        std::fill(beta_dash_arc.begin(), beta_dash_arc.end(), 0.0);
        for (int q = Q; q >= 1; q--) {
            beta_dash_arc[q] += beta_dash[1][q];
            beta_dash_arc[q - 1] += beta_dash_arc[q];
            add(q, 0, beta_dash_arc[q], state_times[1] * beta_dash_arc[q], state_times[1] * beta_dash_arc[q]);
        }
        gamma_out.assign(Q, {});
        times_out.assign(Q, {});
        for (int q = 1; q <= Q; q++) {
            auto &gq = gamma_out[q - 1];
            std::vector<Acc> &aq = acc[q];
            std::sort(aq.begin(), aq.end(), [](const Acc &x, const Acc &y) {  // GammaCompare on the float posteriors
                const float gx = (float)x.g, gy = (float)y.g;
                if (gx > gy) return true;
                if (gx < gy) return false;
                return x.word > y.word;
            });
            for (const Acc &x : aq) {
                const float g = (float)x.g;
                gq.emplace_back(x.word, g);
                times_out[q - 1].emplace_back((float)(x.tb / g), (float)(x.te / g));
            }
        }
        // ---- MbrDecode step ----
        double delta_Q = 0.0;
        one_best_times.clear();
        one_best_conf.clear();
        for (int q = 0; q < Q; q++) {
            const auto &g = gamma_out[q];
            if (g.empty()) continue;
            double old_gamma = 0, new_gamma = g[0].second;
            const int rq = R[q], rhat = g[0].first;
            for (auto &pr : g)
                if (pr.first == rq) old_gamma = pr.second;
            delta_Q += old_gamma - new_gamma;
            R[q] = rhat;
            if (R[q] != 0) {
                one_best_times.push_back(times_out[q][0]);
                const size_t i = one_best_times.size();
                if (i > 1 && one_best_times[i - 2].second > one_best_times[i - 1].first) {
                    // overlapping words: both share the union of their spans, split in proportion to their durations
                    const float prev_right = i > 2 ? one_best_times[i - 3].second : 0.0f;
                    const float left = std::max(prev_right, std::min(one_best_times[i - 2].first, one_best_times[i - 1].first));
                    const float right = std::max(one_best_times[i - 2].second, one_best_times[i - 1].second);
                    const float first_dur = one_best_times[i - 2].second - one_best_times[i - 2].first;
                    const float second_dur = one_best_times[i - 1].second - one_best_times[i - 1].first;
                    float mid = first_dur > 0 ? left + (right - left) * first_dur / (first_dur + second_dur) : left;
                    one_best_times[i - 2].first = left;
                    one_best_times[i - 2].second = one_best_times[i - 1].first = mid;
                    one_best_times[i - 1].second = right;
                }
                float conf = 0.f;
                for (auto &pr : g)
                    if (pr.first == R[q]) {
                        conf = pr.second;
                        break;
                    }
                one_best_conf.push_back(conf);
            }
        }
        if (delta_Q == 0 || counter > 100) break;
    }
    std::vector<int> words;
    for (int w : R)
        if (w != 0) words.push_back(w);
    for (size_t i = 0; i < words.size() && i < one_best_times.size(); i++)
        result.push_back(WordSpan{words[i], one_best_times[i].first, one_best_times[i].second, one_best_conf[i]});
    return result;
}

std::vector<WordSpan> lattice_to_words(const RawLattice &raw, const Model &m, float lattice_beam, float lm_scale) {
    LatticeCtx ctx{&m.graph, &m.tid2phone, &m.phone_type};
    CLat det, aligned;
    if (!determinize_lattice(raw, ctx, lattice_beam, &det)) return {};
    scale_graph_costs(&det, lm_scale);
    word_align_lattice(det, ctx, &aligned);
    return mbr_one_best(aligned);
}

}  // namespace vb
