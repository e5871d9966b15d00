// orc_lattice.cc — CPU restatement of the reference's lattice result chain.  TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// What the reference does with the raw lattice of a finished segment:
//   * inside the cudadecoder pipeline: DeterminizeLatticePhonePrunedWrapper(trans_model, &raw, lattice_beam, &clat,
//     det_opts) — the CompactLattice handed to the callback [REF src/batch_recognizer.cc:138-149]
//   * BatchRecognizer::PushLattice [REF src/batch_recognizer.cc:43-107]:
//       fst::ScaleLattice(fst::GraphLatticeScale(0.9), &clat)            [REF :45]
//       WordAlignLattice(clat, trans_model, winfo, 0, &aligned_lat)      [REF :47-48]
//       MinimumBayesRisk mbr(aligned_lat); GetOneBestConfidences / GetOneBest / GetOneBestTimes  [REF :50-54]
//       JSON / NLSML text                                                [REF :58-106]
//     (CPU twin: [REF src/recognizer.cc:430-482].)
// The algorithms live in alphacep/kaldi (absent from /root/reference — SURVEY.md §8c): lat/determinize-lattice-pruned.cc
// (LatticeDeterminizerPruned, DeterminizeLatticePhonePruned: phone pass then word pass), fstext/remove-eps-local-inl.h,
// lat/word-align-lattice.cc (LatticeWordAligner), lat/sausages.cc (MinimumBayesRisk).  They are restated here in their
// published structure — two determinization passes, initial/minimal subset lookup, RemoveEpsLocal with reweighting,
// std::map accumulators — independently of the engine's flat-array implementation (vosk-api_b200/csrc/vb_lattice.cc).
// Not restated: the max_mem / max_arcs retry loop of DeterminizeLatticePruned (never reached at these lattice sizes).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <queue>
#include <sstream>
#include <string>
#include <vector>

#include "oracle.h"

namespace {
const float kInf = std::numeric_limits<float>::infinity();
const double kInfD = std::numeric_limits<double>::infinity();
const float kDelta = 1.0f / 1024.0f;  // fst::kDelta

// ---------------------------------------------------------------- LatticeWeight (graph cost, acoustic cost)
struct LW {
    float g, a;
};
inline LW lw_one() { return LW{0.f, 0.f}; }
inline LW lw_zero() { return LW{kInf, kInf}; }
inline bool is_zero(const LW &w) { return w.g == kInf && w.a == kInf; }
inline bool lw_eq(const LW &x, const LW &y) { return x.g == y.g && x.a == y.a; }
inline LW times(const LW &x, const LW &y) { return LW{x.g + y.g, x.a + y.a}; }
inline LW divide(const LW &x, const LW &y) { return LW{x.g - y.g, x.a - y.a}; }
inline int compare(const LW &x, const LW &y) {  // 1 = x is better (lower total cost, then lower graph cost)
    const float fx = x.g + x.a, fy = y.g + y.a;
    if (fx < fy) return 1;
    if (fx > fy) return -1;
    if (x.g < y.g) return 1;
    if (x.g > y.g) return -1;
    return 0;
}
inline LW plus(const LW &x, const LW &y) { return compare(x, y) >= 0 ? x : y; }
inline double to_cost(const LW &w) { return (double)w.g + (double)w.a; }  // ConvertToCost
inline bool approx_equal(const LW &x, const LW &y, float delta) {
    if (x.g + x.a == y.g + y.a) return true;
    return std::fabs((x.g + x.a) - (y.g + y.a)) <= delta;
}

// ---------------------------------------------------------------- Lattice / CompactLattice
struct Arc {
    int il, ol;
    LW w;
    int next;
};
struct Lat {
    int start = -1;
    std::vector<std::vector<Arc>> arcs;
    std::vector<LW> fin;
    int add_state() {
        arcs.emplace_back();
        fin.push_back(lw_zero());
        return (int)arcs.size() - 1;
    }
    int n() const { return (int)arcs.size(); }
};
struct CW {  // CompactLatticeWeight
    LW w;
    std::vector<int> s;
};
inline CW cw_zero() { return CW{lw_zero(), {}}; }
inline CW cw_one() { return CW{lw_one(), {}}; }
inline bool is_zero(const CW &c) { return is_zero(c.w) && c.s.empty(); }
inline int compare_str(const std::vector<int> &x, const std::vector<int> &y) {  // the shorter string is better
    if (x.size() > y.size()) return -1;
    if (x.size() < y.size()) return 1;
    for (size_t i = 0; i < x.size(); i++) {
        if (x[i] < y[i]) return -1;
        if (x[i] > y[i]) return 1;
    }
    return 0;
}
inline int compare(const CW &x, const CW &y) {
    const int c = compare(x.w, y.w);
    return c ? c : compare_str(x.s, y.s);
}
inline CW plus(const CW &x, const CW &y) { return compare(x, y) >= 0 ? x : y; }
inline CW times(const CW &x, const CW &y) {
    if (is_zero(x.w) || is_zero(y.w)) return cw_zero();
    CW r{times(x.w, y.w), x.s};
    r.s.insert(r.s.end(), y.s.begin(), y.s.end());
    return r;
}
// Divide(w1, w2, DIVIDE_LEFT); *bad is set where Kaldi raises "Cannot divide" (length / data mismatch)
inline CW divide_left(const CW &x, const CW &y, bool *bad) {
    if (is_zero(x.w)) return cw_zero();
    CW r{divide(x.w, y.w), {}};
    if (y.s.size() > x.s.size() || !std::equal(y.s.begin(), y.s.end(), x.s.begin())) {
        *bad = true;
        r.s = x.s;
        return r;
    }
    r.s.assign(x.s.begin() + y.s.size(), x.s.end());
    return r;
}
struct CArc {
    int label;
    CW w;
    int next;
};
struct CLat {
    int start = -1;
    std::vector<std::vector<CArc>> arcs;
    std::vector<CW> fin;
    int add_state() {
        arcs.emplace_back();
        fin.push_back(cw_zero());
        return (int)arcs.size() - 1;
    }
    int n() const { return (int)arcs.size(); }
};

// fst::TopSort: states renumbered in reverse DFS finishing order (start first)
template <class L>
std::vector<int> top_order(const L &f) {
    const int n = f.n();
    std::vector<int> finish;
    std::vector<char> color(n, 0);
    auto dfs = [&](int root) {
        std::vector<std::pair<int, size_t>> st;
        st.emplace_back(root, 0);
        color[root] = 1;
        while (!st.empty()) {
            auto &top = st.back();
            if (top.second < f.arcs[top.first].size()) {
                const int d = f.arcs[top.first][top.second++].next;
                if (!color[d]) {
                    color[d] = 1;
                    st.emplace_back(d, 0);
                }
            } else {
                finish.push_back(top.first);
                st.pop_back();
            }
        }
    };
    if (f.start >= 0) dfs(f.start);
    for (int s = 0; s < n; s++)
        if (!color[s]) dfs(s);
    std::vector<int> order(n);  // order[old] = new
    for (int i = 0; i < n; i++) order[finish[n - 1 - i]] = i;
    return order;
}
template <class L>
void renumber(L *f, const std::vector<int> &order) {
    L g;
    for (int s = 0; s < f->n(); s++) g.add_state();
    for (int s = 0; s < f->n(); s++) {
        g.fin[order[s]] = f->fin[s];
        for (auto a : f->arcs[s]) {
            a.next = order[a.next];
            g.arcs[order[s]].push_back(a);
        }
    }
    g.start = f->start >= 0 ? order[f->start] : -1;
    *f = g;
}
template <class L>
void top_sort(L *f) {
    renumber(f, top_order(*f));
}
// fst::Connect: keep the states that are accessible from the start and from which a final state can be reached
template <class L, class IsFinal>
void connect(L *f, IsFinal is_final) {
    const int n = f->n();
    std::vector<char> acc(n, 0), coacc(n, 0);
    std::vector<int> st;
    if (f->start >= 0) {
        acc[f->start] = 1;
        st.push_back(f->start);
    }
    while (!st.empty()) {
        const int s = st.back();
        st.pop_back();
        for (auto &a : f->arcs[s])
            if (!acc[a.next]) {
                acc[a.next] = 1;
                st.push_back(a.next);
            }
    }
    std::vector<std::vector<int>> rev(n);
    for (int s = 0; s < n; s++)
        for (auto &a : f->arcs[s]) rev[a.next].push_back(s);
    for (int s = 0; s < n; s++)
        if (is_final(s)) {
            coacc[s] = 1;
            st.push_back(s);
        }
    while (!st.empty()) {
        const int s = st.back();
        st.pop_back();
        for (int p : rev[s])
            if (!coacc[p]) {
                coacc[p] = 1;
                st.push_back(p);
            }
    }
    std::vector<int> map(n, -1);
    int cnt = 0;
    for (int s = 0; s < n; s++)
        if (acc[s] && coacc[s]) map[s] = cnt++;
    L g;
    for (int i = 0; i < cnt; i++) g.add_state();
    for (int s = 0; s < n; s++) {
        if (map[s] < 0) continue;
        g.fin[map[s]] = f->fin[s];
        for (auto a : f->arcs[s])
            if (map[a.next] >= 0) {
                a.next = map[a.next];
                g.arcs[map[s]].push_back(a);
            }
    }
    g.start = f->start >= 0 ? map[f->start] : -1;
    *f = g;
}

// ---------------------------------------------------------------- LatticeStringRepository
class StringRepo {
   public:
    StringRepo() {
        parent_.push_back(-1);
        label_.push_back(0);
    }
    int empty() const { return 0; }
    int successor(int id, int label) {
        auto key = std::make_pair(id, label);
        auto it = succ_.find(key);
        if (it != succ_.end()) return it->second;
        parent_.push_back(id);
        label_.push_back(label);
        return succ_[key] = (int)parent_.size() - 1;
    }
    std::vector<int> to_vector(int id) const {
        std::vector<int> v;
        for (; id > 0; id = parent_[id]) v.push_back(label_[id]);
        std::reverse(v.begin(), v.end());
        return v;
    }
    int from_vector(const std::vector<int> &v) {
        int id = 0;
        for (int x : v) id = successor(id, x);
        return id;
    }
    int concatenate(int a, int b) {
        for (int x : to_vector(b)) a = successor(a, x);
        return a;
    }
    void reduce_to_common_prefix(int id, std::vector<int> *prefix) const {
        const std::vector<int> v = to_vector(id);
        size_t k = 0;
        while (k < prefix->size() && k < v.size() && v[k] == (*prefix)[k]) k++;
        prefix->resize(k);
    }
    int remove_prefix(int id, size_t n) {
        const std::vector<int> v = to_vector(id);
        return from_vector(std::vector<int>(v.begin() + n, v.end()));
    }

   private:
    std::vector<int> parent_, label_;
    std::map<std::pair<int, int>, int> succ_;
};

// ---------------------------------------------------------------- LatticeDeterminizerPruned
// Input: topologically sorted acceptor-like Lattice with the labels to determinize on as ilabel and the symbols that
// become the weight strings as olabel; arcs of a state sorted by ilabel.
class DeterminizerPruned {
   public:
    DeterminizerPruned(const Lat &ifst, double beam, float delta) : ifst_(ifst), beam_(beam), delta_(delta) {}

    void determinize() {
        initialize();
        while (!queue_.empty()) {
            Task *t = queue_.top();
            queue_.pop();
            process_transition(t->state, t->label, &t->subset);
            delete t;
        }
    }
    // compact form: one arc per determinized transition, the string in the weight
    void output(CLat *o) const {
        *o = CLat();
        if (out_.empty()) return;
        for (size_t s = 0; s < out_.size(); s++) o->add_state();
        o->start = 0;
        for (size_t s = 0; s < out_.size(); s++)
            for (const TempArc &t : out_[s].arcs) {
                CW w{t.weight, repo_.to_vector(t.string)};
                if (t.next < 0) o->fin[s] = w;
                else o->arcs[s].push_back(CArc{t.label, w, t.next});
            }
    }
    // non-compact form: the string is spelled out as olabels along a chain of new states, weight and ilabel on the first arc
    void output(Lat *o) const {
        *o = Lat();
        if (out_.empty()) return;
        for (size_t s = 0; s < out_.size(); s++) o->add_state();
        o->start = 0;
        for (size_t s = 0; s < out_.size(); s++)
            for (const TempArc &t : out_[s].arcs) {
                const std::vector<int> seq = repo_.to_vector(t.string);
                int cur = (int)s;
                if (t.next < 0) {
                    for (size_t i = 0; i < seq.size(); i++) {
                        const int nx = o->add_state();
                        o->arcs[cur].push_back(Arc{0, seq[i], i == 0 ? t.weight : lw_one(), nx});
                        cur = nx;
                    }
                    o->fin[cur] = seq.empty() ? t.weight : lw_one();
                } else {
                    for (size_t i = 0; i + 1 < seq.size(); i++) {
                        const int nx = o->add_state();
                        o->arcs[cur].push_back(Arc{i == 0 ? t.label : 0, seq[i], i == 0 ? t.weight : lw_one(), nx});
                        cur = nx;
                    }
                    o->arcs[cur].push_back(Arc{seq.size() <= 1 ? t.label : 0, seq.empty() ? 0 : seq.back(),
                                               seq.size() <= 1 ? t.weight : lw_one(), t.next});
                }
            }
    }

   private:
    struct Element {
        int state;
        int string;
        LW weight;
        bool operator<(const Element &o) const { return state < o.state; }
        bool operator>(const Element &o) const { return state > o.state; }
        bool operator!=(const Element &o) const { return state != o.state || string != o.string || !lw_eq(weight, o.weight); }
    };
    struct TempArc {
        int label;
        int string;
        int next;  // -1 = "this is the final weight"
        LW weight;
    };
    struct OutputState {
        std::vector<Element> minimal_subset;
        std::vector<TempArc> arcs;
        double forward_cost;
    };
    struct Task {
        int state;
        int label;
        std::vector<Element> subset;
        double priority_cost = kInfD;
    };
    struct TaskCompare {
        bool operator()(const Task *x, const Task *y) const { return x->priority_cost > y->priority_cost; }
    };
    typedef std::vector<std::pair<int, int>> SubsetKey;  // (state, string) per element; the weights are compared within delta

    static SubsetKey key_of(const std::vector<Element> &v) {
        SubsetKey k;
        for (const Element &e : v) k.emplace_back(e.state, e.string);
        return k;
    }
    bool subset_equal(const std::vector<Element> &x, const std::vector<Element> &y) const {
        if (x.size() != y.size()) return false;
        for (size_t i = 0; i < x.size(); i++)
            if (x[i].state != y[i].state || x[i].string != y[i].string || !approx_equal(x[i].weight, y[i].weight, delta_)) return false;
        return true;
    }
    int compare_ws(const LW &aw, int as, const LW &bw, int bs) const {
        const int c = compare(aw, bw);
        if (c) return c;
        if (as == bs) return 0;
        return compare_str(repo_.to_vector(as), repo_.to_vector(bs));
    }

    void compute_backward_weight() {
        const int n = ifst_.n();
        backward_.assign(n, kInfD);
        for (int s = n - 1; s >= 0; s--) {
            double c = to_cost(ifst_.fin[s]);
            for (const Arc &a : ifst_.arcs[s]) c = std::min(c, to_cost(a.w) + backward_[a.next]);
            backward_[s] = c;
        }
        if (ifst_.start >= 0) cutoff_ = backward_[ifst_.start] + beam_;
    }

    void epsilon_closure(std::vector<Element> *subset) {
        std::priority_queue<Element, std::vector<Element>, std::greater<Element>> queue;
        std::map<int, Element> cur;
        for (const Element &e : *subset) {
            queue.push(e);
            cur[e.state] = e;
        }
        bool replaced = false;
        while (!queue.empty()) {
            const Element elem = queue.top();
            queue.pop();
            if (replaced && cur[elem.state] != elem) continue;
            for (const Arc &arc : ifst_.arcs[elem.state]) {
                if (arc.il != 0) break;  // arcs are sorted on ilabel: no more epsilons
                if (is_zero(arc.w)) continue;
                Element nx;
                nx.state = arc.next;
                nx.weight = times(elem.weight, arc.w);
                auto it = cur.find(nx.state);
                if (it == cur.end()) {
                    nx.string = arc.ol == 0 ? elem.string : repo_.successor(elem.string, arc.ol);
                    cur[nx.state] = nx;
                    queue.push(nx);
                } else {
                    int comp = compare(nx.weight, it->second.weight);
                    if (comp == 0) {
                        nx.string = arc.ol == 0 ? elem.string : repo_.successor(elem.string, arc.ol);
                        comp = compare_ws(nx.weight, nx.string, it->second.weight, it->second.string);
                    }
                    if (comp == 1) {
                        nx.string = arc.ol == 0 ? elem.string : repo_.successor(elem.string, arc.ol);
                        it->second.string = nx.string;
                        it->second.weight = nx.weight;
                        queue.push(nx);
                        replaced = true;
                    }
                }
            }
        }
        subset->clear();
        for (auto &kv : cur) subset->push_back(kv.second);  // (std::map: already sorted by state)
    }

    void convert_to_minimal(std::vector<Element> *subset) const {
        std::vector<Element> keep;
        for (const Element &e : *subset) {
            bool ok = !is_zero(ifst_.fin[e.state]);
            for (const Arc &a : ifst_.arcs[e.state])
                if (a.il != 0) {
                    ok = true;
                    break;
                }
            if (ok) keep.push_back(e);
        }
        subset->swap(keep);
    }

    void normalize_subset(std::vector<Element> *elems, LW *tot, int *common) {
        if (elems->empty()) {
            *common = repo_.empty();
            *tot = lw_zero();
            return;
        }
        std::vector<int> prefix = repo_.to_vector((*elems)[0].string);
        LW w = (*elems)[0].weight;
        for (size_t i = 1; i < elems->size(); i++) {
            w = plus(w, (*elems)[i].weight);
            repo_.reduce_to_common_prefix((*elems)[i].string, &prefix);
        }
        for (Element &e : *elems) {
            e.weight = divide(e.weight, w);
            e.string = repo_.remove_prefix(e.string, prefix.size());
        }
        *common = repo_.from_vector(prefix);
        *tot = w;
    }

    static void make_subset_unique(std::vector<Element> *subset, const DeterminizerPruned *self) {
        // the subset is sorted on state; of several elements with one state the best (weight, then string) stays
        std::vector<Element> out;
        for (const Element &e : *subset) {
            if (!out.empty() && out.back().state == e.state) {
                if (self->compare_ws(e.weight, e.string, out.back().weight, out.back().string) == 1) out.back() = e;
            } else {
                out.push_back(e);
            }
        }
        subset->swap(out);
    }

    void process_final(int sid) {
        OutputState &st = out_[sid];
        int fstr = repo_.empty();
        LW fw = lw_zero();
        bool is_final = false;
        for (const Element &e : st.minimal_subset) {
            if (is_zero(ifst_.fin[e.state])) continue;
            const LW w = times(e.weight, ifst_.fin[e.state]);
            if (!is_final || compare_ws(w, e.string, fw, fstr) == 1) {
                is_final = true;
                fw = w;
                fstr = e.string;
            }
        }
        if (is_final && to_cost(fw) + st.forward_cost <= cutoff_) st.arcs.push_back(TempArc{0, fstr, -1, fw});
    }

    void process_transitions(int sid) {
        const std::vector<Element> minimal = out_[sid].minimal_subset;  // copy: out_ grows below
        std::vector<std::pair<int, Element>> all;
        for (const Element &e : minimal)
            for (const Arc &arc : ifst_.arcs[e.state]) {
                if (arc.il == 0 || is_zero(arc.w)) continue;
                Element nx;
                nx.state = arc.next;
                nx.weight = times(e.weight, arc.w);
                nx.string = arc.ol == 0 ? e.string : repo_.successor(e.string, arc.ol);
                all.emplace_back(arc.il, nx);
            }
        std::stable_sort(all.begin(), all.end(), [](const std::pair<int, Element> &x, const std::pair<int, Element> &y) {
            return x.first != y.first ? x.first < y.first : x.second.state < y.second.state;
        });
        size_t cur = 0;
        while (cur < all.size()) {
            Task *t = new Task;
            t->state = sid;
            t->label = all[cur].first;
            while (cur < all.size() && all[cur].first == t->label) {
                const Element &e = all[cur].second;
                t->subset.push_back(e);
                t->priority_cost = std::min(t->priority_cost, to_cost(e.weight) + backward_[e.state]);
                cur++;
            }
            t->priority_cost += out_[sid].forward_cost;
            if (t->priority_cost > cutoff_) {
                delete t;
            } else {
                make_subset_unique(&t->subset, this);
                queue_.push(t);
            }
        }
    }

    int minimal_to_state_id(const std::vector<Element> &subset, double forward_cost) {
        auto &bucket = minimal_hash_[key_of(subset)];
        for (int id : bucket)
            if (subset_equal(out_[id].minimal_subset, subset)) return id;
        const int id = (int)out_.size();
        out_.push_back(OutputState{subset, {}, forward_cost});
        bucket.push_back(id);
        process_final(id);
        process_transitions(id);
        return id;
    }

    int initial_to_state_id(const std::vector<Element> &subset_in, double forward_cost, LW *remaining, int *common) {
        auto &bucket = initial_hash_[key_of(subset_in)];
        for (const auto &entry : bucket)
            if (subset_equal(entry.first, subset_in)) {
                *remaining = entry.second.weight;
                *common = entry.second.string;
                return entry.second.state;
            }
        std::vector<Element> subset(subset_in);
        epsilon_closure(&subset);
        convert_to_minimal(&subset);
        Element elem;
        normalize_subset(&subset, &elem.weight, &elem.string);
        const int ans = minimal_to_state_id(subset, forward_cost + to_cost(elem.weight));
        *remaining = elem.weight;
        *common = elem.string;
        elem.state = ans;
        initial_hash_[key_of(subset_in)].emplace_back(subset_in, elem);  // (re-lookup: the map may have grown meanwhile)
        return ans;
    }

    void process_transition(int sid, int label, std::vector<Element> *subset) {
        double forward_cost = out_[sid].forward_cost;
        int common;
        LW tot;
        normalize_subset(subset, &tot, &common);
        forward_cost += to_cost(tot);
        LW next_tot;
        int next_common;
        const int next = initial_to_state_id(*subset, forward_cost, &next_tot, &next_common);
        common = repo_.concatenate(common, next_common);
        tot = times(tot, next_tot);
        out_[sid].arcs.push_back(TempArc{label, common, next, tot});
    }

    void initialize() {
        compute_backward_weight();
        if (ifst_.start < 0) return;
        // the start subset is not normalized (its weight / string would need a super-initial state)
        std::vector<Element> subset{Element{ifst_.start, repo_.empty(), lw_one()}};
        epsilon_closure(&subset);
        convert_to_minimal(&subset);
        out_.push_back(OutputState{subset, {}, 0.0});
        minimal_hash_[key_of(subset)].push_back(0);
        process_final(0);
        process_transitions(0);
    }

    const Lat &ifst_;
    double beam_, cutoff_ = kInfD;
    float delta_;
    std::vector<double> backward_;
    std::vector<OutputState> out_;
    std::map<SubsetKey, std::vector<int>> minimal_hash_;
    std::map<SubsetKey, std::vector<std::pair<std::vector<Element>, Element>>> initial_hash_;
    std::priority_queue<Task *, std::vector<Task *>, TaskCompare> queue_;
    mutable StringRepo repo_;
};

struct TransInfo {  // what the chain needs from the TransitionModel and WordBoundaryInfo
    const int *tid2phone;
    const unsigned char *tid_flags;  // bit 0: self-loop, bit 1: final (enters the HMM's final state), bit 2: leaves HMM state 0
    int num_tids;
    const int *phone_type;
    int num_phones;
    int phone(int tid) const { return tid > 0 && tid < num_tids ? tid2phone[tid] : 0; }
    // default (tid_flags == NULL): chain topology of the model generator, tid = 2*tstate+1 self-loop, 2*tstate+2 forward
    bool self_loop(int tid) const { return tid_flags ? (tid_flags[tid] & 1) != 0 : (tid > 0 && (tid % 2) == 1); }
    bool is_final(int tid) const { return tid_flags ? (tid_flags[tid] & 2) != 0 : (tid > 0 && (tid % 2) == 0); }
    bool from_state0(int tid) const { return tid_flags ? (tid_flags[tid] & 4) != 0 : tid > 0; }
    int type(int phone) const { return phone >= 0 && phone <= num_phones ? phone_type[phone] : 0; }
};

// DeterminizeLatticeInsertPhones: a phone label (first_phone_label + phone) where a phone begins — on the arc itself if it
// has no word, else on an extra arc behind it.  Arcs leaving the start state are skipped, as in Kaldi.
int insert_phones(const TransInfo &tm, Lat *f) {
    int highest = 0;
    for (auto &as : f->arcs)
        for (auto &a : as) highest = std::max(highest, a.il);
    const int first_phone_label = highest + 1;
    const int n0 = f->n();
    for (int s = 0; s < n0; s++) {
        if (s == f->start) continue;
        const size_t na = f->arcs[s].size();
        for (size_t k = 0; k < na; k++) {
            Arc arc = f->arcs[s][k];
            if (arc.ol != 0 && tm.from_state0(arc.ol) && !tm.self_loop(arc.ol)) {
                const int phone = tm.phone(arc.ol);
                if (arc.il == 0) {
                    arc.il = first_phone_label + phone;
                } else {
                    const int extra = f->add_state();
                    const int nx = arc.next;
                    arc.next = extra;
                    f->arcs[extra].push_back(Arc{first_phone_label + phone, 0, lw_one(), nx});
                }
            }
            f->arcs[s][k] = arc;
        }
    }
    return first_phone_label;
}
void delete_phones(int first_phone_label, Lat *f) {
    for (auto &as : f->arcs)
        for (auto &a : as)
            if (a.il >= first_phone_label) a.il = 0;
}
void arc_sort_ilabel(Lat *f) {
    for (auto &as : f->arcs) std::stable_sort(as.begin(), as.end(), [](const Arc &x, const Arc &y) { return x.il < y.il; });
}

// DeterminizeLatticePhonePrunedWrapper with the default options (phone_determinize, word_determinize, no minimize).
// raw: ilabel = transition-id, olabel = word, as GetRawLattice builds it.
void determinize_phone_pruned(const TransInfo &tm, Lat raw, double beam, bool phone_pass, CLat *out) {
    for (auto &as : raw.arcs)
        for (auto &a : as) std::swap(a.il, a.ol);  // Invert: words become the input labels
    top_sort(&raw);
    arc_sort_ilabel(&raw);
    if (phone_pass) {
        const int first_phone_label = insert_phones(tm, &raw);
        top_sort(&raw);
        arc_sort_ilabel(&raw);
        Lat det;
        {
            DeterminizerPruned d(raw, beam, kDelta);
            d.determinize();
            d.output(&det);
        }
        delete_phones(first_phone_label, &det);
        top_sort(&det);
        arc_sort_ilabel(&det);
        raw = det;
    }
    DeterminizerPruned d(raw, beam, kDelta);
    d.determinize();
    d.output(out);
    connect(out, [&](int s) { return !is_zero(out->fin[s]); });
}

// ---------------------------------------------------------------- RemoveEpsLocal on a CompactLattice (acceptor)
class RemoveEpsLocal {
   public:
    explicit RemoveEpsLocal(CLat *f, bool *divide_error) : f_(f), bad_(divide_error) {}
    void run() {
        if (f_->start < 0) return;
        dead_ = f_->add_state();
        const int n = f_->n();
        nin_.assign(n, 0);
        nout_.assign(n, 0);
        nin_[f_->start]++;
        for (int s = 0; s < n; s++) {
            if (!is_zero(f_->fin[s])) nout_[s]++;
            for (const CArc &a : f_->arcs[s]) {
                nin_[a.next]++;
                nout_[s]++;
            }
        }
        for (int s = 0; s < n; s++)
            for (size_t pos = 0; pos < f_->arcs[s].size(); pos++) remove_eps(s, pos);
        connect(f_, [&](int s) { return !is_zero(f_->fin[s]); });
    }

   private:
    static bool can_combine(const CArc &a, const CArc &b, CArc *c) {
        if (a.label != 0 && b.label != 0) return false;
        c->w = times(a.w, b.w);
        c->label = a.label != 0 ? a.label : b.label;
        c->next = b.next;
        return true;
    }
    static bool can_combine_final(const CArc &a, const CW &fin, CW *out) {
        if (a.label != 0) return false;
        *out = times(a.w, fin);
        return true;
    }
    void reweight(int s, size_t pos, const CW &rw) {
        CArc &arc = f_->arcs[s][pos];
        arc.w = times(arc.w, rw);
        const int t = arc.next;
        for (CArc &nx : f_->arcs[t])
            if (nx.next != dead_) nx.w = divide_left(nx.w, rw, bad_);
        if (!is_zero(f_->fin[t])) f_->fin[t] = divide_left(f_->fin[t], rw, bad_);
    }
    void pattern1(int s, size_t pos, CArc arc) {
        const int t = arc.next;
        CW removed = cw_zero(), kept = cw_zero();
        std::vector<CArc> add;
        for (CArc &nx : f_->arcs[t]) {
            if (nx.next == dead_) continue;
            CArc c;
            if (can_combine(arc, nx, &c)) {
                removed = plus(removed, nx.w);
                nout_[t]--;
                nin_[nx.next]--;
                nx.next = dead_;
                add.push_back(c);
            } else {
                kept = plus(kept, nx.w);
            }
        }
        if (!is_zero(f_->fin[t])) {
            CW nf;
            if (can_combine_final(arc, f_->fin[t], &nf)) {
                removed = plus(removed, f_->fin[t]);
                if (is_zero(f_->fin[s])) nout_[s]++;
                f_->fin[s] = plus(f_->fin[s], nf);
                nout_[t]--;
                f_->fin[t] = cw_zero();
            } else {
                kept = plus(kept, f_->fin[t]);
            }
        }
        if (!is_zero(removed)) {
            if (is_zero(kept)) {
                nout_[s]--;
                nin_[t]--;
                f_->arcs[s][pos].next = dead_;
            } else {
                const CW total = plus(removed, kept);
                const CW rw = divide_left(kept, total, bad_);
                if (getenv("ORC_DEBUG") && (!lw_eq(rw.w, lw_one()) || !rw.s.empty())) fprintf(stderr, "reweight %g %g strlen %zu\n", rw.w.g, rw.w.a, rw.s.size());
                reweight(s, pos, rw);
            }
        }
        for (const CArc &c : add) {
            nout_[s]++;
            nin_[c.next]++;
            f_->arcs[s].push_back(c);
        }
    }
    void pattern2(int s, size_t pos, CArc arc) {
        const int t = arc.next;
        bool del = false;
        if (!is_zero(f_->fin[t])) {
            CW nf;
            if (can_combine_final(arc, f_->fin[t], &nf)) {
                if (is_zero(f_->fin[s])) nout_[s]++;
                f_->fin[s] = plus(f_->fin[s], nf);
                del = true;
            }
        } else {
            size_t k = 0;
            while (f_->arcs[t][k].next == dead_) k++;
            CArc nx = f_->arcs[t][k], c;
            if (can_combine(arc, nx, &c)) {
                del = true;
                if (nin_[t] == 1) {
                    nout_[t]--;
                    nin_[nx.next]--;
                    f_->arcs[t][k].next = dead_;
                }
                nout_[s]++;
                nin_[c.next]++;
                f_->arcs[s].push_back(c);
            }
        }
        if (del) {
            nout_[s]--;
            nin_[t]--;
            f_->arcs[s][pos].next = dead_;
        }
    }
    void remove_eps(int s, size_t pos) {
        const CArc arc = f_->arcs[s][pos];
        const int t = arc.next;
        if (t == dead_ || t == s) return;
        if (nin_[t] == 1 && nout_[t] > 1) pattern1(s, pos, arc);
        else if (nout_[t] == 1) pattern2(s, pos, arc);
    }
    CLat *f_;
    bool *bad_;
    int dead_ = -1;
    std::vector<int> nin_, nout_;
};

// fst::CreateSuperFinal
int create_super_final(CLat *f) {
    std::vector<int> finals;
    for (int s = 0; s < f->n(); s++)
        if (!is_zero(f->fin[s])) finals.push_back(s);
    if (finals.size() == 1) {
        const CW &w = f->fin[finals[0]];
        if (lw_eq(w.w, lw_one()) && w.s.empty() && f->arcs[finals[0]].empty()) return finals[0];
    }
    const int sf = f->add_state();
    f->fin[sf] = cw_one();
    for (int s : finals) {
        f->arcs[s].push_back(CArc{0, f->fin[s], sf});
        f->fin[s] = cw_zero();
    }
    return sf;
}

// ---------------------------------------------------------------- LatticeWordAligner (reorder = true, silence / partial label 0)
class WordAligner {
   public:
    WordAligner(const CLat &lat, const TransInfo &tm, CLat *out) : lat_(lat), tm_(tm), out_(out) {
        // labels 0 for silence / partial words would vanish in RemoveEpsLocal: temporary unused labels are used instead
        int unused = 1;
        for (auto &as : lat_.arcs)
            for (auto &a : as) unused = std::max(unused, a.label + 1);
        partial_label_ = unused++;
        silence_label_ = unused;
        create_super_final(&lat_);
    }
    bool align() {
        *out_ = CLat();
        if (lat_.start < 0) return false;
        Tuple t0;
        t0.input_state = lat_.start;
        out_->start = state_for(t0);
        while (!queue_.empty()) {
            const std::pair<Tuple, int> item = queue_.back();
            queue_.pop_back();
            process(item.first, item.second);
        }
        RemoveEpsLocal(out_, &error_).run();
        for (auto &as : out_->arcs)
            for (CArc &a : as)
                if (a.label == silence_label_ || a.label == partial_label_) a.label = 0;
        if (getenv("ORC_DEBUG")) fprintf(stderr, "word-align error flag %d\n", (int)error_);
        return !error_;
    }

   private:
    struct Comp {  // ComputationState
        std::vector<int> tids, words;
        LW weight = lw_one();
        bool empty() const { return tids.empty() && words.empty(); }
    };
    struct Tuple {
        int input_state = -1;
        Comp c;
        bool operator<(const Tuple &o) const {
            if (input_state != o.input_state) return input_state < o.input_state;
            if (c.tids != o.c.tids) return c.tids < o.c.tids;
            if (c.words != o.c.words) return c.words < o.c.words;
            if (c.weight.g != o.c.weight.g) return c.weight.g < o.c.weight.g;
            return c.weight.a < o.c.weight.a;
        }
    };
    int state_for(const Tuple &t) {
        auto it = map_.find(t);
        if (it != map_.end()) return it->second;
        const int id = out_->add_state();
        map_[t] = id;
        queue_.emplace_back(t, id);
        return id;
    }
    void take(Comp *c, size_t i, int label, bool word, CArc *arc) const {
        *arc = CArc{label, CW{c->weight, std::vector<int>(c->tids.begin(), c->tids.begin() + i)}, -1};
        c->tids.erase(c->tids.begin(), c->tids.begin() + i);
        if (word) c->words.erase(c->words.begin());
        c->weight = lw_one();
    }
    bool output_silence_arc(Comp *c, CArc *arc) {
        if (c->tids.empty()) return false;
        const int phone = tm_.phone(c->tids[0]);
        if (tm_.type(phone) != 1) return false;
        const size_t len = c->tids.size();
        size_t i;
        for (i = 0; i < len; i++) {
            if (tm_.phone(c->tids[i]) != phone) error_ = true;
            if (tm_.is_final(c->tids[i])) break;
        }
        if (i == len) return false;
        i++;
        while (i < len && tm_.self_loop(c->tids[i])) i++;  // reorder: the self-loops follow the final transition
        if (i == len) return false;
        take(c, i, silence_label_, false, arc);
        return true;
    }
    bool output_one_phone_word_arc(Comp *c, CArc *arc) {
        if (c->tids.empty() || c->words.empty()) return false;
        const int phone = tm_.phone(c->tids[0]);
        if (tm_.type(phone) != 5) return false;
        const size_t len = c->tids.size();
        size_t i;
        for (i = 0; i < len; i++) {
            if (tm_.phone(c->tids[i]) != phone) error_ = true;
            if (tm_.is_final(c->tids[i])) break;
        }
        if (i == len) return false;
        i++;
        while (i < len && tm_.self_loop(c->tids[i])) i++;
        if (i == len) return false;
        take(c, i, c->words[0], true, arc);
        return true;
    }
    bool output_normal_word_arc(Comp *c, CArc *arc) {
        if (c->tids.empty() || c->words.empty()) return false;
        const int begin_phone = tm_.phone(c->tids[0]);
        if (tm_.type(begin_phone) != 2) return false;
        const size_t len = c->tids.size();
        size_t i;
        for (i = 0; i < len && !tm_.is_final(c->tids[i]); i++) {}
        if (i == len) return false;
        i++;
        while (i < len && tm_.self_loop(c->tids[i])) i++;
        if (i == len) return false;
        for (; i < len; i++) {  // word-internal phones up to the word-ending phone
            const int ty = tm_.type(tm_.phone(c->tids[i]));
            if (ty == 3 || ty == 5) break;
        }
        if (i == len) return false;
        const int final_phone = tm_.phone(c->tids[i]);
        for (; i < len; i++) {
            if (tm_.phone(c->tids[i]) != final_phone) error_ = true;
            if (tm_.is_final(c->tids[i])) break;
        }
        if (i == len) return false;
        i++;
        while (i < len && tm_.self_loop(c->tids[i])) i++;
        if (i == len) return false;
        take(c, i, c->words[0], true, arc);
        return true;
    }
    bool output_arc(Comp *c, CArc *arc) { return output_silence_arc(c, arc) || output_one_phone_word_arc(c, arc) || output_normal_word_arc(c, arc); }
    void output_arc_force(Comp *c, CArc *arc) {
        if (!c->tids.empty()) {
            const int phone = tm_.phone(c->tids[0]);
            if (tm_.type(phone) == 1) {
                take(c, c->tids.size(), silence_label_, false, arc);
            } else if (!c->words.empty()) {
                take(c, c->tids.size(), c->words[0], true, arc);
            } else {
                take(c, c->tids.size(), partial_label_, false, arc);
            }
        } else {
            take(c, 0, c->words[0], true, arc);
        }
    }
    void process(const Tuple &tuple_in, int out_state) {
        Tuple tuple = tuple_in;
        CArc arc;
        if (output_arc(&tuple.c, &arc)) {
            arc.next = state_for(tuple);
            out_->arcs[out_state].push_back(arc);
            return;
        }
        // ProcessFinal: after CreateSuperFinal the only final weight is One
        if (!is_zero(lat_.fin[tuple.input_state])) {
            if (tuple.c.empty()) {
                out_->fin[out_state] = plus(out_->fin[out_state], CW{tuple.c.weight, {}});
            } else {
                Tuple t2 = tuple;
                output_arc_force(&t2.c, &arc);
                arc.next = state_for(t2);
                out_->arcs[out_state].push_back(arc);
            }
        }
        for (const CArc &in : lat_.arcs[tuple.input_state]) {
            Tuple nx = tuple;
            // ComputationState::Advance: the arc's symbols are queued, its weight goes out on an epsilon arc
            nx.c.tids.insert(nx.c.tids.end(), in.w.s.begin(), in.w.s.end());
            if (in.label != 0) nx.c.words.push_back(in.label);
            const LW w = times(nx.c.weight, in.w.w);
            nx.c.weight = lw_one();
            nx.input_state = in.next;
            const int d = state_for(nx);
            out_->arcs[out_state].push_back(CArc{0, CW{w, {}}, d});
        }
    }

    CLat lat_;
    const TransInfo &tm_;
    CLat *out_;
    int silence_label_, partial_label_;
    bool error_ = false;
    std::map<Tuple, int> map_;
    std::vector<std::pair<Tuple, int>> queue_;
};

// ---------------------------------------------------------------- MinimumBayesRisk (lat/sausages.cc; decode_mbr, no silence)
struct MbrResult {
    std::vector<int> words;
    std::vector<std::pair<float, float>> times;
    std::vector<float> conf;
};
class Mbr {
   public:
    explicit Mbr(CLat clat) {
        if (clat.start < 0 || clat.n() == 0) return;
        bool any = false;
        for (int s = 0; s < clat.n(); s++) any = any || !is_zero(clat.fin[s]);
        if (!any) return;
        // PrepareLatticeAndInitStats
        create_super_final(&clat);
        top_sort(&clat);
        const int N = clat.n();
        {  // CompactLatticeStateTimes, then 1-based
            std::vector<int> t(N, -1);
            t[clat.start] = 0;
            for (int s = 0; s < N; s++)
                for (const CArc &a : clat.arcs[s])
                    if (t[s] >= 0) t[a.next] = t[s] + (int)a.w.s.size();
            state_times_.assign(N + 1, 0);
            for (int s = 0; s < N; s++) state_times_[s + 1] = t[s];
        }
        pre_.resize(N + 1);
        for (int n = 1; n <= N; n++)
            for (const CArc &a : clat.arcs[n - 1]) {
                MArc m;
                m.word = a.label;
                m.start = n;
                m.end = a.next + 1;
                m.loglike = -(a.w.w.g + a.w.w.a);
                pre_[m.end].push_back((int)arcs_.size());
                arcs_.push_back(m);
            }
        {  // R = words of the best path (fst::ShortestPath on the lattice without its alignments)
            std::vector<double> best(N + 1, kInfD);
            std::vector<int> back(N + 1, -1);
            best[1] = 0;
            for (int n = 2; n <= N; n++)
                for (int ai : pre_[n]) {
                    const double c = best[arcs_[ai].start] - (double)arcs_[ai].loglike;
                    if (c < best[n]) {
                        best[n] = c;
                        back[n] = ai;
                    }
                }
            for (int n = N; n > 1 && back[n] >= 0; n = arcs_[back[n]].start)
                if (arcs_[back[n]].word != 0) R_.push_back(arcs_[back[n]].word);
            std::reverse(R_.begin(), R_.end());
        }
        mbr_decode();
    }
    MbrResult result() const { return MbrResult{R_, one_best_times_, one_best_conf_}; }

   private:
    struct MArc {
        int word, start, end;
        float loglike;
    };
    static double l(int a, int b, bool penalize = false) { return a == b ? 0.0 : (penalize ? 1.0 + 1.0e-05 : 1.0); }
    int r(int q) const { return R_[q - 1]; }
    static double log_add(double x, double y) {
        if (x == -kInfD) return y;
        if (y == -kInfD) return x;
        const double hi = std::max(x, y), diff = std::min(x, y) - hi;
        return hi + std::log1p(std::exp(diff));
    }
    static void normalize_eps(std::vector<int> *v) {
        std::vector<int> w;
        w.push_back(0);
        for (int x : *v)
            if (x != 0) {
                w.push_back(x);
                w.push_back(0);
            }
        v->swap(w);
    }
    void acc_stats() {
        const int N = (int)pre_.size() - 1, Q = (int)R_.size();
        std::vector<double> alpha(N + 1, 0.0);
        std::vector<std::vector<double>> alpha_dash(N + 1, std::vector<double>(Q + 1, 0.0)), beta_dash(N + 1, std::vector<double>(Q + 1, 0.0));
        std::vector<double> alpha_dash_arc(Q + 1, 0.0), beta_dash_arc(Q + 1, 0.0);
        std::vector<char> b_arc(Q + 1, 0);
        std::vector<std::map<int, double>> gamma(Q + 1), tau_b(Q + 1), tau_e(Q + 1);
        // EditDistance
        alpha[1] = 0.0;
        alpha_dash[1][0] = 0.0;
        for (int q = 1; q <= Q; q++) alpha_dash[1][q] = alpha_dash[1][q - 1] + l(0, r(q));
        for (int n = 2; n <= N; n++) {
            double alpha_n = -kInfD;
            for (int ai : pre_[n]) alpha_n = log_add(alpha_n, alpha[arcs_[ai].start] + arcs_[ai].loglike);
            alpha[n] = alpha_n;
            for (int ai : pre_[n]) {
                const MArc &arc = arcs_[ai];
                const int s_a = arc.start, w_a = arc.word;
                const float p_a = arc.loglike;
                for (int q = 0; q <= Q; q++) {
                    if (q == 0) {
                        alpha_dash_arc[q] = alpha_dash[s_a][q] + l(w_a, 0, true);
                    } else {
                        const double a1 = alpha_dash[s_a][q - 1] + l(w_a, r(q)), a2 = alpha_dash[s_a][q] + l(w_a, 0, true),
                                     a3 = alpha_dash_arc[q - 1] + l(0, r(q));
                        alpha_dash_arc[q] = std::min(a1, std::min(a2, a3));
                    }
                    alpha_dash[n][q] += std::exp(alpha[s_a] + p_a - alpha[n]) * alpha_dash_arc[q];
                }
            }
        }
        beta_dash[N][Q] = 1.0;
        for (int n = N; n >= 2; n--)
            for (int ai : pre_[n]) {
                const MArc &arc = arcs_[ai];
                const int s_a = arc.start, w_a = arc.word;
                const float p_a = arc.loglike;
                alpha_dash_arc[0] = alpha_dash[s_a][0] + l(w_a, 0, true);
                for (int q = 1; q <= Q; q++) {
                    const double a1 = alpha_dash[s_a][q - 1] + l(w_a, r(q)), a2 = alpha_dash[s_a][q] + l(w_a, 0, true),
                                 a3 = alpha_dash_arc[q - 1] + l(0, r(q));
                    if (a1 <= a2) {
                        if (a1 <= a3) {
                            b_arc[q] = 1;
                            alpha_dash_arc[q] = a1;
                        } else {
                            b_arc[q] = 3;
                            alpha_dash_arc[q] = a3;
                        }
                    } else {
                        if (a2 <= a3) {
                            b_arc[q] = 2;
                            alpha_dash_arc[q] = a2;
                        } else {
                            b_arc[q] = 3;
                            alpha_dash_arc[q] = a3;
                        }
                    }
                }
                std::fill(beta_dash_arc.begin(), beta_dash_arc.end(), 0.0);
                for (int q = Q; q >= 1; q--) {
                    beta_dash_arc[q] += std::exp(alpha[s_a] + p_a - alpha[n]) * beta_dash[n][q];
                    switch (b_arc[q]) {
                        case 1:
                            beta_dash[s_a][q - 1] += beta_dash_arc[q];
                            gamma[q][w_a] += beta_dash_arc[q];
                            tau_b[q][w_a] += state_times_[s_a] * beta_dash_arc[q];
                            tau_e[q][w_a] += state_times_[n] * beta_dash_arc[q];
                            break;
                        case 2:
                            beta_dash[s_a][q] += beta_dash_arc[q];
                            break;
                        case 3:
                            beta_dash_arc[q - 1] += beta_dash_arc[q];
                            gamma[q][0] += beta_dash_arc[q];
                            tau_b[q][0] += state_times_[n] * beta_dash_arc[q];
                            tau_e[q][0] += state_times_[n] * beta_dash_arc[q];
                            break;
                    }
                }
                beta_dash_arc[0] += std::exp(alpha[s_a] + p_a - alpha[n]) * beta_dash[n][0];
                beta_dash[s_a][0] += beta_dash_arc[0];
            }
        std::fill(beta_dash_arc.begin(), beta_dash_arc.end(), 0.0);
        for (int q = Q; q >= 1; q--) {
            beta_dash_arc[q] += beta_dash[1][q];
            beta_dash_arc[q - 1] += beta_dash_arc[q];
            gamma[q][0] += beta_dash_arc[q];
            tau_b[q][0] += state_times_[1] * beta_dash_arc[q];
            tau_e[q][0] += state_times_[1] * beta_dash_arc[q];
        }
        gamma_.assign(Q, {});
        times_.assign(Q, {});
        for (int q = 1; q <= Q; q++) {
            for (auto &kv : gamma[q]) gamma_[q - 1].emplace_back(kv.first, (float)kv.second);
            std::sort(gamma_[q - 1].begin(), gamma_[q - 1].end(), [](const std::pair<int, float> &x, const std::pair<int, float> &y) {
                if (x.second > y.second) return true;
                if (x.second < y.second) return false;
                return x.first > y.first;
            });
            for (auto &pr : gamma_[q - 1]) {
                const double wb = tau_b[q][pr.first], we = tau_e[q][pr.first];
                times_[q - 1].emplace_back((float)(wb / pr.second), (float)(we / pr.second));
            }
        }
    }
    void mbr_decode() {
        for (int counter = 0;; counter++) {
            normalize_eps(&R_);
            acc_stats();
            double delta_Q = 0.0;
            one_best_times_.clear();
            one_best_conf_.clear();
            for (size_t q = 0; q < R_.size(); q++) {
                const auto &g = gamma_[q];
                double old_gamma = 0, new_gamma = g[0].second;
                const int rq = R_[q], rhat = g[0].first;
                for (auto &pr : g)
                    if (pr.first == rq) old_gamma = pr.second;
                delta_Q += old_gamma - new_gamma;
                R_[q] = rhat;
                if (R_[q] != 0) {
                    one_best_times_.push_back(times_[q][0]);
                    const size_t i = one_best_times_.size();
                    if (i > 1 && one_best_times_[i - 2].second > one_best_times_[i - 1].first) {
                        const float prev_right = i > 2 ? one_best_times_[i - 3].second : 0.0f;
                        const float left = std::max(prev_right, std::min(one_best_times_[i - 2].first, one_best_times_[i - 1].first));
                        const float right = std::max(one_best_times_[i - 2].second, one_best_times_[i - 1].second);
                        const float first_dur = one_best_times_[i - 2].second - one_best_times_[i - 2].first;
                        const float second_dur = one_best_times_[i - 1].second - one_best_times_[i - 1].first;
                        const float mid = first_dur > 0 ? left + (right - left) * first_dur / (first_dur + second_dur) : left;
                        one_best_times_[i - 2].first = left;
                        one_best_times_[i - 2].second = one_best_times_[i - 1].first = mid;
                        one_best_times_[i - 1].second = right;
                    }
                    float confidence = 0.0f;
                    for (auto &pr : g)
                        if (pr.first == R_[q]) {
                            confidence = pr.second;
                            break;
                        }
                    one_best_conf_.push_back(confidence);
                }
            }
            if (delta_Q == 0) break;
            if (counter > 100) break;
        }
        std::vector<int> w;
        for (int x : R_)
            if (x != 0) w.push_back(x);
        R_.swap(w);
    }

    std::vector<std::vector<int>> pre_;
    std::vector<MArc> arcs_;
    std::vector<int> state_times_;
    std::vector<int> R_;
    std::vector<std::vector<std::pair<int, float>>> gamma_;
    std::vector<std::vector<std::pair<float, float>>> times_;
    std::vector<std::pair<float, float>> one_best_times_;
    std::vector<float> one_best_conf_;
};

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

std::string dump_clat(const CLat &L) {
    std::string text;
    char b[160];
    snprintf(b, sizeof b, "S %d\n", L.start);
    text += b;
    auto tids = [](const std::vector<int> &t) {
        std::string s;
        for (size_t i = 0; i < t.size(); i++) s += (i ? "," : "") + std::to_string(t[i]);
        return s.empty() ? std::string("-") : s;
    };
    for (int s = 0; s < L.n(); s++) {
        for (const CArc &a : L.arcs[s]) {
            snprintf(b, sizeof b, "A %d %d %d %.9g %.9g ", s, a.next, a.label, a.w.w.g, a.w.w.a);
            text += b + tids(a.w.s) + "\n";
        }
        if (!is_zero(L.fin[s])) {
            snprintf(b, sizeof b, "F %d %.9g %.9g ", s, L.fin[s].w.g, L.fin[s].w.a);
            text += b + tids(L.fin[s].s) + "\n";
        }
    }
    return text;
}
}  // namespace

extern "C" char *orc_lattice_result(const OrcResultCtx *c, const float *arc_w, const unsigned char *tid_flags, int num_tids, int n_states,
                                    int start, int64_t n_links, const int64_t *src, const int64_t *dst, const int *arc, const float *acoustic,
                                    int64_t n_final, const int64_t *final_state, const float *final_cost, float lattice_beam, double lm_scale,
                                    float offset_seconds, int nlsml, int stage, int phone_pass) {
    // GetRawLattice: states = surviving tokens, arcs (ilabel = transition-id, olabel = word, (graph, acoustic))
    Lat raw;
    for (int s = 0; s < n_states; s++) raw.add_state();
    raw.start = start;
    for (int64_t k = 0; k < n_links; k++) {
        const int a = arc[k];
        raw.arcs[src[k]].push_back(Arc{c->arc_ilabel[a], c->arc_olabel[a], LW{arc_w[a], acoustic[k]}, (int)dst[k]});
    }
    for (int64_t k = 0; k < n_final; k++) raw.fin[final_state[k]] = LW{final_cost[k], 0.f};
    TransInfo tm{c->tid2phone, tid_flags, num_tids, c->phone_type, c->num_phones};
    std::string out;
    CLat clat;
    if (n_states > 0 && start >= 0) determinize_phone_pruned(tm, raw, (double)lattice_beam, phone_pass != 0, &clat);
    // fst::ScaleLattice(fst::GraphLatticeScale(0.9), &clat)
    // (ScaleTupleWeight with the double matrix {{lm_scale, 0}, {0, 1}}: evaluated in double, stored as float)
    auto scale = [&](LW *w) {
        if (w->g == kInf) return;
        const double g = (double)w->g, a = (double)w->a;
        w->g = (float)(lm_scale * g + 0.0 * a);
        w->a = (float)(0.0 * g + 1.0 * a);
    };
    for (auto &as : clat.arcs)
        for (CArc &a : as) scale(&a.w.w);
    for (CW &f : clat.fin)
        if (!is_zero(f)) scale(&f.w);
    if (stage == 1) {
        out = dump_clat(clat);
    } else {
        CLat aligned;
        WordAligner(clat, tm, &aligned).align();
        if (stage == 2) {
            out = dump_clat(aligned);
        } else {
            const MbrResult r = Mbr(aligned).result();
            const int size = (int)r.words.size();
            auto word = [&](int id) { return std::string(id >= 0 && id < c->num_words ? c->words[id] : ""); };
            if (stage == 5) {  // the raw MBR output: one "word-id begin end confidence" line per word (test hook)
                char b[128];
                for (int i = 0; i < size; i++) {
                    snprintf(b, sizeof b, "%d %.9g %.9g %.9g\n", r.words[i], r.times[i].first, r.times[i].second, r.conf[i]);
                    out += b;
                }
            } else if (nlsml) {
                std::stringstream ss, text;
                ss << "<?xml version=\"1.0\"?>\n";
                ss << "<result grammar=\"default\">\n";
                float confidence = 0.0;
                for (int i = 0; i < size; i++) {
                    if (i) text << " ";
                    confidence += r.conf[i];
                    text << word(r.words[i]);
                }
                confidence /= size;
                // (ostream << float prints as %g with precision 6; formatted with snprintf because float insertion into a
                // stringstream crashes in this library when it is built with -mavx2 and loaded into the test interpreter)
                char cbuf[64];
                snprintf(cbuf, sizeof cbuf, "%g", (double)confidence);
                ss << "<interpretation grammar=\"default\" confidence=\"" << cbuf << "\">\n";
                ss << "<input mode=\"speech\">" << text.str() << "</input>\n";
                ss << "<instance>" << text.str() << "</instance>\n";
                ss << "</interpretation>\n";
                ss << "</result>\n";
                out = ss.str();
            } else {
                std::string text;
                out = "{\n";
                if (size > 0) {
                    out += "  \"result\" : [";
                    for (int i = 0; i < size; i++) {
                        const double st = std::round(r.times[i].first) * 0.03 + offset_seconds;
                        const double en = std::round(r.times[i].second) * 0.03 + offset_seconds;
                        if (i) out += ", ";
                        out += "{\n      \"conf\" : " + std::to_string((double)r.conf[i]) + ",\n      \"end\" : " + std::to_string(en) +
                               ",\n      \"start\" : " + std::to_string(st) + ",\n      \"word\" : \"" + json_escape(word(r.words[i])) + "\"\n    }";
                        if (i) text += " ";
                        text += word(r.words[i]);
                    }
                    out += "],\n";
                }
                out += "  \"text\" : \"" + json_escape(text) + "\"\n}";
            }
        }
    }
    char *res = (char *)malloc(out.size() + 1);
    memcpy(res, out.c_str(), out.size() + 1);
    return res;
}
