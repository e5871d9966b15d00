// vb_lattice.cc — see vb_lattice.h.
//
// Everything here implements published Kaldi algorithms whose sources are NOT in /root/reference (they live in
// alphacep/kaldi); the reference only names them at its call sites [REF src/batch_recognizer.cc:45-54]
// and receives the determinized lattice from the cudadecoder pipeline [REF src/batch_recognizer.cc:138-149]:
//   * DeterminizeLatticePhonePrunedWrapper, default options: Invert, phone labels inserted where a phone begins (not on
//     the arcs leaving the start state), pruned determinization on phone + word labels, phone labels deleted, pruned
//     determinization on the word labels, Connect.  Pruning: a transition is expanded only while forward cost + best
//     completion stays within best + lattice_beam; subsets are matched on states, strings and total weight within 1/1024;
//   * fst::ScaleLattice(GraphLatticeScale(0.9)) (evaluated in double);
//   * WordAlignLattice (reorder = true, silence / partial-word label 0) with RemoveEpsLocal;
//   * MinimumBayesRisk (Xu, Povey, Mangu, Zhu; decode_mbr = true, print_silence = false).
// Layout: lattices are CSR arrays, strings live in a hash-consed trie, subsets / tuples in arenas with open-addressing
// lookup; all of it in one per-thread workspace that is reused from segment to segment.
// Stated differences: no max_mem / max_arcs retry loop (a lattice beyond kMaxDetStates output states fails and the caller
// falls back to the best path, logged and counted); RemoveEpsLocal's stochasticity re-weighting is omitted (it moves weight
// along a path without changing any path weight, and never triggers on a word-aligned lattice: the oracle, which has it,
// counts zero non-trivial re-weights).
#include "vb_lattice.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <queue>

namespace vb {

namespace {

#ifdef VB_LAT_PROF
#include <x86intrin.h>
static thread_local unsigned long long g_prof[16];
struct ProfScope {
    int k;
    unsigned long long t0;
    explicit ProfScope(int k_) : k(k_), t0(__rdtsc()) {}
    ~ProfScope() { g_prof[k] += __rdtsc() - t0; }
};
#define LATP(k) ProfScope prof_scope_##k(k)
#else
#define LATP(k)
#endif
constexpr float kInfF = std::numeric_limits<float>::infinity();
constexpr double kInfD = std::numeric_limits<double>::infinity();
constexpr int kMaxDetStates = 400000;
constexpr float kDelta = 1.0f / 1024.0f;
constexpr int kPhoneBase = 1 << 28;         // labels of the phone pass: kPhoneBase + phone
constexpr int kSilTmp = -1, kPartialTmp = -2;  // temporary labels so that epsilon removal leaves these arcs alone
constexpr int kDead = -7;                   // nextstate of an arc deleted by RemoveEpsLocal

// fst::Compare(LatticeWeight): 1 if (g1, a1) is better (lower total cost, then lower graph cost)
inline int compare_w(float g1, float a1, float g2, float a2) {
    const float f1 = g1 + a1, f2 = g2 + a2;
    if (f1 < f2) return 1;
    if (f1 > f2) return -1;
    if (g1 < g2) return 1;
    if (g1 > g2) return -1;
    return 0;
}
inline double cost_of(float g, float a) { return (double)g + (double)a; }  // ConvertToCost

// ------------------------------------------------------------------------------------------------------------
// flat lattice: input of a determinization pass (topologically numbered; arcs of a state sorted by il, epsilons first).
// An arc stands for a chain of arcs of the lattice Kaldi would see (one symbol each): interior states of such a chain —
// one arc in, one unlabelled arc out, not final — can hold no subset element that survives, so the chain is followed in
// one step; its weights are still added one by one, in order, so every float sum is the one Kaldi forms.
// ------------------------------------------------------------------------------------------------------------
struct Sym {  // one arc of the uncompressed lattice: its string symbol (0 = none) and weight
    int tid;
    float g, a;
};
struct FArc {
    int il;            // label determinized on (0 = epsilon), carried by the first arc of the chain
    int s_off, s_len;  // the chain, as a span of the pass's Sym pool (s_len >= 1)
    int next;
};
struct Edge {
    int src;
    FArc a;
};
struct FLat {
    int start = -1, n = 0;
    std::vector<int> off;
    std::vector<FArc> arcs;
    std::vector<float> fin_g, fin_a;  // fin_g = +inf: not final
    const std::vector<Sym> *syms = nullptr;
};
struct MEdge {  // arc of the word-aligned lattice as MinimumBayesRisk sees it
    int src, word, len, next;
    float loglike;
};

// open-addressing table hash -> id (entries are never removed; equality is the caller's)
struct IdTable {
    struct Slot {
        uint64_t hash;
        int id;  // -1 = free
    };
    std::vector<Slot> slots;  // (hash and id side by side: one cache line per probe)
    size_t used = 0, mask = 0;
    void reset(size_t cap_pow2) {
        if (slots.size() != cap_pow2) slots.assign(cap_pow2, Slot{0, -1});
        else std::fill(slots.begin(), slots.end(), Slot{0, -1});
        mask = cap_pow2 - 1;
        used = 0;
    }
    template <class Eq>
    int find(uint64_t h, Eq eq) const {
        for (size_t i = h & mask;; i = (i + 1) & mask) {
            const Slot &s = slots[i];
            if (s.id < 0) return -1;
            if (s.hash == h && eq(s.id)) return s.id;
        }
    }
    void insert(uint64_t h, int v) {
        if ((used + 1) * 2 > slots.size()) grow();
        size_t i = h & mask;
        while (slots[i].id >= 0) i = (i + 1) & mask;
        slots[i] = Slot{h, v};
        used++;
    }
    void grow() {
        std::vector<Slot> old;
        old.swap(slots);
        slots.assign(old.size() * 2, Slot{0, -1});
        mask = slots.size() - 1;
        used = 0;
        for (const Slot &s : old)
            if (s.id >= 0) insert(s.hash, s.id);
    }
};
inline uint64_t mix(uint64_t h, uint64_t v) {
    h = (h ^ v) * 0x9e3779b97f4a7c15ull;
    return h ^ (h >> 29);
}

// LatticeStringRepository: trie of symbol strings; id 0 = the empty string.  Strings that are stored in subsets are canonical
// (hash-consed from the root: equal content = equal id); the strings built while arcs are followed are plain appended
// nodes (no lookup), compared by content where it matters.
struct Repo {
    struct Node {
        int parent, label, depth, canon;
    };
    std::vector<Node> nodes;  // one record per string node: the walks touch one cache line per step
    IdTable tab;
    std::vector<int> tmp;
    void reset() {
        nodes.assign(1, Node{-1, 0, 0, 1});
        tab.reset(1 << 12);
    }
    int append(int id, int lab) {
        const int k = (int)nodes.size();
        nodes.push_back(Node{id, lab, nodes[id].depth + 1, 0});
        return k;
    }
    int succ(int id, int lab) {  // canonical successor of a canonical string
        const uint64_t h = mix((uint64_t)(uint32_t)id, (uint64_t)(uint32_t)lab);
        const int f = tab.find(h, [&](int k) { return nodes[k].parent == id && nodes[k].label == lab; });
        if (f >= 0) return f;
        const int k = append(id, lab);
        nodes[k].canon = 1;
        tab.insert(h, k);
        return k;
    }
    // a node of x's ancestry that spells the longest common prefix of x and y (by content)
    int common_prefix(int x, int y) const {
        while (nodes[x].depth > nodes[y].depth) x = nodes[x].parent;
        while (nodes[y].depth > nodes[x].depth) y = nodes[y].parent;
        int ans = x;
        while (x != y) {
            if (nodes[x].label != nodes[y].label) ans = nodes[x].parent;
            x = nodes[x].parent;
            y = nodes[y].parent;
        }
        return ans;
    }
    int remove_prefix(int id, int n) {  // canonical id of the string without its first n symbols
        tmp.clear();
        int r = 0;
        if (n == 0) {
            // nothing to remove: only the symbols appended on top of the deepest canonical ancestor need a lookup (a canonical
            // node's ancestors are canonical, the root is)
            int k = id;
            for (; !nodes[k].canon; k = nodes[k].parent) tmp.push_back(nodes[k].label);
            r = k;
        } else {
            for (int k = id; nodes[k].depth > n; k = nodes[k].parent) tmp.push_back(nodes[k].label);
        }
        for (size_t i = tmp.size(); i-- > 0;) r = succ(r, tmp[i]);
        return r;
    }
    int concat(int x, int y) {  // (not canonical: only the content of an output string matters)
        if (y == 0) return x;
        if (x == 0) return y;
        tmp.clear();
        for (int k = y; k > 0; k = nodes[k].parent) tmp.push_back(nodes[k].label);
        for (size_t i = tmp.size(); i-- > 0;) x = append(x, tmp[i]);
        return x;
    }
    void to_vec(int id, std::vector<int> *out) const {
        const size_t base = out->size();
        out->resize(base + nodes[id].depth);
        for (int k = id, i = nodes[id].depth; k > 0; k = nodes[k].parent) (*out)[base + --i] = nodes[k].label;
    }
    // LatticeDeterminizerPruned::Compare on strings: the shorter one is better, then the lexicographically larger (sic)
    int compare(int x, int y) const {
        if (x == y) return 0;
        if (nodes[x].depth > nodes[y].depth) return -1;
        if (nodes[x].depth < nodes[y].depth) return 1;
        int res = 0;  // decided by the first differing symbol = the mismatch closest to the root
        while (x != y) {
            if (nodes[x].label != nodes[y].label) res = nodes[x].label < nodes[y].label ? -1 : 1;
            x = nodes[x].parent;
            y = nodes[y].parent;
        }
        return res;
    }
};

struct Elem {
    int state, str;
    float g, a;
};
struct TArc {  // output transition of a determinization pass; next = -1: this is the final weight of src
    int src, label, str, next;
    float g, a;
};

// compact lattice in CSR form (output of the word pass; input of the word aligner)
struct CArcF {
    int label;
    float g, a;
    int t_off, t_len;  // transition-id string in CLatF::tids
    int next;
};
struct CLatF {
    int start = -1, n = 0;
    std::vector<int> off;
    std::vector<CArcF> arcs;
    std::vector<char> is_final;
    std::vector<float> fin_g, fin_a;
    std::vector<int> fin_off, fin_len;
    std::vector<int> tids;
};

// word-aligned lattice: arcs of a state form a linked list in insertion order (RemoveEpsLocal appends while it iterates)
struct AArc {
    int label;
    float g, a;
    int t_off, t_len;  // span in Workspace::seq
    int next, link;
};
struct ALat {
    int start = -1;
    std::vector<AArc> arcs;
    std::vector<int> head, tail;
    std::vector<char> is_final;
    std::vector<float> fin_g, fin_a;
    int add_state() {
        head.push_back(-1);
        tail.push_back(-1);
        is_final.push_back(0);
        fin_g.push_back(0.f);
        fin_a.push_back(0.f);
        return (int)head.size() - 1;
    }
    void add_arc(int s, const AArc &a) {
        const int k = (int)arcs.size();
        arcs.push_back(a);
        arcs[k].link = -1;
        if (tail[s] < 0) head[s] = k;
        else arcs[tail[s]].link = k;
        tail[s] = k;
    }
    void clear() {
        start = -1;
        arcs.clear();
        head.clear();
        tail.clear();
        is_final.clear();
        fin_g.clear();
        fin_a.clear();
    }
    int n() const { return (int)head.size(); }
};

struct Workspace {
    // build_flat
    std::vector<int> csr_off, csr_idx, order, dfs_stack, dfs_pos;
    std::vector<char> color;
    std::vector<Edge> edges, edges2;
    std::vector<MEdge> medges;
    std::vector<Sym> sym_in, sym1, sym2;
    std::vector<float> fin_g, fin_a;
    FLat lat1, lat2;
    // determinizer
    Repo repo;
    std::vector<Elem> pool, cur, sub, allv;
    std::vector<int> all_label;
    std::vector<double> all_prio, bw_first;
    std::vector<int> heap, stamp, done, slot, idx;
    std::vector<char> keepable;
    std::vector<double> backward;
    std::vector<TArc> tarcs;
    IdTable min_tab, init_tab;
    std::vector<int> init1, min1;  // single-element subsets by state: index + 1 into the initial-subset records / output states
    std::vector<int> tvec;
    CLatF clat;
    // aligner
    std::vector<int> seq, wseq;
    ALat ali;
    IdTable tup_tab;
    std::vector<int> nin, nout, remap, stack;
    // mbr
    std::vector<double> alpha, alpha_dash, beta_dash, post, mbr_cq, mbr_m12a, mbr_m12b, mbr_vala, mbr_valb;
    std::vector<char> b_arc;
    std::vector<int> lo, hi;
};
Workspace &workspace() {
    static thread_local Workspace w;
    return w;
}

// Chain compression: an edge absorbs the arcs behind it as long as it ends in an interior state (exactly one arc in, one
// unlabelled arc out, not final, not the start).  The symbols / weights of the absorbed arcs are appended to its span.
static int g_comp_pass = 0;
void compress_chains(Workspace &ws, int n, int start, const std::vector<Edge> &in, const std::vector<float> &fin_g,
                     const std::vector<Sym> &syms_in, std::vector<Edge> *out, std::vector<Sym> *syms_out) {
    LATP(7);
    std::vector<int> &indeg = ws.nin, &outdeg = ws.nout, &out_edge = ws.remap;
    indeg.assign(n, 0);
    outdeg.assign(n, 0);
    out_edge.assign(n, -1);
    for (size_t k = 0; k < in.size(); k++) {
        indeg[in[k].a.next]++;
        outdeg[in[k].src]++;
        out_edge[in[k].src] = (int)k;
    }
    const bool off_ = getenv("VB_NOCOMP") && atoi(getenv("VB_NOCOMP")) == g_comp_pass;
    auto interior = [&](int x) {
        return !off_ && x != start && indeg[x] == 1 && outdeg[x] == 1 && fin_g[x] == kInfF && in[out_edge[x]].a.il == 0;
    };
    out->clear();
    syms_out->clear();
    for (const Edge &e : in) {
        if (interior(e.src)) continue;  // absorbed by the edge that leads here
        Edge o = e;
        o.a.s_off = (int)syms_out->size();
        syms_out->insert(syms_out->end(), syms_in.begin() + e.a.s_off, syms_in.begin() + e.a.s_off + e.a.s_len);
        while (interior(o.a.next)) {
            const FArc &nx = in[out_edge[o.a.next]].a;
            syms_out->insert(syms_out->end(), syms_in.begin() + nx.s_off, syms_in.begin() + nx.s_off + nx.s_len);
            o.a.next = nx.next;
        }
        o.a.s_len = (int)syms_out->size() - o.a.s_off;
        out->push_back(o);
    }
}

// Topological numbering (reverse DFS finishing order from the start state; unreachable states are dropped) and CSR
// construction with the arcs of a state ordered by il.  fin_g/fin_a are indexed by the old state ids.
bool build_flat(Workspace &ws, int n_old, int start_old, const std::vector<Edge> &edges, const std::vector<float> &fin_g,
                const std::vector<float> &fin_a, const std::vector<Sym> *syms, FLat *out) {
    LATP(6);
    out->start = -1;
    out->n = 0;
    out->arcs.clear();
    out->syms = syms;
    if (start_old < 0 || start_old >= n_old) return false;
    std::vector<int> &off = ws.csr_off, &idx = ws.csr_idx, &order = ws.order;
    off.assign(n_old + 1, 0);
    for (const Edge &e : edges) off[e.src + 1]++;
    for (int s = 0; s < n_old; s++) off[s + 1] += off[s];
    idx.resize(edges.size());
    std::vector<int> &st = ws.dfs_stack, &pos = ws.dfs_pos, &finish = ws.stack;
    pos.assign(off.begin(), off.end() - 1);
    for (size_t k = 0; k < edges.size(); k++) idx[pos[edges[k].src]++] = (int)k;
    order.assign(n_old, -1);
    ws.color.assign(n_old, 0);
    st.clear();
    finish.clear();
    pos.assign(off.begin(), off.end() - 1);
    st.push_back(start_old);
    ws.color[start_old] = 1;
    while (!st.empty()) {
        const int s = st.back();
        if (pos[s] < off[s + 1]) {
            const int d = edges[idx[pos[s]++]].a.next;
            if (ws.color[d] == 1) return false;  // a cycle: not a lattice
            if (!ws.color[d]) {
                ws.color[d] = 1;
                st.push_back(d);
            }
        } else {
            ws.color[s] = 2;
            finish.push_back(s);
            st.pop_back();
        }
    }
    const int n = (int)finish.size();
    for (int i = 0; i < n; i++) order[finish[n - 1 - i]] = i;
    out->n = n;
    out->start = order[start_old];
    out->off.assign(n + 1, 0);
    for (const Edge &e : edges)
        if (order[e.src] >= 0) out->off[order[e.src] + 1]++;
    for (int s = 0; s < n; s++) out->off[s + 1] += out->off[s];
    out->arcs.resize(out->off[n]);
    pos.assign(out->off.begin(), out->off.end() - 1);
    for (const Edge &e : edges) {
        const int s = order[e.src];
        if (s < 0) continue;
        FArc a = e.a;
        a.next = order[a.next];
        // insertion by il keeps the edge order among equal labels
        int p = pos[s]++;
        while (p > out->off[s] && out->arcs[p - 1].il > a.il) {
            out->arcs[p] = out->arcs[p - 1];
            p--;
        }
        out->arcs[p] = a;
    }
    out->fin_g.assign(n, kInfF);
    out->fin_a.assign(n, kInfF);
    for (int s = 0; s < n_old; s++)
        if (order[s] >= 0 && s < (int)fin_g.size()) {
            out->fin_g[order[s]] = fin_g[s];
            out->fin_a[order[s]] = fin_a[s];
        }
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// LatticeDeterminizerPruned
// ------------------------------------------------------------------------------------------------------------
class Determinizer {
   public:
    Determinizer(Workspace &ws, const FLat &in, double beam) : ws_(ws), in_(in), sy_(in.syms->data()), beam_(beam), repo_(ws.repo) {}

    bool run() {
        LATP(5);
        repo_.reset();
        ws_.pool.clear();
        ws_.tarcs.clear();
        states_.clear();
        ws_.min_tab.reset(1 << 12);
        ws_.init_tab.reset(1 << 12);
        inits_.clear();
        const int n = in_.n;
        // most subsets (about nine in ten on speech lattices) hold one element; normalized, such a subset is {state, empty string,
        // One}: it is looked up by its state in an array, not through the hash tables
        ws_.init1.assign((size_t)n + in_.arcs.size(), 0);
        ws_.min1.assign((size_t)n + in_.arcs.size(), 0);
        if (n == 0 || in_.start < 0) return false;
        // ComputeBackwardWeight (per arc of the uncompressed lattice: the chain is walked from its end)
        ws_.backward.assign(n, kInfD);
        ws_.keepable.assign(n, 0);
        ws_.bw_first.resize(in_.arcs.size());
        for (int s = n - 1; s >= 0; s--) {
            double c = in_.fin_g[s] == kInfF ? kInfD : cost_of(in_.fin_g[s], in_.fin_a[s]);
            bool keep = in_.fin_g[s] != kInfF;
            for (int k = in_.off[s]; k < in_.off[s + 1]; k++) {
                const FArc &a = in_.arcs[k];
                const Sym *sy = sy_ + a.s_off;
                double b = ws_.backward[a.next];
                for (int j = a.s_len - 1; j >= 1; j--) b = cost_of(sy[j].g, sy[j].a) + b;
                ws_.bw_first[k] = b;  // the backward cost of the state behind the chain's first arc
                c = std::min(c, cost_of(sy[0].g, sy[0].a) + b);
                keep = keep || a.il != 0;
            }
            ws_.backward[s] = c;
            ws_.keepable[s] = keep;
        }
        if (ws_.backward[in_.start] == kInfD) return false;
        cutoff_ = ws_.backward[in_.start] + beam_;
        ws_.stamp.assign(n, 0);
        ws_.done.assign(n, 0);
        ws_.slot.assign(n, 0);
        epoch_ = 0;
        // the start subset is not normalized (its weight and string stay inside it)
        std::vector<Elem> &sub = ws_.sub;
        sub.clear();
        sub.push_back(Elem{in_.start, 0, 0.f, 0.f});
        closure(&sub);
        for (Elem &e : sub) e.str = repo_.remove_prefix(e.str, 0);  // canonical strings, so that an equal subset can match
        new_state(sub, 0.0);
        while (!tasks_.empty()) {
            if ((int)states_.size() > kMaxDetStates) return false;
            const Task t = tasks_.top();
            tasks_.pop();
            process_transition(t);
        }
        return true;
    }
    int num_states() const { return (int)states_.size(); }

   private:
    struct OutState {
        int off, len;
        double fwd;
    };
    struct Task {
        double prio;
        int src, label, off, len;
        bool operator<(const Task &o) const { return prio > o.prio; }  // min-heap on the priority cost
    };
    struct Init {  // an initial (pre-closure) subset that has been seen, and where it leads
        int off, len, state, str;
        float g, a;
    };

    static uint64_t hash_of(const Elem *e, int n) {
        uint64_t h = (uint64_t)n;
        for (int i = 0; i < n; i++) h = mix(h, ((uint64_t)(uint32_t)e[i].state << 32) | (uint32_t)e[i].str);
        return h;
    }
    static bool same_subset(const Elem *x, const Elem *y, int n) {
        for (int i = 0; i < n; i++) {
            if (x[i].state != y[i].state || x[i].str != y[i].str) return false;
            const float fx = x[i].g + x[i].a, fy = y[i].g + y[i].a;
            if (fx != fy && !(std::fabs(fx - fy) <= kDelta)) return false;
        }
        return true;
    }
    int compare_ws(float g1, float a1, int s1, float g2, float a2, int s2) const {
        const int c = compare_w(g1, a1, g2, a2);
        return c ? c : repo_.compare(s1, s2);
    }
    // Times(elem.weight, arc.weight) arc by arc along the chain (from its `from`-th arc on); the string is built only when needed
    inline void walk_weight(const FArc &arc, int from, float *g, float *a) const {
        const Sym *sy = sy_ + arc.s_off;
        for (int j = from; j < arc.s_len; j++) {
            *g += sy[j].g;
            *a += sy[j].a;
        }
    }
    inline int walk_string(const FArc &arc, int from, int str) {
        const Sym *sy = sy_ + arc.s_off;
        for (int j = from; j < arc.s_len; j++)
            if (sy[j].tid) str = repo_.append(str, sy[j].tid);
        return str;
    }

    // EpsilonClosure + ConvertToMinimal: follow the arcs without a label; per state keep the best (weight, string); drop
    // the elements whose state neither is final nor has a labelled arc.  States are numbered topologically, so a state
    // popped from the min-heap has its final element.  An element with state >= n sits inside the chain of arc
    // (state - n), behind its first (labelled) arc: that is where Kaldi's initial subsets live, and what they are keyed on.
    // The result is sorted by state.
    void closure(std::vector<Elem> *sub) {
        LATP(1);
        if (sub->size() == 1 && (*sub)[0].state < in_.n) {
            // one element whose state has no epsilon arc (they sort first): the closure is the element itself
            const int s0 = (*sub)[0].state;
            if (in_.off[s0] == in_.off[s0 + 1] || in_.arcs[in_.off[s0]].il != 0) {
                if (!ws_.keepable[s0]) sub->clear();
                return;
            }
        }
        std::vector<Elem> &cur = ws_.cur;
        std::vector<int> &heap = ws_.heap;
        cur.clear();
        heap.clear();
        epoch_++;
        const int n = in_.n;
        // one relaxation: the element (weight ng/na, string made by mk()) arrives at state `to`
        auto relax = [&](int to, float ng, float na, auto mk) {
            if (ws_.stamp[to] != epoch_) {
                ws_.stamp[to] = epoch_;
                ws_.slot[to] = (int)cur.size();
                cur.push_back(Elem{to, mk(), ng, na});
                heap.push_back(to);
                std::push_heap(heap.begin(), heap.end(), std::greater<int>());
            } else {
                Elem &old = cur[ws_.slot[to]];
                int c = compare_w(ng, na, old.g, old.a);
                int nstr = -1;
                if (c == 0) {
                    nstr = mk();
                    c = repo_.compare(nstr, old.str);
                }
                if (c == 1) {
                    if (nstr < 0) nstr = mk();
                    old.str = nstr;
                    old.g = ng;
                    old.a = na;
                }
            }
        };
        for (const Elem &e : *sub) {
            if (e.state >= n) continue;
            ws_.stamp[e.state] = epoch_;
            ws_.slot[e.state] = (int)cur.size();
            cur.push_back(e);
            heap.push_back(e.state);
        }
        std::make_heap(heap.begin(), heap.end(), std::greater<int>());
        for (const Elem &e : *sub) {
            if (e.state < n) continue;
            const FArc &arc = in_.arcs[e.state - n];
            float ng = e.g, na = e.a;
            walk_weight(arc, 1, &ng, &na);
            relax(arc.next, ng, na, [&] { return walk_string(arc, 1, e.str); });
        }
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), std::greater<int>());
            const int s = heap.back();
            heap.pop_back();
            if (ws_.done[s] == epoch_) continue;
            ws_.done[s] = epoch_;
            const Elem e = cur[ws_.slot[s]];
            for (int k = in_.off[s]; k < in_.off[s + 1]; k++) {
                const FArc &arc = in_.arcs[k];
                if (arc.il != 0) break;  // sorted: no more epsilons
                float ng = e.g, na = e.a;
                walk_weight(arc, 0, &ng, &na);
                relax(arc.next, ng, na, [&] { return walk_string(arc, 0, e.str); });
            }
        }
        sub->clear();
        for (const Elem &e : cur)
            if (ws_.keepable[e.state]) sub->push_back(e);
        std::sort(sub->begin(), sub->end(), [](const Elem &x, const Elem &y) { return x.state < y.state; });
    }

    // NormalizeSubset: take the best weight and the longest common string prefix out of the subset
    void normalize(Elem *e, int n, float *tg, float *ta, int *common) {
        LATP(2);
        if (n == 0) {
            *tg = *ta = kInfF;
            *common = 0;
            return;
        }
        float bg = e[0].g, ba = e[0].a;
        int pre = e[0].str;
        for (int i = 1; i < n; i++) {
            if (compare_w(bg, ba, e[i].g, e[i].a) < 0) {
                bg = e[i].g;
                ba = e[i].a;
            }
            pre = repo_.common_prefix(pre, e[i].str);
        }
        const int plen = repo_.nodes[pre].depth;
        for (int i = 0; i < n; i++) {
            e[i].g -= bg;
            e[i].a -= ba;
            e[i].str = repo_.remove_prefix(e[i].str, plen);
        }
        *tg = bg;
        *ta = ba;
        *common = pre;
    }

    int new_state(const std::vector<Elem> &sub, double fwd) {
        const int id = (int)states_.size();
        const int off = (int)ws_.pool.size();
        ws_.pool.insert(ws_.pool.end(), sub.begin(), sub.end());
        states_.push_back(OutState{off, (int)sub.size(), fwd});
        if (plain_single(sub.data(), (int)sub.size())) ws_.min1[sub[0].state] = id + 1;
        else ws_.min_tab.insert(hash_of(sub.data(), (int)sub.size()), id);
        process_final(id);
        process_transitions(id);
        return id;
    }
    static bool plain_single(const Elem *e, int n) { return n == 1 && e[0].str == 0 && e[0].g == 0.f && e[0].a == 0.f; }
    int minimal_to_state(const std::vector<Elem> &sub, double fwd) {
        const int n = (int)sub.size();
        if (plain_single(sub.data(), n)) {
            int &slot = ws_.min1[sub[0].state];
            if (!slot) slot = new_state(sub, fwd) + 1;
            return slot - 1;
        }
        const int f = ws_.min_tab.find(hash_of(sub.data(), n), [&](int id) {
            return states_[id].len == n && same_subset(&ws_.pool[states_[id].off], sub.data(), n);
        });
        return f >= 0 ? f : new_state(sub, fwd);
    }

    void process_final(int sid) {
        const OutState st = states_[sid];
        bool is_final = false;
        float fg = kInfF, fa = kInfF;
        int fs = 0;
        for (int i = 0; i < st.len; i++) {
            const Elem e = ws_.pool[st.off + i];
            if (in_.fin_g[e.state] == kInfF) continue;
            const float g = e.g + in_.fin_g[e.state], a = e.a + in_.fin_a[e.state];
            if (!is_final || compare_ws(g, a, e.str, fg, fa, fs) == 1) {
                is_final = true;
                fg = g;
                fa = a;
                fs = e.str;
            }
        }
        if (is_final && cost_of(fg, fa) + st.fwd <= cutoff_) ws_.tarcs.push_back(TArc{sid, 0, fs, -1, fg, fa});
    }

    void process_transitions(int sid) {
        LATP(3);
        const OutState st = states_[sid];
        std::vector<Elem> &all = ws_.allv;
        std::vector<int> &lab = ws_.all_label, &idx = ws_.idx;
        std::vector<double> &pri = ws_.all_prio;
        all.clear();
        lab.clear();
        pri.clear();
        for (int i = 0; i < st.len; i++) {
            const Elem e = ws_.pool[st.off + i];
            for (int k = in_.off[e.state + 1] - 1; k >= in_.off[e.state]; k--) {
                const FArc &arc = in_.arcs[k];
                if (arc.il == 0) break;  // sorted: the labelled arcs are at the end
                // the priority is taken where Kaldi takes it: behind the labelled (first) arc of the chain
                const Sym &s0 = sy_[arc.s_off];
                const float ng = e.g + s0.g, na = e.a + s0.a;
                pri.push_back(cost_of(ng, na) + ws_.bw_first[k]);
                // the element behind the labelled arc: at the arc's end, or inside its chain (state id n + arc)
                all.push_back(Elem{arc.s_len == 1 ? arc.next : in_.n + k, s0.tid ? repo_.append(e.str, s0.tid) : e.str, ng, na});
                lab.push_back(arc.il);
            }
        }
        if (all.empty()) return;
        idx.resize(all.size());
        for (size_t i = 0; i < idx.size(); i++) idx[i] = (int)i;
        std::sort(idx.begin(), idx.end(), [&](int x, int y) {
            if (lab[x] != lab[y]) return lab[x] < lab[y];
            if (all[x].state != all[y].state) return all[x].state < all[y].state;
            return x < y;
        });
        size_t c = 0;
        while (c < idx.size()) {
            const int label = lab[idx[c]];
            double prio = kInfD;
            const int off = (int)ws_.pool.size();
            while (c < idx.size() && lab[idx[c]] == label) {
                const Elem &e = all[idx[c]];
                prio = std::min(prio, pri[idx[c]]);
                // MakeSubsetUnique: of several elements with one state the best (weight, then string) stays
                if ((int)ws_.pool.size() > off && ws_.pool.back().state == e.state) {
                    Elem &b = ws_.pool.back();
                    if (compare_ws(e.g, e.a, e.str, b.g, b.a, b.str) == 1) b = e;
                } else {
                    ws_.pool.push_back(e);
                }
                c++;
            }
            prio += st.fwd;
            if (prio > cutoff_) {
                ws_.pool.resize(off);  // beyond the beam: never expanded
                continue;
            }
            tasks_.push(Task{prio, sid, label, off, (int)ws_.pool.size() - off});
        }
    }

    void process_transition(const Task &t) {
        LATP(4);
        double fwd = states_[t.src].fwd;
        Elem *e = &ws_.pool[t.off];
        float tg, ta;
        int common;
        const bool single = t.len == 1;
        if (single) {  // NormalizeSubset of one element: its weight and its whole string come out
            tg = e[0].g;
            ta = e[0].a;
            common = e[0].str;
            e[0].g = 0.f;
            e[0].a = 0.f;
            e[0].str = 0;
        } else {
            normalize(e, t.len, &tg, &ta, &common);
        }
        fwd += cost_of(tg, ta);
        // InitialToStateId
        int next, nstr;
        float ng, na;
        uint64_t h = 0;
        int f = -1;
        if (single) {
            f = ws_.init1[e[0].state] - 1;
        } else {
            h = hash_of(e, t.len);
            f = ws_.init_tab.find(h, [&](int id) { return inits_[id].len == t.len && same_subset(&ws_.pool[inits_[id].off], e, t.len); });
        }
        if (f >= 0) {
            next = inits_[f].state;
            nstr = inits_[f].str;
            ng = inits_[f].g;
            na = inits_[f].a;
        } else {
            std::vector<Elem> &sub = ws_.sub;
            const int key_state = e[0].state;
            sub.assign(e, e + t.len);
            closure(&sub);
            if (sub.size() == 1) {
                ng = sub[0].g;
                na = sub[0].a;
                nstr = sub[0].str;
                sub[0].g = 0.f;
                sub[0].a = 0.f;
                sub[0].str = 0;
            } else {
                normalize(sub.data(), (int)sub.size(), &ng, &na, &nstr);
            }
            next = minimal_to_state(sub, fwd + cost_of(ng, na));
            inits_.push_back(Init{t.off, t.len, next, nstr, ng, na});
            if (single) ws_.init1[key_state] = (int)inits_.size();
            else ws_.init_tab.insert(h, (int)inits_.size() - 1);
        }
        ws_.tarcs.push_back(TArc{t.src, t.label, repo_.concat(common, nstr), next, tg + ng, ta + na});
    }

    Workspace &ws_;
    const FLat &in_;
    const Sym *sy_;
    double beam_, cutoff_ = 0;
    Repo &repo_;
    int epoch_ = 0;
    std::vector<OutState> states_;
    std::vector<Init> inits_;
    std::priority_queue<Task> tasks_;
};

// output of the phone pass in non-compact form with the phone labels deleted: Kaldi spells every string out as symbols
// along a chain of new states, weight and label on the first arc of the chain — here one arc with that chain as its span
void output_pass1(Workspace &ws, int n_det, std::vector<Edge> *edges, std::vector<Sym> *syms, std::vector<float> *fin_g, std::vector<float> *fin_a,
                  int *n_out) {
    LATP(8);
    edges->clear();
    syms->clear();
    fin_g->assign(n_det, kInfF);
    fin_a->assign(n_det, kInfF);
    int n = n_det;
    std::vector<int> &seq = ws.tvec;
    for (const TArc &t : ws.tarcs) {
        seq.clear();
        ws.repo.to_vec(t.str, &seq);
        if (t.next < 0 && seq.empty()) {
            (*fin_g)[t.src] = t.g;
            (*fin_a)[t.src] = t.a;
            continue;
        }
        const int off = (int)syms->size();
        if (seq.empty()) syms->push_back(Sym{0, t.g, t.a});
        for (size_t i = 0; i < seq.size(); i++) syms->push_back(Sym{seq[i], i == 0 ? t.g : 0.f, i == 0 ? t.a : 0.f});
        int next = t.next;
        if (next < 0) {  // a final weight with a string: the chain ends in a new state whose final weight is One
            next = n++;
            fin_g->push_back(0.f);
            fin_a->push_back(0.f);
        }
        const int label = t.next < 0 || t.label >= kPhoneBase ? 0 : t.label;  // DeterminizeLatticeDeletePhones
        edges->push_back(Edge{t.src, FArc{label, off, (int)syms->size() - off, next}});
    }
    *n_out = n;
}

// output of the word pass in compact form + fst::Connect (drop the states from which no final weight can be reached)
void compact_output(Workspace &ws, int n_det, CLatF *out) {
    LATP(9);
    out->tids.clear();
    std::vector<int> &off = ws.csr_off;
    off.assign(n_det + 1, 0);
    std::vector<char> fin(n_det, 0);
    for (const TArc &t : ws.tarcs)
        if (t.next >= 0) off[t.src + 1]++;
        else fin[t.src] = 1;
    for (int s = 0; s < n_det; s++) off[s + 1] += off[s];
    std::vector<int> &idx = ws.csr_idx, &pos = ws.dfs_pos;
    idx.resize(off[n_det]);
    pos.assign(off.begin(), off.end() - 1);
    for (size_t k = 0; k < ws.tarcs.size(); k++)
        if (ws.tarcs[k].next >= 0) idx[pos[ws.tarcs[k].src]++] = (int)k;
    // co-accessibility: reverse reachability from the final states
    std::vector<char> &co = ws.color;
    co.assign(n_det, 0);
    {
        std::vector<int> roff(n_det + 1, 0), ridx(off[n_det]);
        for (const TArc &t : ws.tarcs)
            if (t.next >= 0) roff[t.next + 1]++;
        for (int s = 0; s < n_det; s++) roff[s + 1] += roff[s];
        std::vector<int> rp(roff.begin(), roff.end() - 1);
        for (const TArc &t : ws.tarcs)
            if (t.next >= 0) ridx[rp[t.next]++] = t.src;
        std::vector<int> &st = ws.dfs_stack;
        st.clear();
        for (int s = 0; s < n_det; s++)
            if (fin[s]) {
                co[s] = 1;
                st.push_back(s);
            }
        while (!st.empty()) {
            const int s = st.back();
            st.pop_back();
            for (int k = roff[s]; k < roff[s + 1]; k++)
                if (!co[ridx[k]]) {
                    co[ridx[k]] = 1;
                    st.push_back(ridx[k]);
                }
        }
    }
    // (every output state is accessible from state 0 by construction)
    std::vector<int> &map = ws.remap;
    map.assign(n_det, -1);
    int n = 0;
    for (int s = 0; s < n_det; s++)
        if (co[s]) map[s] = n++;
    out->n = n;
    out->start = n_det > 0 ? map[0] : -1;
    out->off.assign(n + 1, 0);
    out->arcs.clear();
    out->is_final.assign(n, 0);
    out->fin_g.assign(n, 0.f);
    out->fin_a.assign(n, 0.f);
    out->fin_off.assign(n, 0);
    out->fin_len.assign(n, 0);
    for (int s = 0; s < n_det; s++) {
        if (map[s] < 0) continue;
        for (int k = off[s]; k < off[s + 1]; k++) {
            const TArc &t = ws.tarcs[idx[k]];
            if (map[t.next] < 0) continue;
            const int o = (int)out->tids.size();
            ws.repo.to_vec(t.str, &out->tids);
            out->arcs.push_back(CArcF{t.label, t.g, t.a, o, (int)out->tids.size() - o, map[t.next]});
        }
        out->off[map[s] + 1] = (int)out->arcs.size();
    }
    for (const TArc &t : ws.tarcs)
        if (t.next < 0 && map[t.src] >= 0) {
            const int s = map[t.src], o = (int)out->tids.size();
            ws.repo.to_vec(t.str, &out->tids);
            out->is_final[s] = 1;
            out->fin_g[s] = t.g;
            out->fin_a[s] = t.a;
            out->fin_off[s] = o;
            out->fin_len[s] = (int)out->tids.size() - o;
        }
}

// DeterminizeLatticePhonePrunedWrapper
bool determinize_phone_pruned(Workspace &ws, const RawLattice &raw, const Model &m, double beam, bool phone_pass, CLatF *out,
                              LatticeStats *stats) {
    LATP(0);
    const Graph &g = m.graph;
    const int n_raw = raw.n_states;
    if (n_raw <= 0 || raw.start < 0 || raw.start >= n_raw) return false;
    // Invert (words become the labels determinized on) + DeterminizeLatticeInsertPhones
    std::vector<Edge> &edges = ws.edges;
    std::vector<Sym> &syms = ws.sym_in;
    edges.clear();
    syms.clear();
    ws.fin_g.assign(n_raw, kInfF);
    ws.fin_a.assign(n_raw, kInfF);
    int n = n_raw;
    const int n_tids = (int)m.tid_flags.size();
    for (size_t k = 0; k < raw.src.size(); k++) {
        const int a = raw.arc[k], s = raw.src[k], d = raw.dst[k];
        if (s < 0 || s >= n_raw || d < 0 || d >= n_raw || a < 0 || a >= g.num_arcs) continue;
        const int tid = g.arc_ilabel[a], word = g.arc_olabel[a];
        FArc arc{word, (int)syms.size(), 1, d};
        syms.push_back(Sym{tid, g.arc_w[a], raw.acoustic[k]});
        if (phone_pass && s != raw.start && tid > 0 && tid < n_tids && (m.tid_flags[tid] & 4) && !(m.tid_flags[tid] & 1)) {
            const int phone_label = kPhoneBase + m.tid2phone[tid];
            if (word == 0) {
                arc.il = phone_label;
            } else {  // a word and a phone start on one arc: the phone label goes on an extra arc behind it
                const int extra = n++;
                ws.fin_g.push_back(kInfF);
                ws.fin_a.push_back(kInfF);
                arc.next = extra;
                edges.push_back(Edge{s, arc});
                edges.push_back(Edge{extra, FArc{phone_label, (int)syms.size(), 1, d}});
                syms.push_back(Sym{0, 0.f, 0.f});
                continue;
            }
        }
        edges.push_back(Edge{s, arc});
    }
    for (size_t k = 0; k < raw.final_state.size(); k++)
        if (raw.final_state[k] >= 0 && raw.final_state[k] < n_raw) {
            ws.fin_g[raw.final_state[k]] = raw.final_cost[k];
            ws.fin_a[raw.final_state[k]] = 0.f;
        }
    g_comp_pass = 1;
    compress_chains(ws, n, raw.start, edges, ws.fin_g, syms, &ws.edges2, &ws.sym1);
    if (!build_flat(ws, n, raw.start, ws.edges2, ws.fin_g, ws.fin_a, &ws.sym1, &ws.lat1)) return false;
    const FLat *in = &ws.lat1;
    if (phone_pass) {
        Determinizer d1(ws, ws.lat1, beam);
        if (!d1.run()) return false;
        int n2 = 0;
        output_pass1(ws, d1.num_states(), &edges, &syms, &ws.fin_g, &ws.fin_a, &n2);
        if (stats) {
            stats->det1_states = d1.num_states();
            stats->det1_arcs = (int)edges.size();
        }
        g_comp_pass = 2;
        compress_chains(ws, n2, 0, edges, ws.fin_g, syms, &ws.edges2, &ws.sym2);
        if (!build_flat(ws, n2, 0, ws.edges2, ws.fin_g, ws.fin_a, &ws.sym2, &ws.lat2)) return false;
        in = &ws.lat2;
    }
    Determinizer d2(ws, *in, beam);
    if (!d2.run()) return false;
    compact_output(ws, d2.num_states(), out);
    if (stats) {
        stats->det_states = out->n;
        stats->det_arcs = (int)out->arcs.size();
    }
    return out->n > 0 && out->start >= 0;
}

// fst::ScaleLattice(fst::GraphLatticeScale(s)): ScaleTupleWeight with the double matrix {{s, 0}, {0, 1}}
inline void scale_weight(double s, float *g, float *a) {
    if (*g == kInfF) return;
    const double dg = (double)*g, da = (double)*a;
    *g = (float)(s * dg + 0.0 * da);
    *a = (float)(0.0 * dg + 1.0 * da);
}
void scale_graph_costs(CLatF *lat, double s) {
    for (CArcF &a : lat->arcs) scale_weight(s, &a.g, &a.a);
    for (int i = 0; i < lat->n; i++)
        if (lat->is_final[i]) scale_weight(s, &lat->fin_g[i], &lat->fin_a[i]);
}

// ------------------------------------------------------------------------------------------------------------
// WordAlignLattice
// ------------------------------------------------------------------------------------------------------------
class WordAligner {
   public:
    WordAligner(Workspace &ws, const CLatF &in, const Model &m) : ws_(ws), in_(in), m_(m), out_(ws.ali) {}

    void run() {
        out_.clear();
        ws_.seq.clear();
        ws_.wseq.clear();
        tuples_.clear();
        ws_.tup_tab.reset(1 << 12);
        queue_.clear();
        if (in_.start < 0) return;
        // CreateSuperFinal: the only final weight left is One on one extra state `sf_` (arcs into it are implicit)
        sf_ = in_.n;
        const int n_final = (int)std::count(in_.is_final.begin(), in_.is_final.end(), (char)1);
        if (n_final == 1) {
            for (int s = 0; s < in_.n; s++)
                if (in_.is_final[s] && in_.fin_g[s] == 0.f && in_.fin_a[s] == 0.f && in_.fin_len[s] == 0 && in_.off[s] == in_.off[s + 1]) sf_ = s;
        }
        out_.start = state_for(Tuple{in_.start, 0, 0, 0, 0});
        while (!queue_.empty()) {
            const int id = queue_.back();
            queue_.pop_back();
            process(id);
        }
        remove_eps_local();
        for (AArc &a : out_.arcs)
            if (a.label < 0) a.label = 0;
    }

   private:
    struct Tuple {
        int in_state;
        int t_off, t_len;  // pending transition-ids (span in ws.seq)
        int w_off, w_len;  // pending word labels (span in ws.wseq)
    };
    int phone_of(int tid) const { return tid >= 0 && tid < (int)m_.tid2phone.size() ? m_.tid2phone[tid] : -1; }
    int type_of(int phone) const { return phone >= 0 && phone < (int)m_.phone_type.size() ? m_.phone_type[phone] : 0; }
    bool is_final_tid(int tid) const { return tid > 0 && tid < (int)m_.tid_flags.size() && (m_.tid_flags[tid] & 2); }
    bool is_self_loop(int tid) const { return tid > 0 && tid < (int)m_.tid_flags.size() && (m_.tid_flags[tid] & 1); }

    uint64_t hash_tuple(const Tuple &t) const {
        uint64_t h = mix((uint64_t)(uint32_t)t.in_state, ((uint64_t)(uint32_t)t.t_len << 32) | (uint32_t)t.w_len);
        for (int i = 0; i < t.t_len; i++) h = mix(h, (uint64_t)(uint32_t)ws_.seq[t.t_off + i]);
        for (int i = 0; i < t.w_len; i++) h = mix(h, (uint64_t)(uint32_t)ws_.wseq[t.w_off + i] + 0x51ull);
        return h;
    }
    int state_for(const Tuple &t) {
        const uint64_t h = hash_tuple(t);
        const int f = ws_.tup_tab.find(h, [&](int id) {
            const Tuple &o = tuples_[id];
            return o.in_state == t.in_state && o.t_len == t.t_len && o.w_len == t.w_len &&
                   std::equal(ws_.seq.begin() + t.t_off, ws_.seq.begin() + t.t_off + t.t_len, ws_.seq.begin() + o.t_off) &&
                   std::equal(ws_.wseq.begin() + t.w_off, ws_.wseq.begin() + t.w_off + t.w_len, ws_.wseq.begin() + o.w_off);
        });
        if (f >= 0) return f;
        const int id = out_.add_state();
        tuples_.push_back(t);
        ws_.tup_tab.insert(h, id);
        queue_.push_back(id);
        return id;
    }

    // number of leading tids that make up one complete phone, or 0 if its end cannot be decided yet (reorder = true:
    // the self-loops follow the final transition, so a phone is only known to be over when the next tid is in sight)
    int phone_span(const int *t, int len, int from) const {
        int i = from;
        for (; i < len; i++)
            if (is_final_tid(t[i])) break;
        if (i == len) return 0;
        i++;
        while (i < len && is_self_loop(t[i])) i++;
        if (i == len) return 0;
        return i;
    }
    // OutputSilenceArc / OutputOnePhoneWordArc / OutputNormalWordArc: tids consumed (0 = nothing to output yet)
    int output_arc(const Tuple &t, int *label, bool *takes_word) const {
        if (t.t_len == 0) return 0;
        const int *tids = &ws_.seq[t.t_off];
        const int ty = type_of(phone_of(tids[0]));
        if (ty == 1) {
            *label = kSilTmp;
            *takes_word = false;
            return phone_span(tids, t.t_len, 0);
        }
        if (t.w_len == 0) return 0;
        *label = ws_.wseq[t.w_off];
        *takes_word = true;
        if (ty == 5) return phone_span(tids, t.t_len, 0);
        if (ty == 2) {  // begin phone, word-internal phones, end phone
            int i = phone_span(tids, t.t_len, 0);
            if (!i) return 0;
            while (i < t.t_len) {
                const int t2 = type_of(phone_of(tids[i]));
                if (t2 == 3 || t2 == 5) break;
                i++;
            }
            if (i == t.t_len) return 0;
            return phone_span(tids, t.t_len, i);
        }
        return 0;
    }

    void process(int id) {
        const Tuple t = tuples_[id];
        int label = 0;
        bool takes_word = false;
        const int used = output_arc(t, &label, &takes_word);
        if (used > 0) {
            const int w = takes_word ? 1 : 0;
            const int d = state_for(Tuple{t.in_state, t.t_off + used, t.t_len - used, t.w_off + w, t.w_len - w});
            out_.add_arc(id, AArc{label, 0.f, 0.f, t.t_off, used, d, -1});
            return;
        }
        if (t.in_state == sf_) {  // ProcessFinal
            if (t.t_len == 0 && t.w_len == 0) {
                out_.is_final[id] = 1;
                out_.fin_g[id] = 0.f;
                out_.fin_a[id] = 0.f;
            } else {  // OutputArcForce: whatever is pending goes out as one arc (silence, word, or partial word)
                int lab;
                int w = 0;
                if (t.t_len > 0) {
                    const int ty = type_of(phone_of(ws_.seq[t.t_off]));
                    if (ty == 1) {
                        lab = kSilTmp;
                    } else if (t.w_len > 0) {
                        lab = ws_.wseq[t.w_off];
                        w = 1;
                    } else {
                        lab = kPartialTmp;
                    }
                } else {
                    lab = ws_.wseq[t.w_off];
                    w = 1;
                }
                const int d = state_for(Tuple{t.in_state, t.t_off + t.t_len, 0, t.w_off + w, t.w_len - w});
                out_.add_arc(id, AArc{lab, 0.f, 0.f, t.t_off, t.t_len, d, -1});
            }
        }
        if (t.in_state == sf_ && sf_ == in_.n) return;  // the implicit super-final state has no arcs
        // Advance: consume one input arc; its symbols are queued, its weight goes out on an epsilon arc
        auto advance = [&](int label_in, float g, float a, const int *str, int len, int next) {
            Tuple nx{next, t.t_off, t.t_len, t.w_off, t.w_len};
            if (len > 0) {
                if (t.t_off + t.t_len != (int)ws_.seq.size()) {
                    nx.t_off = (int)ws_.seq.size();
                    ws_.seq.resize(ws_.seq.size() + t.t_len);
                    std::copy_n(ws_.seq.begin() + t.t_off, t.t_len, ws_.seq.begin() + nx.t_off);
                }
                ws_.seq.insert(ws_.seq.end(), str, str + len);
                nx.t_len = t.t_len + len;
            }
            if (label_in != 0) {
                if (t.w_off + t.w_len != (int)ws_.wseq.size()) {
                    nx.w_off = (int)ws_.wseq.size();
                    ws_.wseq.resize(ws_.wseq.size() + t.w_len);
                    std::copy_n(ws_.wseq.begin() + t.w_off, t.w_len, ws_.wseq.begin() + nx.w_off);
                }
                ws_.wseq.push_back(label_in);
                nx.w_len = t.w_len + 1;
            }
            const int d = state_for(nx);
            out_.add_arc(id, AArc{0, g, a, 0, 0, d, -1});
        };
        if (t.in_state >= in_.n) return;
        // (the strings are copied out of in_.tids because ws_.seq may reallocate while it is appended to)
        for (int k = in_.off[t.in_state]; k < in_.off[t.in_state + 1]; k++) {
            const CArcF &a = in_.arcs[k];
            advance(a.label, a.g, a.a, in_.tids.data() + a.t_off, a.t_len, a.next);
        }
        if (in_.is_final[t.in_state] && t.in_state != sf_)
            advance(0, in_.fin_g[t.in_state], in_.fin_a[t.in_state], in_.tids.data() + in_.fin_off[t.in_state], in_.fin_len[t.in_state], sf_);
    }

    // RemoveEpsLocal: merge an arc with the arc(s) leaving its destination when one of the two is an epsilon and the
    // destination has a single arc in (pattern 1) or a single arc out (pattern 2); never increases the arc count.
    bool can_combine(const AArc &a, const AArc &b, AArc *c) {
        if (a.label != 0 && b.label != 0) return false;
        c->label = a.label != 0 ? a.label : b.label;
        c->g = a.g + b.g;
        c->a = a.a + b.a;
        if (a.t_len == 0) {
            c->t_off = b.t_off;
            c->t_len = b.t_len;
        } else if (b.t_len == 0) {
            c->t_off = a.t_off;
            c->t_len = a.t_len;
        } else {
            c->t_off = (int)ws_.seq.size();
            c->t_len = a.t_len + b.t_len;
            ws_.seq.resize(ws_.seq.size() + c->t_len);
            std::copy_n(ws_.seq.begin() + a.t_off, a.t_len, ws_.seq.begin() + c->t_off);
            std::copy_n(ws_.seq.begin() + b.t_off, b.t_len, ws_.seq.begin() + c->t_off + a.t_len);
        }
        c->next = b.next;
        c->link = -1;
        return true;
    }
    void set_final_plus(int s, float g, float a) {  // SetFinal(s, Plus(Final(s), w)); final strings are empty here
        ALat &f = out_;
        if (!f.is_final[s]) {
            ws_.nout[s]++;
            f.is_final[s] = 1;
            f.fin_g[s] = g;
            f.fin_a[s] = a;
        } else if (compare_w(g, a, f.fin_g[s], f.fin_a[s]) == 1) {
            f.fin_g[s] = g;
            f.fin_a[s] = a;
        }
    }
    void remove_eps_local() {
        ALat &f = out_;
        const int n = f.n();
        std::vector<int> &nin = ws_.nin, &nout = ws_.nout;
        nin.assign(n, 0);
        nout.assign(n, 0);
        for (int s = 0; s < n; s++) {
            for (int k = f.head[s]; k >= 0; k = f.arcs[k].link) {
                nin[f.arcs[k].next]++;
                nout[s]++;
            }
            if (f.is_final[s]) nout[s]++;
        }
        if (f.start >= 0) nin[f.start]++;
        std::vector<AArc> add;
        for (int s = 0; s < n; s++) {
            for (int pos = f.head[s]; pos >= 0; pos = f.arcs[pos].link) {
                const AArc arc = f.arcs[pos];
                const int t = arc.next;
                if (t == kDead || t == s) continue;
                if (nin[t] == 1 && nout[t] > 1) {  // pattern 1
                    bool removed = false, kept = false;
                    add.clear();
                    for (int k = f.head[t]; k >= 0; k = f.arcs[k].link) {
                        AArc &nx = f.arcs[k];
                        if (nx.next == kDead) continue;
                        AArc c;
                        if (can_combine(arc, nx, &c)) {
                            removed = true;
                            nout[t]--;
                            nin[f.arcs[k].next]--;
                            f.arcs[k].next = kDead;
                            add.push_back(c);
                        } else {
                            kept = true;
                        }
                    }
                    if (f.is_final[t]) {
                        if (arc.label == 0) {  // CanCombineFinal: an epsilon arc into a final state
                            removed = true;
                            set_final_plus(s, arc.g + f.fin_g[t], arc.a + f.fin_a[t]);
                            nout[t]--;
                            f.is_final[t] = 0;
                        } else {
                            kept = true;
                        }
                    }
                    if (removed && !kept) {
                        nout[s]--;
                        nin[t]--;
                        f.arcs[pos].next = kDead;
                    }
                    for (const AArc &c : add) {
                        nout[s]++;
                        nin[c.next]++;
                        f.add_arc(s, c);
                    }
                } else if (nout[t] == 1) {  // pattern 2
                    bool del = false;
                    if (f.is_final[t]) {
                        if (arc.label == 0) {
                            set_final_plus(s, arc.g + f.fin_g[t], arc.a + f.fin_a[t]);
                            del = true;
                        }
                    } else {
                        int k = f.head[t];
                        while (k >= 0 && f.arcs[k].next == kDead) k = f.arcs[k].link;
                        AArc c;
                        if (k >= 0 && can_combine(arc, f.arcs[k], &c)) {
                            del = true;
                            if (nin[t] == 1) {
                                nout[t]--;
                                nin[f.arcs[k].next]--;
                                f.arcs[k].next = kDead;
                            }
                            nout[s]++;
                            nin[c.next]++;
                            f.add_arc(s, c);
                        }
                    }
                    if (del) {
                        nout[s]--;
                        nin[t]--;
                        f.arcs[pos].next = kDead;
                    }
                }
            }
        }
    }

    Workspace &ws_;
    const CLatF &in_;
    const Model &m_;
    ALat &out_;
    int sf_ = -1;
    std::vector<Tuple> tuples_;
    std::vector<int> queue_;
};

// ------------------------------------------------------------------------------------------------------------
// MinimumBayesRisk (Xu, Povey, Mangu, Zhu: "Minimum Bayes Risk decoding and system combination based on a
// recursion for edit distance", as implemented by Kaldi lat/sausages.cc)
// ------------------------------------------------------------------------------------------------------------
inline double log_add(double a, double b) {
    if (a == -kInfD) return b;
    if (b == -kInfD) return a;
    const double hi = std::max(a, b), diff = std::min(a, b) - hi;
    return hi + std::log1p(std::exp(diff));
}

// PrepareLatticeAndInitStats: the word-aligned lattice as MinimumBayesRisk sees it (false: no complete path)
bool mbr_prepare(Workspace &ws, MbrJob *job, LatticeStats *stats) {
    const bool result = false;
    job->N = 0;
    const ALat &lat = ws.ali;
    const int n_all = lat.n();
    if (lat.start < 0 || n_all == 0) return result;
    // live arcs in CSR form; Connect (states reachable from the start); CreateSuperFinal; topological numbering
    std::vector<MEdge> &edges = ws.medges;
    edges.clear();
    int sf = n_all, n_final = 0, last_final = -1;
    for (int s = 0; s < n_all; s++)
        if (lat.is_final[s]) {
            n_final++;
            last_final = s;
        }
    if (!n_final) return result;
    if (n_final == 1 && lat.fin_g[last_final] == 0.f && lat.fin_a[last_final] == 0.f) {
        // CreateSuperFinal: a single final state with weight One and no arcs out already is the super-final state
        bool live = false;
        for (int k = lat.head[last_final]; k >= 0; k = lat.arcs[k].link) live = live || lat.arcs[k].next != kDead;
        if (!live) sf = last_final;
    }
    for (int s = 0; s < n_all; s++) {
        for (int k = lat.head[s]; k >= 0; k = lat.arcs[k].link) {
            const AArc &a = lat.arcs[k];
            if (a.next == kDead) continue;
            edges.push_back(MEdge{s, a.label, a.t_len, a.next, -(a.g + a.a)});
        }
        if (lat.is_final[s] && s != sf) edges.push_back(MEdge{s, 0, 0, sf, -(lat.fin_g[s] + lat.fin_a[s])});
    }
    std::vector<int> &off = ws.csr_off, &idx = ws.csr_idx, &order = ws.order, &pos = ws.dfs_pos, &st = ws.dfs_stack, &finish = ws.remap;
    const int n_tot = n_all + 1;
    off.assign(n_tot + 1, 0);
    for (const MEdge &e : edges) off[e.src + 1]++;
    for (int s = 0; s < n_tot; s++) off[s + 1] += off[s];
    idx.resize(edges.size());
    pos.assign(off.begin(), off.end() - 1);
    for (size_t k = 0; k < edges.size(); k++) idx[pos[edges[k].src]++] = (int)k;
    order.assign(n_tot, -1);
    ws.color.assign(n_tot, 0);
    pos.assign(off.begin(), off.end() - 1);
    st.clear();
    finish.clear();
    st.push_back(lat.start);
    ws.color[lat.start] = 1;
    while (!st.empty()) {
        const int s = st.back();
        if (pos[s] < off[s + 1]) {
            const int d = edges[idx[pos[s]++]].next;
            if (!ws.color[d]) {
                ws.color[d] = 1;
                st.push_back(d);
            }
        } else {
            finish.push_back(s);
            st.pop_back();
        }
    }
    if (!ws.color[sf] || finish[0] != sf) {
        // the super-final state has no arcs out: the first DFS branch ends there unless it dead-ends in a trimmed-away state
        auto it = std::find(finish.begin(), finish.end(), sf);
        if (it == finish.end()) return result;
        finish.erase(it);
        finish.insert(finish.begin(), sf);
    }
    const int N = (int)finish.size();
    for (int i = 0; i < N; i++) order[finish[N - 1 - i]] = i + 1;  // 1-based nodes, super-final = N
    // arcs grouped by end node, in start-node order (pre_[n] of sausages.cc)
    std::vector<MbrArc> arcs;
    std::vector<int> pre_off(N + 2, 0);
    {
        std::vector<MbrArc> tmp;
        for (int i = N - 1; i >= 0; i--) {
            const int s = finish[i];  // node N - i
            for (int k = off[s]; k < off[s + 1]; k++) {
                const MEdge &e = edges[idx[k]];
                if (order[e.next] < 0) continue;
                tmp.push_back(MbrArc{e.word, order[s], order[e.next], e.loglike, e.len});
            }
        }
        for (const MbrArc &a : tmp) pre_off[a.end + 1]++;
        for (int n = 0; n <= N; n++) pre_off[n + 1] += pre_off[n];
        arcs.resize(tmp.size());
        std::vector<int> p(pre_off.begin(), pre_off.end() - 1);
        for (const MbrArc &a : tmp) arcs[p[a.end]++] = a;
    }
    job->N = N;
    job->arcs.swap(arcs);
    job->pre_off.swap(pre_off);
    const std::vector<MbrArc> &A = job->arcs;
    const std::vector<int> &P = job->pre_off;
    job->state_times.assign(N + 1, 0);
    for (int n = 2; n <= N; n++)
        for (int k = P[n]; k < P[n + 1]; k++) job->state_times[n] = job->state_times[A[k].start] + A[k].t_len;
    // initial R: words of the best path
    job->R0.clear();
    {
        std::vector<double> best(N + 1, kInfD);
        std::vector<int> back(N + 1, -1);
        best[1] = 0;
        for (int n = 2; n <= N; n++)
            for (int k = P[n]; k < P[n + 1]; k++) {
                const double c = best[A[k].start] - (double)A[k].loglike;
                if (c < best[n]) {
                    best[n] = c;
                    back[n] = k;
                }
            }
        for (int n = N; n > 1 && back[n] >= 0; n = A[back[n]].start)
            if (A[back[n]].word != 0) job->R0.push_back(A[back[n]].word);
        std::reverse(job->R0.begin(), job->R0.end());
    }
    // forward log-probabilities and arc posteriors given their end node: the only transcendental part of MinimumBayesRisk
    // (they do not depend on the hypothesis, so they are computed once; the edit-distance recursion itself is plain arithmetic)
    std::vector<double> &alpha = ws.alpha;
    alpha.assign(N + 1, 0.0);
    job->post.resize(A.size());
    for (int n = 2; n <= N; n++) {
        double alpha_n = -kInfD;
        for (int k = P[n]; k < P[n + 1]; k++) alpha_n = log_add(alpha_n, alpha[A[k].start] + A[k].loglike);
        alpha[n] = alpha_n;
        for (int k = P[n]; k < P[n + 1]; k++) job->post[k] = std::exp(alpha[A[k].start] + A[k].loglike - alpha_n);
    }
    if (stats) {
        stats->ali_states = N;
        stats->ali_arcs = (int)A.size();
    }
    return true;
}

}  // namespace

// the edit-distance recursion and the decision loop (vb_mbr.h) on the host, with buffers that grow on demand
std::vector<WordSpan> mbr_solve_host(const MbrJob &job, LatticeStats *stats) {
    std::vector<WordSpan> result;
    if (job.N <= 0) return result;
    struct Buf {
        std::vector<double> ad, bd, cq, m12a, m12b, vala, valb;
        std::vector<char> b_all;
        std::vector<MbrAcc> acc;
        std::vector<int> acc_n, R, R2, words, rh, nz_lo, nz_hi;
        std::vector<float> tb, te, conf, otb, ote, oconf;
    };
    static thread_local Buf b;
    int w_cap = 4 * (int)job.R0.size() + 16, acc_cap = 64;
    for (int attempt = 0; attempt < 8; attempt++) {
        const size_t nw = (size_t)(job.N + 1) * w_cap;
        if (b.ad.size() < nw) {
            b.ad.resize(nw);
            b.bd.resize(nw);
        }
        if (b.b_all.size() < job.arcs.size() * (size_t)w_cap) b.b_all.resize(job.arcs.size() * (size_t)w_cap);
        for (std::vector<double> *v : {&b.cq, &b.m12a, &b.m12b, &b.vala, &b.valb})
            if ((int)v->size() < w_cap) v->resize(w_cap);
        if (b.acc.size() < (size_t)w_cap * acc_cap) b.acc.resize((size_t)w_cap * acc_cap);
        for (std::vector<int> *v : {&b.acc_n, &b.R, &b.R2, &b.words, &b.rh})
            if ((int)v->size() < w_cap) v->resize(w_cap);
        if ((int)b.nz_lo.size() < job.N + 1) {
            b.nz_lo.resize(job.N + 1);
            b.nz_hi.resize(job.N + 1);
        }
        for (std::vector<float> *v : {&b.tb, &b.te, &b.conf, &b.otb, &b.ote, &b.oconf})
            if ((int)v->size() < w_cap) v->resize(w_cap);
        MbrView view{job.N, (int)job.arcs.size(), job.arcs.data(), job.pre_off.data(), job.state_times.data(), job.post.data()};
        MbrScratch sc{w_cap, acc_cap, b.ad.data(), b.bd.data(), b.b_all.data(), b.cq.data(), b.m12a.data(), b.m12b.data(), b.vala.data(), b.valb.data(),
                      b.rh.data(), b.nz_lo.data(), b.nz_hi.data(), b.acc.data(), b.acc_n.data(), b.R.data(), b.R2.data(), b.tb.data(), b.te.data(), b.conf.data()};
        MbrOut out{};
        mbr_solve(view, sc, job.R0.data(), (int)job.R0.size(), &out, b.words.data(), b.otb.data(), b.ote.data(), b.oconf.data());
        if (out.status == 1) {
            w_cap *= 2;
            continue;
        }
        if (out.status == 2) {
            acc_cap *= 4;
            continue;
        }
        if (stats) {
            stats->mbr_iters = out.iters;
            stats->mbr_q = out.q;
        }
        for (int i = 0; i < out.n_words; i++) result.push_back(WordSpan{b.words[i], b.otb[i], b.ote[i], b.oconf[i]});
        return result;
    }
    return result;
}

namespace {
inline double ms_since(std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
}  // namespace

bool lattice_to_mbr_job(const RawLattice &raw, const Model &m, float lattice_beam, double lm_scale, MbrJob *job, LatticeStats *stats) {
#ifdef VB_LAT_PROF
    struct Dump { ~Dump() { if (getenv("VB_LAT_PROF_DUMP")) { for (int k = 0; k < 12; k++) fprintf(stderr, "prof[%d]=%.3f Mcyc\n", k, g_prof[k] / 1e6); } } };
    static thread_local Dump dump;
#endif
    Workspace &ws = workspace();
    job->N = 0;
    auto t0 = std::chrono::steady_clock::now();
    if (stats) {
        stats->raw_states = raw.n_states;
        stats->raw_arcs = (int)raw.src.size();
    }
    if (!determinize_phone_pruned(ws, raw, m, (double)lattice_beam, true, &ws.clat, stats)) return false;
    scale_graph_costs(&ws.clat, lm_scale);
    if (stats) stats->ms_det = ms_since(t0);
    t0 = std::chrono::steady_clock::now();
    WordAligner(ws, ws.clat, m).run();
    mbr_prepare(ws, job, stats);  // (N = 0: a lattice without a complete aligned path — MinimumBayesRisk then has no words)
    if (stats) stats->ms_align = ms_since(t0);
    return true;
}

std::vector<WordSpan> lattice_to_words(const RawLattice &raw, const Model &m, float lattice_beam, double lm_scale, LatticeStats *stats, bool *ok) {
    if (ok) *ok = false;
    static thread_local MbrJob job;
    if (!lattice_to_mbr_job(raw, m, lattice_beam, lm_scale, &job, stats)) return {};
    if (ok) *ok = true;  // the chain ran; an empty result now means a lattice without words (silence), not a failure
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<WordSpan> r = mbr_solve_host(job, stats);
    if (stats) stats->ms_mbr = ms_since(t0);
    return r;
}

std::string lattice_debug_text(const RawLattice &raw, const Model &m, float lattice_beam, int stage, bool phone_pass, double lm_scale) {
    Workspace &ws = workspace();
    std::string text;
    char b[160];
    auto tids = [](const int *t, int n) {
        std::string s;
        for (int i = 0; i < n; i++) s += (i ? "," : "") + std::to_string(t[i]);
        return s.empty() ? std::string("-") : s;
    };
    if (!determinize_phone_pruned(ws, raw, m, (double)lattice_beam, phone_pass, &ws.clat, nullptr)) return "S -1\n";
    scale_graph_costs(&ws.clat, lm_scale);
    if (stage == 1) {
        const CLatF &L = ws.clat;
        snprintf(b, sizeof b, "S %d\n", L.start);
        text += b;
        for (int s = 0; s < L.n; s++) {
            for (int k = L.off[s]; k < L.off[s + 1]; k++) {
                const CArcF &a = L.arcs[k];
                snprintf(b, sizeof b, "A %d %d %d %.9g %.9g ", s, a.next, a.label, a.g, a.a);
                text += b + tids(L.tids.data() + a.t_off, a.t_len) + "\n";
            }
            if (L.is_final[s]) {
                snprintf(b, sizeof b, "F %d %.9g %.9g ", s, L.fin_g[s], L.fin_a[s]);
                text += b + tids(L.tids.data() + L.fin_off[s], L.fin_len[s]) + "\n";
            }
        }
        return text;
    }
    WordAligner(ws, ws.clat, m).run();
    const ALat &L = ws.ali;
    snprintf(b, sizeof b, "S %d\n", L.start);
    text += b;
    for (int s = 0; s < L.n(); s++) {
        for (int k = L.head[s]; k >= 0; k = L.arcs[k].link) {
            const AArc &a = L.arcs[k];
            if (a.next == kDead) continue;
            snprintf(b, sizeof b, "A %d %d %d %.9g %.9g ", s, a.next, a.label, a.g, a.a);
            text += b + tids(ws.seq.data() + a.t_off, a.t_len) + "\n";
        }
        if (L.is_final[s]) {
            snprintf(b, sizeof b, "F %d %.9g %.9g -\n", s, L.fin_g[s], L.fin_a[s]);
            text += b;
        }
    }
    return text;
}

}  // namespace vb
