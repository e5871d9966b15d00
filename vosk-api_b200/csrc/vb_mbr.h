// vb_mbr.h — MinimumBayesRisk decoding of a word-aligned lattice (Xu, Povey, Mangu, Zhu: "Minimum Bayes Risk decoding and
// system combination based on a recursion for edit distance", as implemented by Kaldi lat/sausages.cc), the last stage of the
// reference's result path [REF src/batch_recognizer.cc:47-56]: MbrDecode's loop of AccStats (edit-distance forward / backward over
// the lattice against the current hypothesis R) and the per-position argmax, until the hypothesis no longer changes.
//
// mbr_solve() is plain arithmetic on flat arrays: the arc posteriors (the only place exp / log1p come in) are prepared by the
// caller.  The passes over the hypothesis positions are written element-wise (candidates that do not depend on the row; the one
// serial chain; the weighted sums; in the backward pass the carry chain, then the case-1 and the case-2 contributions) so that
// the compiler vectorizes what is independent; every floating-point operation and its order per element are those of the plain
// loops, products and sums rounded separately (no FMA contraction: the recursion breaks ties on exact comparisons).
// Round 2 also ran this body on the device (one warp per lattice, lanes over the positions; bit-identical, all GPU tests green):
// a lattice took 30-150 ms there and its CTAs competed with the search for SM residency — 1.8 s per bench step against 0.3 s —
// so the recursion stays on the host lattice pool.
#pragma once
#include <stdint.h>

#define VB_HD inline

namespace vb {

struct MbrArc {
    int word, start, end;  // 1-based nodes in topological order, the super-final node last
    float loglike;
    int t_len;             // frames (transition ids) on the arc
};
struct MbrAcc {            // one (position, word) accumulator of AccStats: gamma, tau_b, tau_e
    int word, pad;
    double g, tb, te;
};

struct MbrView {           // the problem: arcs grouped by end node (pre_off), in start-node order
    int N, n_arcs;
    const MbrArc *arcs;
    const int *pre_off;      // [N + 2]
    const int *state_times;  // [N + 1]
    const double *post;      // [n_arcs]  exp(alpha(start) + loglike - alpha(end))
};
struct MbrScratch {
    int w_cap;    // hypothesis positions + 1 the buffers hold
    int acc_cap;  // accumulators per position
    double *alpha_dash, *beta_dash;           // [(N + 1) * w_cap]
    char *b_all;                              // [n_arcs * w_cap]
    double *cq, *m12a, *m12b, *vala, *valb;   // [w_cap] rows of the arc in hand
    int *rh;                                  // [w_cap] decode step: the position's best word
    int *nz_lo, *nz_hi;                       // [N + 1] backward: range of a node's non-zero beta_dash cells
    MbrAcc *acc;                              // [w_cap * acc_cap]
    int *acc_n;                               // [w_cap]
    int *R, *R2;                              // [w_cap]
    float *tb, *te, *conf;                    // [w_cap] one-best begin / end / confidence
};
struct MbrOut {
    int status;   // 0 ok, 1 hypothesis longer than w_cap, 2 more words at a position than acc_cap
    int n_words, iters, q;
};
// results: out_word[n_words], out_tb / out_te / out_conf [n_words]

#define VB_MBR_LANE 0
#define VB_MBR_LANES 1
#define VB_MBR_SYNC()
#define VB_MBR_MUL(a, b) ((a) * (b))   // (the build has -ffp-contract=off)
#define VB_MBR_ADD(a, b) ((a) + (b))
#define VB_MBR_ANY(x) (x)

// hypothesis R0[nR0] (words, no epsilons) -> MBR one-best.  Returns through *out and the out_* arrays (sized w_cap).
VB_HD void mbr_solve(const MbrView &v, const MbrScratch &s, const int *R0, int nR0, MbrOut *out, int *out_word, float *out_tb, float *out_te,
                     float *out_conf) {
    const int lane = VB_MBR_LANE;
    const int N = v.N;
    const double kPen = 1.0 + 1.0e-05;  // l(a, eps, penalize = true)
    int *R = s.R, *R2 = s.R2;
    int nR = nR0;
    for (int i = lane; i < nR0; i += VB_MBR_LANES) R[i] = R0[i];
    VB_MBR_SYNC();
    int status = 0, iters = 0, n_best = 0, Q = 0;
    for (int counter = 0;; counter++) {
        // NormalizeEps: epsilons between all words and at both ends
        int nw = 0;
        for (int i = 0; i < nR; i++) nw += R[i] != 0;
        Q = 2 * nw + 1;
        if (Q + 1 > s.w_cap) {
            status = 1;
            break;
        }
        if (lane == 0) {
            int o = 0;
            R2[o++] = 0;
            for (int i = 0; i < nR; i++)
                if (R[i] != 0) {
                    R2[o++] = R[i];
                    R2[o++] = 0;
                }
        }
        VB_MBR_SYNC();
        {
            int *t = R;
            R = R2;
            R2 = t;
        }
        nR = Q;
        const int W = Q + 1;
        const int *r = R - 1;  // r[q], q = 1..Q
        // beta_dash is accumulated cell by cell and starts from zero; a row of alpha_dash is the sum over the node's incoming arcs, the
        // first of which assigns it (0 + x = x for the non-negative x here), so only the rows of nodes without incoming arcs are cleared
        for (long long i = lane; i < (long long)(N + 1) * W; i += VB_MBR_LANES) s.beta_dash[i] = 0.0;
        for (int n = 2; n <= N; n++)
            if (v.pre_off[n] == v.pre_off[n + 1])
                for (int q = lane; q <= Q; q += VB_MBR_LANES) s.alpha_dash[(long long)n * W + q] = 0.0;
        for (int q = lane; q <= Q; q += VB_MBR_LANES) {
            s.acc_n[q] = 0;
            s.cq[q] = (q >= 1 && r[q] != 0) ? 1.0 : 0.0;  // c(q) = l(eps, r(q)): what skipping reference position q costs
        }
        VB_MBR_SYNC();
        // ---- EditDistance (forward) ----
        if (lane == 0) {
            double *ad1 = s.alpha_dash + (long long)1 * W;
            ad1[0] = 0.0;
            for (int q = 1; q <= Q; q++) ad1[q] = ad1[q - 1] + (r[q] == 0 ? 0.0 : 1.0);
        }
        VB_MBR_SYNC();
        // One arc's row alpha_dash_arc(.) in three passes: the two candidates that do not depend on the row itself (a1: substitute /
        // match, a2: insert the arc's word) for all q — independent; then the chain a3(q) = row(q-1) + c(q), the only serial part;
        // then the weighted sum into the node.  Two arcs into the same node are taken together (their chains on two lanes).
        for (int n = 2; n <= N; n++) {
            double *adn = s.alpha_dash + (long long)n * W;
            int k = v.pre_off[n];
            const int kend = v.pre_off[n + 1];
            while (k < kend) {
                const int na = kend - k >= 2 ? 2 : 1;
                double first[2];
                for (int j = 0; j < na; j++) {
                    const MbrArc arc = v.arcs[k + j];
                    const double *ads = s.alpha_dash + (long long)arc.start * W;
                    const int w_a = arc.word;
                    const double l_eps = w_a == 0 ? 0.0 : kPen;  // l(w_a, eps, true)
                    double *m12 = j ? s.m12b : s.m12a;
                    char *bk = s.b_all + (long long)(k + j) * W;
                    for (int q = 1 + lane; q <= Q; q += VB_MBR_LANES) {
                        const double a1 = ads[q - 1] + (w_a == r[q] ? 0.0 : 1.0), a2 = ads[q] + l_eps;
                        const bool one = a1 <= a2;
                        m12[q] = one ? a1 : a2;
                        bk[q] = one ? 1 : 2;
                    }
                    first[j] = ads[0] + l_eps;
                }
                VB_MBR_SYNC();
                // The chain.  The normalized hypothesis alternates epsilon (odd q: c = 0) and word (even q: c = 1), and the row is never
                // negative, so at the odd positions a3 = prev + 0 is prev itself: the add is left out there (same value, same
                // comparison), which shortens the dependency chain from an add and a compare per position to three operations per two.
                if (na == 2) {  // two independent chains side by side (the chain is latency-bound)
                    char *bka = s.b_all + (long long)k * W, *bkb = s.b_all + (long long)(k + 1) * W;
                    double prev_a = first[0], prev_b = first[1];
                    s.vala[0] = prev_a;
                    s.valb[0] = prev_b;
                    for (int q = 1; q <= Q; q += 2) {
                        {
                            const double ma = s.m12a[q], mb = s.m12b[q];
                            const bool ta = !(ma <= prev_a), tb = !(mb <= prev_b);
                            prev_a = ta ? prev_a : ma;
                            prev_b = tb ? prev_b : mb;
                            s.vala[q] = prev_a;
                            s.valb[q] = prev_b;
                            bka[q] = ta ? 3 : bka[q];
                            bkb[q] = tb ? 3 : bkb[q];
                        }
                        if (q + 1 <= Q) {
                            const double a3a = prev_a + 1.0, a3b = prev_b + 1.0;
                            const double ma = s.m12a[q + 1], mb = s.m12b[q + 1];
                            const bool ta = !(ma <= a3a), tb = !(mb <= a3b);
                            prev_a = ta ? a3a : ma;
                            prev_b = tb ? a3b : mb;
                            s.vala[q + 1] = prev_a;
                            s.valb[q + 1] = prev_b;
                            bka[q + 1] = ta ? 3 : bka[q + 1];
                            bkb[q + 1] = tb ? 3 : bkb[q + 1];
                        }
                    }
                } else {
                    char *bk = s.b_all + (long long)k * W;
                    double prev = first[0];
                    s.vala[0] = prev;
                    for (int q = 1; q <= Q; q += 2) {
                        {
                            const double m = s.m12a[q];
                            const bool three = !(m <= prev);
                            prev = three ? prev : m;
                            s.vala[q] = prev;
                            bk[q] = three ? 3 : bk[q];
                        }
                        if (q + 1 <= Q) {
                            const double a3 = prev + 1.0;
                            const double m = s.m12a[q + 1];
                            const bool three = !(m <= a3);
                            prev = three ? a3 : m;
                            s.vala[q + 1] = prev;
                            bk[q + 1] = three ? 3 : bk[q + 1];
                        }
                    }
                }
                for (int j = 0; j < na; j++) {  // (arc order: the sums into the node are those of the plain loop)
                    const double p = v.post[k + j];
                    const double *val = j ? s.valb : s.vala;
                    if (k + j == v.pre_off[n]) {  // the node's first arc: 0 + p * val
                        for (int q = lane; q <= Q; q += VB_MBR_LANES) adn[q] = VB_MBR_MUL(p, val[q]);
                    } else {
                        for (int q = lane; q <= Q; q += VB_MBR_LANES) adn[q] = VB_MBR_ADD(adn[q], VB_MBR_MUL(p, val[q]));
                    }
                }
                VB_MBR_SYNC();
                k += na;
            }
        }
        // ---- backward: serial (every sum below is order-sensitive).  beta_dash is sparse — only the cells on optimal alignments are
        // non-zero — so every node carries the range of its non-zero cells and an arc visits that range only (a cell whose
        // beta_dash_arc is zero contributes nothing: skipping it changes no sum) ----
        auto acc_add = [&](int q, int word, double g, double tb, double te) -> bool {
            MbrAcc *aq = s.acc + (long long)q * s.acc_cap;
            const int na = s.acc_n[q];
            int i = 0;
            for (; i < na; i++)
                if (aq[i].word == word) break;
            if (i < na) {
                aq[i].g += g;
                aq[i].tb += tb;
                aq[i].te += te;
            } else if (na < s.acc_cap) {
                aq[na].word = word;
                aq[na].pad = 0;
                aq[na].g = g;
                aq[na].tb = tb;
                aq[na].te = te;
                s.acc_n[q] = na + 1;
            } else {
                return false;
            }
            return true;
        };
        for (int n = 0; n <= N; n++) {
            s.nz_lo[n] = W;
            s.nz_hi[n] = -1;
        }
        s.beta_dash[(long long)N * W + Q] = 1.0;
        s.nz_lo[N] = s.nz_hi[N] = Q;
        for (int n = N; n >= 2 && !status; n--) {
            const int lo_n = s.nz_lo[n], hi_n = s.nz_hi[n];
            if (hi_n < 0) continue;  // nothing reaches the end from here
            const double *bdn = s.beta_dash + (long long)n * W;
            for (int k = v.pre_off[n]; k < v.pre_off[n + 1] && !status; k++) {
                const MbrArc arc = v.arcs[k];
                const int s_a = arc.start, w_a = arc.word;
                double *bds = s.beta_dash + (long long)s_a * W;
                const char *b_arc = s.b_all + (long long)k * W;
                const double p = v.post[k];
                const double t_s = v.state_times[s_a], t_n = v.state_times[n];
                int lo_s = s.nz_lo[s_a], hi_s = s.nz_hi[s_a];
                double carry = 0.0;  // beta_dash_arc(q) accumulated from case 3 of q + 1
                int q = hi_n;
                for (; q >= 1; q--) {
                    if (q < lo_n && carry == 0.0) break;  // below the node's range and nothing carried: all zero from here on
                    const double b = carry + p * bdn[q];
                    carry = 0.0;
                    if (b == 0.0) continue;  // (adding an exact zero changes no sum; Kaldi's maps would only gain zero entries)
                    switch (b_arc[q]) {
                        case 1:
                            bds[q - 1] += b;
                            lo_s = q - 1 < lo_s ? q - 1 : lo_s;
                            hi_s = q - 1 > hi_s ? q - 1 : hi_s;
                            if (!acc_add(q, w_a, b, t_s * b, t_n * b)) status = 2;
                            break;
                        case 2:
                            bds[q] += b;
                            lo_s = q < lo_s ? q : lo_s;
                            hi_s = q > hi_s ? q : hi_s;
                            break;
                        case 3: {
                            carry = b;
                            const double tt = t_n * b;
                            if (!acc_add(q, 0, b, tt, tt)) status = 2;
                            break;
                        }
                    }
                }
                if (q == 0) {  // (an early exit means the carry and beta_dash(n, 0) are zero)
                    const double b0 = carry + p * bdn[0];
                    bds[0] += b0;
                    if (b0 != 0.0) {
                        lo_s = 0;
                        hi_s = hi_s < 0 ? 0 : hi_s;
                    }
                }
                s.nz_lo[s_a] = lo_s;
                s.nz_hi[s_a] = hi_s;
            }
        }
        if (!status) {
            const double *bd1 = s.beta_dash + (long long)1 * W;
            const double t1 = v.state_times[1];
            double carry = 0.0;
            for (int q = Q; q >= 1; q--) {
                const double b = carry + bd1[q];
                carry = b;
                const double tt = t1 * b;
                if (!acc_add(q, 0, b, tt, tt)) status = 2;
            }
        }
        if (status) break;
        // ---- MbrDecode step: the positions' best words side by side, then the serial part (R, delta, one-best times) on lane 0 ----
        for (int q = 1 + lane; q <= Q; q += VB_MBR_LANES) {
            const MbrAcc *aq = s.acc + (long long)q * s.acc_cap;
            const int na = s.acc_n[q];
            s.rh[q] = -1;
            if (na == 0) continue;
            // GammaCompare on the float posteriors: largest first, then the larger word id
            const MbrAcc *top = &aq[0];
            for (int i = 0; i < na; i++) {
                const float gx = (float)aq[i].g, gt = (float)top->g;
                if (gx > gt || (gx == gt && aq[i].word > top->word)) top = &aq[i];
            }
            double old_gamma = 0;
            const int rq = R[q - 1];
            for (int i = 0; i < na; i++)
                if (aq[i].word == rq) old_gamma = (float)aq[i].g;
            s.rh[q] = top->word;
            s.m12a[q] = old_gamma;
            s.m12b[q] = top->g;
            s.vala[q] = top->tb;
            s.valb[q] = top->te;
        }
        VB_MBR_SYNC();
        int done = 0;
        if (lane == 0) {
            double delta_Q = 0.0;
            n_best = 0;
            for (int q = 1; q <= Q; q++) {
                const int rhat = s.rh[q];
                if (rhat < 0) continue;
                const float g = (float)s.m12b[q];
                const double new_gamma = g;
                delta_Q += s.m12a[q] - new_gamma;
                R[q - 1] = rhat;
                if (rhat != 0) {
                    const int i = n_best++;
                    s.tb[i] = (float)(s.vala[q] / g);
                    s.te[i] = (float)(s.valb[q] / g);
                    if (i > 0 && s.te[i - 1] > s.tb[i]) {
                        // overlapping words: both share the union of their spans, split in proportion to their durations
                        const float prev_right = i > 1 ? s.te[i - 2] : 0.0f;
                        const float mn = s.tb[i - 1] < s.tb[i] ? s.tb[i - 1] : s.tb[i];
                        const float left = prev_right > mn ? prev_right : mn;
                        const float right = s.te[i - 1] > s.te[i] ? s.te[i - 1] : s.te[i];
                        const float first_dur = s.te[i - 1] - s.tb[i - 1];
                        const float second_dur = s.te[i] - s.tb[i];
                        const float mid = first_dur > 0 ? left + (right - left) * first_dur / (first_dur + second_dur) : left;
                        s.tb[i - 1] = left;
                        s.te[i - 1] = s.tb[i] = mid;
                        s.te[i] = right;
                    }
                    s.conf[i] = g;
                }
            }
            done = delta_Q == 0 || counter > 100;
        }
        iters = counter + 1;
        VB_MBR_SYNC();
        if (done) break;
    }
    if (lane == 0) {
        int n = 0;
        if (!status) {
            for (int i = 0; i < nR; i++)
                if (R[i] != 0 && n < n_best) {
                    out_word[n] = R[i];
                    out_tb[n] = s.tb[n];
                    out_te[n] = s.te[n];
                    out_conf[n] = s.conf[n];
                    n++;
                }
        }
        out->status = status;
        out->n_words = n;
        out->iters = iters;
        out->q = Q;
    }
}

}  // namespace vb
