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
    double *pb, *bq;                          // [w_cap] backward: p * beta_dash(n, q), beta_dash_arc(q)
    int *rh;                                  // [w_cap] decode step: the position's best word
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
        for (long long i = lane; i < (long long)(N + 1) * W; i += VB_MBR_LANES) {
            s.alpha_dash[i] = 0.0;
            s.beta_dash[i] = 0.0;
        }
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
                for (int j = 0; j < na; j++) {
                    if (lane != (VB_MBR_LANES > 1 ? j : 0)) continue;
                    const double *m12 = j ? s.m12b : s.m12a;
                    double *val = j ? s.valb : s.vala;
                    char *bk = s.b_all + (long long)(k + j) * W;
                    double prev = first[j];
                    val[0] = prev;
                    for (int q = 1; q <= Q; q++) {
                        const double a3 = prev + s.cq[q];
                        const double m = m12[q];
                        const bool three = !(m <= a3);
                        prev = three ? a3 : m;
                        val[q] = prev;
                        if (three) bk[q] = 3;
                    }
                }
                VB_MBR_SYNC();
                for (int j = 0; j < na; j++) {  // (arc order: the sums into the node are those of the plain loop)
                    const double p = v.post[k + j];
                    const double *val = j ? s.valb : s.vala;
                    for (int q = lane; q <= Q; q += VB_MBR_LANES) adn[q] = VB_MBR_ADD(adn[q], VB_MBR_MUL(p, val[q]));
                }
                VB_MBR_SYNC();
                k += na;
            }
        }
        // ---- backward.  Arcs are taken one at a time, in order (every sum below is order-sensitive across arcs); within an arc the
        // positions are independent except for (a) the carry of case 3 into q - 1 — a short serial chain on lane 0 — and (b) the two
        // contributions a cell of the start node can get from one arc: from q + 1 by case 1 and from q by case 2, in that order —
        // two passes.  The accumulators of a position belong to that position alone. ----
        if (lane == 0) s.beta_dash[(long long)N * W + Q] = 1.0;
        VB_MBR_SYNC();
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
        int bad = 0;
        for (int n = N; n >= 2 && !status; n--) {
            const double *bdn = s.beta_dash + (long long)n * W;
            for (int k = v.pre_off[n]; k < v.pre_off[n + 1]; k++) {
                const MbrArc arc = v.arcs[k];
                const int s_a = arc.start, w_a = arc.word;
                double *bds = s.beta_dash + (long long)s_a * W;
                const char *b_arc = s.b_all + (long long)k * W;
                const double p = v.post[k];
                const double t_s = v.state_times[s_a], t_n = v.state_times[n];
                for (int q = lane; q <= Q; q += VB_MBR_LANES) s.pb[q] = VB_MBR_MUL(p, bdn[q]);
                VB_MBR_SYNC();
                if (lane == 0) {
                    double carry = 0.0;  // beta_dash_arc(q) accumulated from case 3 of q + 1
                    for (int q = Q; q >= 1; q--) {
                        const double b = VB_MBR_ADD(carry, s.pb[q]);
                        s.bq[q] = b;
                        carry = (b != 0.0 && b_arc[q] == 3) ? b : 0.0;
                    }
                    s.bq[0] = VB_MBR_ADD(carry, s.pb[0]);
                }
                VB_MBR_SYNC();
                for (int q = 1 + lane; q <= Q; q += VB_MBR_LANES) {
                    const double b = s.bq[q];
                    if (b == 0.0) continue;  // (adding an exact zero changes no sum; Kaldi's maps would only gain zero entries)
                    const char c = b_arc[q];
                    if (c == 1) {
                        bds[q - 1] += b;
                        if (!acc_add(q, w_a, b, VB_MBR_MUL(t_s, b), VB_MBR_MUL(t_n, b))) bad = 1;
                    } else if (c == 3) {
                        const double tt = VB_MBR_MUL(t_n, b);
                        if (!acc_add(q, 0, b, tt, tt)) bad = 1;
                    }
                }
                VB_MBR_SYNC();
                for (int q = 1 + lane; q <= Q; q += VB_MBR_LANES) {
                    const double b = s.bq[q];
                    if (b != 0.0 && b_arc[q] == 2) bds[q] += b;
                }
                if (lane == 0) bds[0] += s.bq[0];
                VB_MBR_SYNC();
            }
            if (VB_MBR_ANY(bad)) status = 2;
        }
        if (!status) {
            const double *bd1 = s.beta_dash + (long long)1 * W;
            const double t1 = v.state_times[1];
            if (lane == 0) {
                double carry = 0.0;
                for (int q = Q; q >= 1; q--) {
                    const double b = carry + bd1[q];
                    carry = b;
                    s.bq[q] = b;
                }
            }
            VB_MBR_SYNC();
            for (int q = 1 + lane; q <= Q; q += VB_MBR_LANES) {
                const double b = s.bq[q];
                const double tt = VB_MBR_MUL(t1, b);
                if (!acc_add(q, 0, b, tt, tt)) bad = 1;
            }
            if (VB_MBR_ANY(bad)) status = 2;
        }
        VB_MBR_SYNC();
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
