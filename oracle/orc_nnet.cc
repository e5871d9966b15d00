// orc_nnet.cc — oracle (TEST INFRASTRUCTURE, see oracle.h): TDNN-F acoustic model forward on the CPU.
//
// Restates the nnet3 graph the reference trains in [REF training/local/chain/run_tdnn.sh:98-129] and
// evaluates after CollapseModel [REF src/batch_model.cc:46-48], [REF src/model.cc:227]:
//   tdnn1   : affine over [mfcc(t-2..t+2), ivector] (idct + batchnorm0 + delta-layer folded in), ReLU, BN
//   tdnnfK  : linear over [x(t-s), x(t)] -> bottleneck ; affine over [b(t), b(t+s)] , ReLU, BN, + 0.75 x(t)
//   prefinal-l (linear) ; prefinal-chain: affine, ReLU, BN, linear(+folded BN) ; output affine
//   no log-softmax [REF run_tdnn.sh:125]; outputs only at t = 0 (mod 3) [REF src/batch_model.cc:82].
// Like nnet3's compiler, only the time indices the requested outputs depend on are computed.
// Input is padded by repeating the first/last frame.  Accumulation is double.
//
// tensors[] order: tdnn1.w tdnn1.b tdnn1.bn_scale tdnn1.bn_offset,
//   then per tdnnf layer: linear.w affine.w affine.b bn_scale bn_offset,
//   then prefinal_l.w, prefinal.affine.w prefinal.affine.b prefinal.bn_scale prefinal.bn_offset,
//   prefinal.linear.w prefinal.linear.b, output.w output.b          (all W are [out][in] row-major)
#include "oracle.h"

#include <cstring>
#include <vector>

namespace {
struct Op {
    int in_node, in_dim, out_dim;
    std::vector<int> offs;
    const float *W, *b;
    bool relu;
    const float *bn_s, *bn_o;
    int byp_node;
    bool uses_ivec;
};
}  // namespace

extern "C" int orc_nnet_forward(const OrcNnetParams *p, const float *mfcc, int T, const float *ivecs,
                                const int *iv_index, float *loglikes) {
    if (T <= 0) return 0;
    const int F = p->feat_dim, I = p->ivec_dim, H = p->hidden, B = p->bottleneck;
    int ctx = 2;
    for (int k = 0; k < p->num_tdnnf; k++) ctx += p->strides[k];
    const int Lt = T + 2 * ctx;
    const float *const *tn = p->tensors;
    std::vector<Op> ops;
    int ti = 0;
    ops.push_back({0, F, H, {-2, -1, 0, 1, 2}, tn[0], tn[1], true, tn[2], tn[3], -1, true});
    ti = 4;
    int cur = 1;  // node index of current activation
    for (int k = 0; k < p->num_tdnnf; k++) {
        int s = p->strides[k];
        std::vector<int> lo = s ? std::vector<int>{-s, 0} : std::vector<int>{0};
        std::vector<int> ao = s ? std::vector<int>{0, s} : std::vector<int>{0};
        ops.push_back({cur, H, B, lo, tn[ti], nullptr, false, nullptr, nullptr, -1, false});
        ops.push_back({cur + 1, B, H, ao, tn[ti + 1], tn[ti + 2], true, tn[ti + 3], tn[ti + 4], cur, false});
        ti += 5;
        cur += 2;
    }
    const int PS = p->prefinal_small, PB = p->prefinal_big, NP = p->num_pdfs;
    ops.push_back({cur, H, PS, {0}, tn[ti], nullptr, false, nullptr, nullptr, -1, false});
    ops.push_back({cur + 1, PS, PB, {0}, tn[ti + 1], tn[ti + 2], true, tn[ti + 3], tn[ti + 4], -1, false});
    ops.push_back({cur + 2, PB, PS, {0}, tn[ti + 5], tn[ti + 6], false, nullptr, nullptr, -1, false});
    ops.push_back({cur + 3, PS, NP, {0}, tn[ti + 7], tn[ti + 8], false, nullptr, nullptr, -1, false});
    const int nn = (int)ops.size() + 1;
    // which time indices are needed at each node (index i = t + ctx)
    std::vector<std::vector<char>> need(nn, std::vector<char>(Lt, 0));
    for (int t = 0; t < T; t += 3) need[nn - 1][t + ctx] = 1;
    for (int o = (int)ops.size() - 1; o >= 0; o--) {
        const Op &op = ops[o];
        for (int i = 0; i < Lt; i++)
            if (need[o + 1][i]) {
                for (int off : op.offs) need[op.in_node][i + off] = 1;
                if (op.byp_node >= 0) need[op.byp_node][i] = 1;
            }
    }
    std::vector<std::vector<float>> act(nn);
    std::vector<int> dim(nn);
    dim[0] = F;
    act[0].assign((size_t)Lt * F, 0.f);
    for (int i = 0; i < Lt; i++) {
        int t = i - ctx;
        t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        memcpy(&act[0][(size_t)i * F], mfcc + (size_t)t * F, sizeof(float) * F);
    }
    std::vector<double> acc;
    std::vector<float> WT, x;
    for (size_t o = 0; o < ops.size(); o++) {
        const Op &op = ops[o];
        const int K = op.in_dim * (int)op.offs.size() + (op.uses_ivec ? I : 0), N = op.out_dim;
        dim[o + 1] = N;
        act[o + 1].assign((size_t)Lt * N, 0.f);
        WT.resize((size_t)K * N);
        for (int n = 0; n < N; n++)
            for (int k = 0; k < K; k++) WT[(size_t)k * N + n] = op.W[(size_t)n * K + k];
        acc.resize(N);
        x.resize(K);
        const std::vector<float> &in = act[op.in_node];
        for (int i = 0; i < Lt; i++) {
            if (!need[o + 1][i]) continue;
            int kk = 0;
            for (int off : op.offs) {
                memcpy(&x[kk], &in[(size_t)(i + off) * op.in_dim], sizeof(float) * op.in_dim);
                kk += op.in_dim;
            }
            if (op.uses_ivec) memcpy(&x[kk], ivecs + (size_t)iv_index[i - 2] * I, sizeof(float) * I);
            for (int n = 0; n < N; n++) acc[n] = op.b ? (double)op.b[n] : 0.0;
            for (int k = 0; k < K; k++) {
                const double xv = x[k];
                const float *w = &WT[(size_t)k * N];
                for (int n = 0; n < N; n++) acc[n] += xv * (double)w[n];
            }
            float *y = &act[o + 1][(size_t)i * N];
            for (int n = 0; n < N; n++) {
                float z = (float)acc[n];
                if (op.relu && z < 0.f) z = 0.f;
                if (op.bn_s) z = z * op.bn_s[n] + op.bn_o[n];
                if (op.byp_node >= 0) z += p->bypass_scale * act[op.byp_node][(size_t)i * N + n];
                y[n] = z;
            }
        }
    }
    int nout = 0;
    for (int t = 0; t < T; t += 3, nout++)
        memcpy(loglikes + (size_t)nout * NP, &act[nn - 1][(size_t)(t + ctx) * NP], sizeof(float) * NP);
    return nout;
}
