// orc_feat.cc — oracle (TEST INFRASTRUCTURE, see oracle.h): MFCC and online i-vector on the CPU.
//
// MFCC restates Kaldi's compute-mfcc-feats as configured by the reference
// [REF training/conf/mfcc.conf:1-7] (+ --dither=0 for determinism) and driven per stream by
// OnlineNnet2FeaturePipeline::AcceptWaveform [REF src/recognizer.cc:305-311]; samples stay in
// int16 units [REF src/recognizer.cc:274-275], [REF src/batch_recognizer.cc:153-155].
// The i-vector branch restates OnlineIvectorFeature with the options of [REF src/model.cc:250-260]
// and the extractor shape of [REF training/local/chain/run_ivector_common.sh:36,46,56];
// i-vectors are solved once per chunk by a direct Cholesky solve, which is what the reference's
// batch path does (SURVEY.md A4 vi).  FFT and all statistics are computed in double here so the
// oracle is the more exact side of every comparison.
#include "oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
const int kWin = 400, kShift = 160, kFft = 512, kMel = 40, kCep = 40;

struct MfccTables {
    std::vector<double> window, cosT, sinT, lifter;
    std::vector<float> mel_w;  // [kMel][256]
    std::vector<double> dct;   // [kCep][kMel]
    MfccTables() {
        window.resize(kWin);
        for (int i = 0; i < kWin; i++) window[i] = std::pow(0.5 - 0.5 * std::cos(2.0 * M_PI * i / (kWin - 1)), 0.85);
        cosT.resize(kFft / 2);
        sinT.resize(kFft / 2);
        for (int k = 0; k < kFft / 2; k++) {
            cosT[k] = std::cos(-2.0 * M_PI * k / kFft);
            sinT[k] = std::sin(-2.0 * M_PI * k / kFft);
        }
        auto mel = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
        const double low = 20.0, high = 8000.0 - 400.0, bw = 16000.0 / kFft;
        const double ml = mel(low), mh = mel(high), delta = (mh - ml) / (kMel + 1);
        mel_w.assign(kMel * 256, 0.f);
        for (int j = 0; j < kMel; j++) {
            double l = ml + j * delta, c = l + delta, r = c + delta;
            // same expression order as the model generator so the fp32 weights agree bit for bit
            c = ml + (j + 1) * delta;
            r = ml + (j + 2) * delta;
            for (int i = 0; i < 256; i++) {
                double m = mel(bw * i);
                if (m > l && m < r) mel_w[j * 256 + i] = (float)(m <= c ? (m - l) / (c - l) : (r - m) / (r - c));
            }
        }
        dct.resize(kCep * kMel);
        for (int n = 0; n < kMel; n++) dct[n] = std::sqrt(1.0 / kMel);
        for (int k = 1; k < kCep; k++)
            for (int n = 0; n < kMel; n++) dct[k * kMel + n] = std::sqrt(2.0 / kMel) * std::cos(M_PI / kMel * (n + 0.5) * k);
        lifter.resize(kCep);
        for (int i = 0; i < kCep; i++) lifter[i] = 1.0 + 0.5 * 22.0 * std::sin(M_PI * i / 22.0);
    }
};
const MfccTables &tables() {
    static MfccTables t;
    return t;
}

// in-place iterative radix-2 complex FFT, double
void fft512(double *re, double *im) {
    const MfccTables &t = tables();
    for (int i = 1, j = 0; i < kFft; i++) {
        int bit = kFft >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            std::swap(re[i], re[j]);
            std::swap(im[i], im[j]);
        }
    }
    for (int len = 2; len <= kFft; len <<= 1) {
        int step = kFft / len;
        for (int i = 0; i < kFft; i += len)
            for (int k = 0; k < len / 2; k++) {
                double wr = t.cosT[k * step], wi = t.sinT[k * step];
                int a = i + k, b = a + len / 2;
                double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr;
                im[b] = im[a] - xi;
                re[a] += xr;
                im[a] += xi;
            }
    }
}
}  // namespace

extern "C" int orc_num_frames(int64_t n) { return n < kWin ? 0 : (int)(1 + (n - kWin) / kShift); }

extern "C" int orc_mfcc(const int16_t *wave, int64_t n, float *out) {
    const MfccTables &t = tables();
    int T = orc_num_frames(n);
    std::vector<double> re(kFft), im(kFft);
    std::vector<float> frame(kWin);
    for (int f = 0; f < T; f++) {
        const int16_t *w = wave + (int64_t)f * kShift;
        // frame extraction follows Kaldi's float pipeline: copy, remove DC, pre-emphasise, window
        float sum = 0.f;
        for (int i = 0; i < kWin; i++) {
            frame[i] = (float)w[i];
            sum += frame[i];
        }
        float mean = sum / kWin;
        for (int i = 0; i < kWin; i++) frame[i] -= mean;
        for (int i = kWin - 1; i > 0; i--) frame[i] -= 0.97f * frame[i - 1];
        frame[0] -= 0.97f * frame[0];
        for (int i = 0; i < kFft; i++) {
            re[i] = i < kWin ? (double)(frame[i] * (float)t.window[i]) : 0.0;
            im[i] = 0.0;
        }
        fft512(re.data(), im.data());
        float logmel[kMel];
        for (int j = 0; j < kMel; j++) {
            double e = 0.0;
            for (int i = 0; i < 256; i++) {
                float wgt = t.mel_w[j * 256 + i];
                if (wgt != 0.f) e += (double)wgt * (re[i] * re[i] + im[i] * im[i]);
            }
            float ef = (float)e;
            if (ef < FLT_EPSILON) ef = FLT_EPSILON;
            logmel[j] = std::log(ef);
        }
        for (int k = 0; k < kCep; k++) {
            double c = 0.0;
            for (int j = 0; j < kMel; j++) c += t.dct[k * kMel + j] * logmel[j];
            out[(int64_t)f * kCep + k] = (float)(c * t.lifter[k]);
        }
    }
    return T;
}

// --------------------------------------------------------------------------------------------
// online i-vector
// --------------------------------------------------------------------------------------------
namespace {
// solve A x = b for SPD A (n x n, row-major, overwritten) by Cholesky
bool chol_solve(std::vector<double> &A, std::vector<double> &b, int n) {
    for (int j = 0; j < n; j++) {
        double d = A[j * n + j];
        for (int k = 0; k < j; k++) d -= A[j * n + k] * A[j * n + k];
        if (d <= 0) return false;
        d = std::sqrt(d);
        A[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = A[i * n + j];
            for (int k = 0; k < j; k++) s -= A[i * n + k] * A[j * n + k];
            A[i * n + j] = s / d;
        }
    }
    for (int i = 0; i < n; i++) {
        double s = b[i];
        for (int k = 0; k < i; k++) s -= A[i * n + k] * b[k];
        b[i] = s / A[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int k = i + 1; k < n; k++) s -= A[k * n + i] * b[k];
        b[i] = s / A[i * n + i];
    }
    return true;
}
}  // namespace

extern "C" int orc_ivectors(const OrcIvectorParams *p, const float *mfcc, int T, const int *ends, const int *avail,
                            int nchunks, float *out) {
    const int F = p->feat_dim, D = p->ivec_dim, G = p->num_gauss;
    const int L = p->splice_left, R = p->splice_right, S = (L + R + 1) * F;
    // derived extractor quantities (Kaldi IvectorExtractor::ComputeDerivedVars): Sigma_i^{-1} M_i, U_i = M_i^T Sigma_i^{-1} M_i
    std::vector<double> SiM((size_t)G * F * D), U((size_t)G * D * D);
    for (int g = 0; g < G; g++) {
        const float *M = p->M + (size_t)g * F * D, *Si = p->sigma_inv + (size_t)g * F * F;
        double *sm = &SiM[(size_t)g * F * D], *u = &U[(size_t)g * D * D];
        for (int a = 0; a < F; a++)
            for (int d = 0; d < D; d++) {
                double s = 0;
                for (int b = 0; b < F; b++) s += (double)Si[a * F + b] * M[b * D + d];
                sm[a * D + d] = s;
            }
        for (int d = 0; d < D; d++)
            for (int e = 0; e < D; e++) {
                double s = 0;
                for (int a = 0; a < F; a++) s += (double)M[a * D + d] * sm[a * D + e];
                u[d * D + e] = s;
            }
    }
    // sliding-window CMN (OnlineCmvn, norm_vars=false) with global-stats smoothing
    std::vector<double> norm((size_t)std::max(T, 1) * F), csum((size_t)(T + 1) * F, 0.0);
    for (int t = 0; t < T; t++)
        for (int d = 0; d < F; d++) csum[(size_t)(t + 1) * F + d] = csum[(size_t)t * F + d] + mfcc[(size_t)t * F + d];
    const double gcount = p->global_cmvn[F];
    for (int t = 0; t < T; t++) {
        int lo = std::max(0, t + 1 - p->cmn_window);
        double n = t + 1 - lo;
        double from_global = 0;
        if (n < p->cmn_window) from_global = std::min<double>(p->cmn_window - n, p->global_frames);
        for (int d = 0; d < F; d++) {
            double s = csum[(size_t)(t + 1) * F + d] - csum[(size_t)lo * F + d];
            if (from_global > 0) s += from_global / gcount * p->global_cmvn[d];
            norm[(size_t)t * F + d] = mfcc[(size_t)t * F + d] - s / (n + from_global);
        }
    }
    std::vector<double> lin(D, 0.0), quad((size_t)D * D, 0.0);
    lin[0] = p->prior_offset;
    for (int d = 0; d < D; d++) quad[d * D + d] = 1.0;
    double num_frames = 0.0;
    std::vector<double> xs(S), xn(S), fu(F), fn(F), ll(G);
    int done = 0;
    for (int c = 0; c < nchunks; c++) {
        int end = std::min(ends[c], T), last = avail[c] - 1;
        for (int t = done; t < end; t++) {
            for (int o = -L; o <= R; o++) {
                int tt = std::min(std::max(t + o, 0), last);
                for (int d = 0; d < F; d++) {
                    xs[(o + L) * F + d] = mfcc[(size_t)tt * F + d];
                    xn[(o + L) * F + d] = norm[(size_t)tt * F + d];
                }
            }
            for (int a = 0; a < F; a++) {
                const float *row = p->lda + (size_t)a * (S + 1);
                double su = row[S], sn = row[S];
                for (int k = 0; k < S; k++) {
                    su += (double)row[k] * xs[k];
                    sn += (double)row[k] * xn[k];
                }
                fu[a] = su;
                fn[a] = sn;
            }
            for (int g = 0; g < G; g++) {
                double s = p->gconsts[g];
                const float *mi = p->means_invvars + (size_t)g * F, *iv = p->inv_vars + (size_t)g * F;
                for (int a = 0; a < F; a++) s += (double)mi[a] * fn[a] - 0.5 * (double)iv[a] * fn[a] * fn[a];
                ll[g] = s;
            }
            // VectorToPosteriorEntry: top num_gselect, softmax, prune < min_post (keep the max), renormalise
            int ng = std::min(p->num_gselect, G);
            std::vector<int> idx(G);
            for (int g = 0; g < G; g++) idx[g] = g;
            std::partial_sort(idx.begin(), idx.begin() + ng, idx.end(), [&](int a, int b) {
                return ll[a] > ll[b] || (ll[a] == ll[b] && a < b);
            });
            double mx = ll[idx[0]], tot = 0;
            std::vector<double> post(ng);
            for (int k = 0; k < ng; k++) {
                post[k] = std::exp(ll[idx[k]] - mx);
                tot += post[k];
            }
            for (int k = 0; k < ng; k++) post[k] /= tot;
            double kept = 0;
            for (int k = 0; k < ng; k++) {
                if (k > 0 && post[k] < p->min_post) post[k] = 0;
                kept += post[k];
            }
            double tot_w = 0;
            for (int k = 0; k < ng; k++) {
                double w = post[k] / kept * p->posterior_scale;
                if (w == 0) continue;
                tot_w += w;
                const double *sm = &SiM[(size_t)idx[k] * F * D], *u = &U[(size_t)idx[k] * D * D];
                for (int a = 0; a < F; a++)
                    for (int d = 0; d < D; d++) lin[d] += w * sm[a * D + d] * fu[a];
                for (int i = 0; i < D * D; i++) quad[i] += w * u[i];
            }
            if (p->max_count > 0) {
                double o = std::max<double>(num_frames, p->max_count) / p->max_count;
                double nn = std::max<double>(num_frames + tot_w, p->max_count) / p->max_count;
                double ch = nn - o;
                if (ch != 0.0) {
                    lin[0] += p->prior_offset * ch;
                    for (int d = 0; d < D; d++) quad[d * D + d] += ch;
                }
            }
            num_frames += tot_w;
        }
        done = std::max(done, end);
        std::vector<double> A(quad), b(lin);
        if (num_frames > 0.0 && chol_solve(A, b, D)) {
            b[0] -= p->prior_offset;
        } else {
            std::fill(b.begin(), b.end(), 0.0);
        }
        for (int d = 0; d < D; d++) out[(size_t)c * D + d] = (float)b[d];
    }
    return nchunks;
}
