// kernels_feat.cu — K1: fused framing / DC removal / pre-emphasis / Povey window / 512-pt FFT / power /
// mel / log / DCT / lifter, and K1c: online CMN + i-vector statistics + per-chunk solve.
//
// Replaces Kaldi cudafeat as driven by the reference's batch pipeline (use_gpu_feature_extraction=true,
// feature_type="mfcc", mfcc.conf, ivector.conf — [REF src/batch_model.cc:73-77]); options from
// [REF training/conf/mfcc.conf:1-7] and [REF src/model.cc:250-260].
//
// Layout: one CTA per lane (= one stream's chunk).  The lane's samples (carried tail + new chunk) are staged
// once in shared memory with coalesced 16-byte loads; each warp then owns whole frames, so the FFT, the mel
// projection and the DCT never leave the SM.  Output rows go straight into the acoustic model's input ring
// (edge padding rows included), so no separate "build batch with context" pass exists.
#include <cfloat>

#include "vb_kernels.h"

namespace vb {

__device__ __forceinline__ int ring_slot(const NodeDesc &n, int t) { return ((t - n.t_start) / n.step) & (n.ring - 1); }
__device__ __forceinline__ float *ring_row(const NodeDesc &n, int ch, int t) {
    return n.buf + ((size_t)ch * n.ring + ring_slot(n, t)) * n.dim;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kFeatThreads = 256;
constexpr int kFeatWarps = kFeatThreads / 32;

// dynamic smem layout (floats): wave[kCarryMax + spc] | fft[kFeatWarps][2][512] | logmel[kFeatWarps][40]
//                               | window[400] | tw[512] | dct_t[1600] | lifter[40]
extern "C" int vbk_feat_smem_bytes(int spc) {
    return (int)sizeof(float) * (kCarryMax + spc + kFeatWarps * 1024 + kFeatWarps * 40 + 400 + 512 + 1600 + 40);
}

__global__ void __launch_bounds__(kFeatThreads) mfcc_kernel(FeatArgs a) {
    extern __shared__ __align__(16) float sm[];
    const LaneDesc ln = a.lanes[blockIdx.x];
    const int spc = a.samples_per_chunk;
    float *wave = sm;
    float *fft = wave + kCarryMax + spc;
    float *logmel = fft + kFeatWarps * 1024;
    float *window = logmel + kFeatWarps * 40;
    float *tw = window + 400;
    float *dct_t = tw + 512;
    float *lifter = dct_t + 1600;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 400; i += kFeatThreads) window[i] = a.tab.window[i];
    for (int i = tid; i < 512; i += kFeatThreads) tw[i] = a.tab.twiddle[i];
    for (int i = tid; i < 1600; i += kFeatThreads) dct_t[i] = a.tab.dct_t[i];
    if (tid < 40) lifter[tid] = a.tab.lifter[tid];
    // stage samples: carried tail then the new chunk (int16 -> float, values unscaled [REF src/batch_recognizer.cc:153-155])
    int16_t *carry = a.carry + (size_t)ln.channel * kCarryMax;
    for (int i = tid; i < ln.carry; i += kFeatThreads) wave[i] = (float)carry[i];
    {
        const int16_t *src = a.staging + (size_t)ln.src_row * a.src_stride + ln.src_off;
        const int n8 = (reinterpret_cast<uintptr_t>(src) & 15) ? 0 : ln.n_samples >> 3;
        const int4 *src4 = reinterpret_cast<const int4 *>(src);  // 16-byte path when the row start is aligned
        for (int i = tid; i < n8; i += kFeatThreads) {
            int4 v = __ldg(src4 + i);
            const short *h = reinterpret_cast<const short *>(&v);
            float *d = wave + ln.carry + i * 8;
#pragma unroll
            for (int j = 0; j < 8; j++) d[j] = (float)h[j];
        }
        for (int i = (n8 << 3) + tid; i < ln.n_samples; i += kFeatThreads) wave[ln.carry + i] = (float)src[i];
    }
    __syncthreads();
    const int total = ln.carry + ln.n_samples;
    const int nf = ln.frames_after - ln.frames_before;
    float *re = fft + warp * 1024, *im = re + 512;
    for (int f = warp; f < nf; f += kFeatWarps) {
        const float *w = wave + f * kFrameShift;
        float s = 0.f;
        for (int i = lane; i < kFrameLen; i += 32) s += w[i];
        const float mean = warp_sum(s) / kFrameLen;
        // bit-reversed scatter of the windowed, pre-emphasised frame; zero padding 400..511
        for (int i = lane; i < kFftSize; i += 32) {
            float v = 0.f;
            if (i < kFrameLen) {
                float x = w[i] - mean;
                float xp = (i > 0 ? w[i - 1] : w[0]) - mean;
                v = (x - 0.97f * xp) * window[i];
            }
            int r = __brev((unsigned)i) >> 23;
            re[r] = v;
            im[r] = 0.f;
        }
        __syncwarp();
#pragma unroll 1
        for (int len = 2; len <= kFftSize; len <<= 1) {
            const int half = len >> 1, tstep = kFftSize / len;
            for (int b = lane; b < kFftSize / 2; b += 32) {
                int k = b & (half - 1);
                int i0 = ((b - k) << 1) + k, i1 = i0 + half;
                float wr = tw[2 * k * tstep], wi = tw[2 * k * tstep + 1];
                float xr = re[i1] * wr - im[i1] * wi, xi = re[i1] * wi + im[i1] * wr;
                float ar = re[i0], ai = im[i0];
                re[i1] = ar - xr;
                im[i1] = ai - xi;
                re[i0] = ar + xr;
                im[i0] = ai + xi;
            }
            __syncwarp();
        }
        for (int i = lane; i < 256; i += 32) re[i] = re[i] * re[i] + im[i] * im[i];
        __syncwarp();
        for (int j = lane; j < kNumMel; j += 32) {
            const int st = a.tab.mel_start[j], n = a.tab.mel_len[j];
            const float *mw = a.tab.mel_w + j * kMelMaxLen;
            float e = 0.f;
            for (int i = 0; i < n; i++) e += __ldg(mw + i) * re[st + i];
            logmel[warp * 40 + j] = logf(fmaxf(e, FLT_EPSILON));
        }
        __syncwarp();
        const int t = ln.frames_before + f;
        float *out = ring_row(a.in_node, ln.channel, t);
        for (int k = lane; k < kNumCeps; k += 32) {
            float c = 0.f;
#pragma unroll 8
            for (int j = 0; j < kNumMel; j++) c += dct_t[j * 40 + k] * logmel[warp * 40 + j];
            out[k] = c * lifter[k];
        }
        __syncwarp();
    }
    __syncthreads();
    // new carry = samples from the start of the next frame
    const int next_start = nf > 0 ? nf * kFrameShift : 0;
    const int keep = (ln.frames_after > 0 || total >= kFrameLen) ? total - next_start : total;
    if (!ln.last)
        for (int i = tid; i < keep && i < kCarryMax; i += kFeatThreads) carry[i] = (int16_t)wave[next_start + i];
    // edge padding of the model input: repeat first / last frame over the model context
    if (ln.first && ln.frames_after > 0 && ln.frames_before == 0) {
        const float *src = ring_row(a.in_node, ln.channel, 0);
        for (int i = tid; i < a.context * 40; i += kFeatThreads) ring_row(a.in_node, ln.channel, -a.context + i / 40)[i % 40] = src[i % 40];
    }
    if (ln.last && ln.frames_after > 0) {
        const float *src = ring_row(a.in_node, ln.channel, ln.frames_after - 1);
        for (int i = tid; i < a.context * 40; i += kFeatThreads) ring_row(a.in_node, ln.channel, ln.frames_after + i / 40)[i % 40] = src[i % 40];
    }
}

// K0: resampling of one step's non-16 kHz segments.  One CTA per (segment, 1024-output tile); the segment's phase table
// (first tap index, tap count, weights) is read through the read-only path, the raw samples are staged by the host.  The
// arithmetic is the host resampler's, operation for operation (vb_result.cc LinearResampler::resample_flush: products and
// sums in fp32, taps in order, taps outside the call skipped, no fused multiply-add), so the int16 samples the feature
// kernel sees are bit-identical to the host path's.
constexpr int kResampleTile = 1024;
__global__ void __launch_bounds__(256) resample_kernel(ResampleArgs a) {
    const ResampleSeg sg = a.segs[blockIdx.x];
    const ResampleTable tb = a.tables[sg.table];
    const int16_t *raw = a.raw + sg.raw_off;
    int16_t *dst = a.staging + (size_t)sg.lane * a.samples_per_chunk + sg.out_pos;
    const int j0 = blockIdx.y * kResampleTile;
    for (int j = j0 + threadIdx.x; j < sg.n_out && j < j0 + kResampleTile; j += blockDim.x) {
        const long long so = (long long)sg.out_first + j;
        const long long unit = so / tb.out_unit;
        const int wrapped = (int)(so - unit * tb.out_unit);
        const long long first = (long long)__ldg(tb.first_index + wrapped) + unit * tb.in_unit;
        const int nt = __ldg(tb.n_taps + wrapped);
        const float *w = tb.weights + (size_t)wrapped * tb.max_taps;
        float acc = 0.f;
        for (int i = 0; i < nt; i++) {
            const long long idx = first + i;
            if (idx >= 0 && idx < sg.n_in) acc = __fadd_rn(acc, __fmul_rn(__ldg(w + i), (float)raw[idx - sg.in_base]));
        }
        acc = fmaxf(-32768.f, fminf(32767.f, acc));
        dst[j] = (int16_t)__float2int_rn(acc);
    }
}

extern "C" cudaError_t vbk_resample(const ResampleArgs *a, cudaStream_t s) {
    if (a->num_segs <= 0) return cudaSuccess;
    dim3 grid(a->num_segs, (a->samples_per_chunk + kResampleTile - 1) / kResampleTile);
    resample_kernel<<<grid, 256, 0, s>>>(*a);
    return cudaGetLastError();
}

extern "C" cudaError_t vbk_mfcc(const FeatArgs *a, cudaStream_t s) {
    if (a->num_lanes <= 0) return cudaSuccess;
    static int configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int smem = vbk_feat_smem_bytes(a->samples_per_chunk);
    if (dev < 16 && configured[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(mfcc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    mfcc_kernel<<<a->num_lanes, kFeatThreads, smem, s>>>(*a);
    return cudaGetLastError();
}

// =====================================================================================================
// K1c — online CMN (600-frame window, global-stats smoothing), splice +-3, LDA, diag-UBM top-N posteriors,
// i-vector statistics (double accumulation), per-chunk Cholesky solve.  One CTA per lane.  The frames of the chunk
// are processed in tiles of kIvTB frames by block-wide passes whose operands live in shared memory / registers:
//   LDA    : thread = (output dim, frame group); the 7-frame input window of the tile is staged once
//   UBM    : thread = 2 Gaussians x all frames of the tile (32 register accumulators); every (mean, inv-var) pair is
//            loaded once per tile and the normalised features are broadcast from shared memory as float4
//   select : warp per frame (top-N by repeated warp arg-max), posteriors, occupancies
//   linear : thread = (i-vector dim, (frame, Gaussian) group)
// =====================================================================================================
constexpr int kIvThreads = 256;
constexpr int kIvWarps = kIvThreads / 32;
constexpr int kIvTB = 16;           // frames per tile
constexpr int kMaxGselect = 8;

struct IvLayout {  // shared-memory carve-up, in floats (doubles at the end)
    int win_raw, win_nrm, fu, fnT, fn2T, ll, gamma, sel_g, sel_w, lin_part, glist, total_floats;
    int cmn_tile;  // frames per CMN staging tile (aliases ll)
};
__host__ __device__ inline IvLayout iv_layout(int G, int F, int D) {
    IvLayout L;
    int o = 0;
    L.win_raw = o; o += (kIvTB + 6) * F;
    L.win_nrm = o; o += (kIvTB + 6) * F;
    L.fu = o; o += kIvTB * F;
    o = (o + 3) & ~3;
    L.fnT = o; o += F * kIvTB;
    L.fn2T = o; o += F * kIvTB;
    L.ll = o; o += kIvTB * G;
    L.gamma = o; o += kIvWarps * G;
    L.sel_g = o; o += kIvTB * kMaxGselect;
    L.sel_w = o; o += kIvTB * kMaxGselect;
    L.lin_part = o; o += kIvThreads;
    L.glist = o; o += G + 8;
    L.total_floats = (o + 1) & ~1;
    L.cmn_tile = (kIvTB * G) / (2 * F);
    return L;
}
static int ivec_smem_bytes(int G, int F, int D) {
    return iv_layout(G, F, D).total_floats * 4 + (D * D + D) * 8;
}

__global__ void __launch_bounds__(kIvThreads, 2) ivector_kernel(IvecArgs a) {
    extern __shared__ __align__(16) float smf[];
    const LaneDesc ln = a.lanes[blockIdx.x];
    const IvecModel &m = a.m;
    const int F = m.feat_dim, D = m.ivec_dim, G = m.num_gauss, S = m.splice_dim;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const IvLayout L = iv_layout(G, F, D);
    float *win_raw = smf + L.win_raw, *win_nrm = smf + L.win_nrm, *fu = smf + L.fu, *fnT = smf + L.fnT, *fn2T = smf + L.fn2T;
    float *ll = smf + L.ll, *gamma = smf + L.gamma, *sel_w = smf + L.sel_w, *lin_part = smf + L.lin_part;
    int *sel_g = reinterpret_cast<int *>(smf + L.sel_g), *glist = reinterpret_cast<int *>(smf + L.glist);
    double *A = reinterpret_cast<double *>(smf + L.total_floats);
    double *bvec = A + D * D;
    __shared__ double s_totw[kIvWarps];
    __shared__ int s_nlist;
    const int ch = ln.channel;
    double *cm_sum = a.st.cmvn_sum + (size_t)ch * F;
    float *nring = a.st.norm_ring + (size_t)ch * kNormRing * F;
    double *lin = a.st.lin + (size_t)ch * D;
    double *quad = a.st.quad + (size_t)ch * D * D;
    if (ln.first) {  // stream start: prior  (OnlineIvectorEstimationStats ctor)
        for (int i = tid; i < D * D; i += kIvThreads) quad[i] = (i / D == i % D) ? 1.0 : 0.0;
        for (int i = tid; i < D; i += kIvThreads) lin[i] = i == 0 ? (double)m.prior_offset : 0.0;
        if (tid < F) cm_sum[tid] = 0.0;
        if (tid == 0) a.st.num_frames[ch] = 0.0;
    }
    for (int i = tid; i < kIvWarps * G; i += kIvThreads) gamma[i] = 0.f;
    if (tid < kIvWarps) s_totw[tid] = 0.0;
    __syncthreads();
    // ---- A. sliding-window CMN over the new frames: rows staged with coalesced loads, then thread per dimension ----
    {
        float *xin = ll, *xold = ll + L.cmn_tile * F;
        for (int t0 = ln.frames_before; t0 < ln.frames_after; t0 += L.cmn_tile) {
            const int n = min(L.cmn_tile, ln.frames_after - t0);
            for (int i = tid; i < n * F; i += kIvThreads) {
                const int t = t0 + i / F, d = i % F;
                xin[i] = ring_row(a.in_node, ch, t)[d];
                xold[i] = t >= m.cmn_window ? ring_row(a.in_node, ch, t - m.cmn_window)[d] : 0.f;
            }
            __syncthreads();
            if (tid < F) {
                double s = t0 == ln.frames_before ? cm_sum[tid] : cm_sum[tid];
                for (int j = 0; j < n; j++) {
                    const int t = t0 + j;
                    const float x = xin[j * F + tid];
                    s += (double)x;
                    if (t >= m.cmn_window) s -= (double)xold[j * F + tid];
                    const double nn = t + 1 < m.cmn_window ? t + 1 : m.cmn_window;
                    const double fg = nn < m.cmn_window ? fmin((double)m.cmn_window - nn, (double)m.global_frames) : 0.0;
                    double tot = s;
                    if (fg > 0.0) tot += fg / m.gcmvn_count * m.gcmvn_sum[tid];
                    nring[(t & (kNormRing - 1)) * F + tid] = (float)((double)x - tot / (nn + fg));
                }
                cm_sum[tid] = s;
            }
            __syncthreads();
        }
    }
    // ---- B. tiles of frames: splice, LDA (raw + normalised), UBM posteriors, statistics ----
    const int last_avail = ln.frames_after - 1;
    const int ng = min(m.num_gselect, G);
    const int ngF = kIvThreads / F, ngD = kIvThreads / D;
    float lin_acc = 0.f;
    double totw = 0.0;
    for (int tb = ln.iv_end_before; tb < ln.iv_end_after; tb += kIvTB) {
        const int nb = min(kIvTB, ln.iv_end_after - tb);
        for (int i = tid; i < (nb + 6) * F; i += kIvThreads) {
            const int r = i / F, d = i % F;
            const int tt = min(max(tb + r - 3, 0), last_avail);
            win_raw[i] = ring_row(a.in_node, ch, tt)[d];
            win_nrm[i] = nring[(tt & (kNormRing - 1)) * F + d];
        }
        __syncthreads();
        // B1. LDA: out[t][d] = offset[d] + sum_k lda[k][d] * splice(t)[k]; splice(t)[o*F + kd] = win[t + o][kd]
        if (tid < ngF * F) {
            const int d = tid % F, tg = tid / F;
            for (int t0 = tg; t0 < nb; t0 += 3 * ngF) {
                const int t1 = t0 + ngF, t2 = t0 + 2 * ngF;
                const int r0 = t0, r1 = min(t1, nb - 1), r2 = min(t2, nb - 1);  // rows beyond the tile repeat the last one (discarded)
                const float off = __ldg(m.lda_t + (size_t)S * F + d);
                float u0 = off, u1 = off, u2 = off, n0 = off, n1 = off, n2 = off;
                for (int o = 0; o < 7; o++) {
                    const float *w = m.lda_t + (size_t)o * F * F + d;
                    const float *xr0 = win_raw + (r0 + o) * F, *xr1 = win_raw + (r1 + o) * F, *xr2 = win_raw + (r2 + o) * F;
                    const float *xn0 = win_nrm + (r0 + o) * F, *xn1 = win_nrm + (r1 + o) * F, *xn2 = win_nrm + (r2 + o) * F;
#pragma unroll 4
                    for (int kd = 0; kd < F; kd++) {
                        const float wv = __ldg(w + (size_t)kd * F);
                        u0 = fmaf(wv, xr0[kd], u0);
                        u1 = fmaf(wv, xr1[kd], u1);
                        u2 = fmaf(wv, xr2[kd], u2);
                        n0 = fmaf(wv, xn0[kd], n0);
                        n1 = fmaf(wv, xn1[kd], n1);
                        n2 = fmaf(wv, xn2[kd], n2);
                    }
                }
                fu[t0 * F + d] = u0;
                fnT[d * kIvTB + t0] = n0;
                fn2T[d * kIvTB + t0] = n0 * n0;
                if (t1 < nb) {
                    fu[t1 * F + d] = u1;
                    fnT[d * kIvTB + t1] = n1;
                    fn2T[d * kIvTB + t1] = n1 * n1;
                }
                if (t2 < nb) {
                    fu[t2 * F + d] = u2;
                    fnT[d * kIvTB + t2] = n2;
                    fn2T[d * kIvTB + t2] = n2 * n2;
                }
            }
        }
        __syncthreads();
        // B2. diag-UBM log-likelihoods of the tile: ll[t][g] = gconst[g] + sum_d (mi[d][g] x - 0.5 iv[d][g] x^2)
        for (int g0 = tid; g0 < G; g0 += 2 * kIvThreads) {
            const int g1 = g0 + kIvThreads;
            const bool h1 = g1 < G;
            const int g1c = h1 ? g1 : g0;
            float acc0[kIvTB], acc1[kIvTB];
            const float c0 = __ldg(m.gconsts + g0), c1 = __ldg(m.gconsts + g1c);
#pragma unroll
            for (int t = 0; t < kIvTB; t++) {
                acc0[t] = c0;
                acc1[t] = c1;
            }
            for (int d = 0; d < F; d++) {
                const float mi0 = __ldg(m.mi_t + (size_t)d * G + g0), mi1 = __ldg(m.mi_t + (size_t)d * G + g1c);
                const float iv0 = -0.5f * __ldg(m.iv_t + (size_t)d * G + g0), iv1 = -0.5f * __ldg(m.iv_t + (size_t)d * G + g1c);
                const float4 *x4 = reinterpret_cast<const float4 *>(fnT + d * kIvTB), *q4 = reinterpret_cast<const float4 *>(fn2T + d * kIvTB);
#pragma unroll
                for (int q = 0; q < kIvTB / 4; q++) {
                    const float4 x = x4[q], xx = q4[q];
                    acc0[4 * q + 0] = fmaf(iv0, xx.x, fmaf(mi0, x.x, acc0[4 * q + 0]));
                    acc0[4 * q + 1] = fmaf(iv0, xx.y, fmaf(mi0, x.y, acc0[4 * q + 1]));
                    acc0[4 * q + 2] = fmaf(iv0, xx.z, fmaf(mi0, x.z, acc0[4 * q + 2]));
                    acc0[4 * q + 3] = fmaf(iv0, xx.w, fmaf(mi0, x.w, acc0[4 * q + 3]));
                    acc1[4 * q + 0] = fmaf(iv1, xx.x, fmaf(mi1, x.x, acc1[4 * q + 0]));
                    acc1[4 * q + 1] = fmaf(iv1, xx.y, fmaf(mi1, x.y, acc1[4 * q + 1]));
                    acc1[4 * q + 2] = fmaf(iv1, xx.z, fmaf(mi1, x.z, acc1[4 * q + 2]));
                    acc1[4 * q + 3] = fmaf(iv1, xx.w, fmaf(mi1, x.w, acc1[4 * q + 3]));
                }
            }
#pragma unroll
            for (int t = 0; t < kIvTB; t++) {
                ll[t * G + g0] = acc0[t];
                if (h1) ll[t * G + g1] = acc1[t];
            }
        }
        __syncthreads();
        // B3. per frame (warp): top-N Gaussians, posteriors (min_post pruning, posterior_scale), occupancies
        for (int t = warp; t < nb; t += kIvWarps) {
            const int per = (G + 31) / 32;  // G <= 1024
            float llv[32];
#pragma unroll
            for (int j = 0; j < 32; j++) {
                const int g = lane + 32 * j;
                llv[j] = (j < per && g < G) ? ll[t * G + g] : -FLT_MAX;
            }
            float best_v[kMaxGselect];
            int best_g[kMaxGselect];
#pragma unroll
            for (int r = 0; r < kMaxGselect; r++) {
                best_v[r] = -FLT_MAX;
                best_g[r] = 0;
                if (r >= ng) continue;
                float bv = -FLT_MAX;
                int bg = 0x7fffffff;
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (llv[j] > bv) { bv = llv[j]; bg = lane + 32 * j; }
                for (int o = 16; o; o >>= 1) {
                    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    int og = __shfl_xor_sync(0xffffffffu, bg, o);
                    if (ov > bv || (ov == bv && og < bg)) { bv = ov; bg = og; }
                }
                best_v[r] = bv;
                best_g[r] = bg;
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if ((bg & 31) == lane && (bg >> 5) == j) llv[j] = -FLT_MAX;
            }
            float post[kMaxGselect], tot = 0.f;
#pragma unroll
            for (int r = 0; r < kMaxGselect; r++) {
                post[r] = r < ng ? expf(best_v[r] - best_v[0]) : 0.f;
                tot += post[r];
            }
            float kept = 0.f;
#pragma unroll
            for (int r = 0; r < kMaxGselect; r++) {
                if (r >= ng) continue;
                post[r] /= tot;
                if (r > 0 && post[r] < m.min_post) post[r] = 0.f;
                kept += post[r];
            }
#pragma unroll
            for (int r = 0; r < kMaxGselect; r++) {
                if (r >= ng) continue;
                const float w = post[r] / kept * m.posterior_scale;
                if (lane == 0) {
                    sel_g[t * kMaxGselect + r] = best_g[r];
                    sel_w[t * kMaxGselect + r] = w;
                    if (w != 0.f) gamma[warp * G + best_g[r]] += w;
                }
                if (w != 0.f) totw += (double)w;
            }
        }
        __syncthreads();
        // B4. linear term: lin[d] += w * sum_a SiM[g][a][d] * fu[t][a]   (thread = dim x (frame, Gaussian) group)
        if (tid < ngD * D) {
            const int d = tid % D, pg = tid / D;
            for (int p = pg; p < nb * ng; p += ngD) {
                const int t = p / ng, r = p % ng;
                const float w = sel_w[t * kMaxGselect + r];
                if (w == 0.f) continue;
                const float *sm_g = m.sim + (size_t)sel_g[t * kMaxGselect + r] * F * D + d;
                const float *f = fu + t * F;
                float s0 = 0.f;
#pragma unroll 8
                for (int q = 0; q < F; q++) s0 = fmaf(__ldg(sm_g + (size_t)q * D), f[q], s0);
                lin_acc += w * s0;
            }
        }
        __syncthreads();
    }
    lin_part[tid] = tid < ngD * D ? lin_acc : 0.f;
    if (lane == 0) s_totw[warp] = totw;
    __syncthreads();
    // ---- C. fold the chunk's statistics into the channel state (fixed summation order) ----
    if (tid == 0) s_nlist = 0;
    __syncthreads();
    for (int g = tid; g < G; g += kIvThreads) {
        float s = 0.f;
        for (int w = 0; w < kIvWarps; w++) s += gamma[w * G + g];
        gamma[g] = s;  // row 0 now holds the per-gaussian occupancy of this chunk
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int g = 0; g < G; g++)
            if (gamma[g] != 0.f) glist[n++] = g;
        s_nlist = n;
    }
    __syncthreads();
    const int nl = s_nlist;
    double tw = 0.0;
    for (int w = 0; w < kIvWarps; w++) tw += s_totw[w];
    const double nf_old = a.st.num_frames[ch], nf_new = nf_old + tw;
    double change = 0.0;
    if (m.max_count > 0.f)
        change = fmax(nf_new, (double)m.max_count) / m.max_count - fmax(nf_old, (double)m.max_count) / m.max_count;
    for (int i = tid; i < D * D; i += kIvThreads) {
        double q = 0.0;
        for (int k = 0; k < nl; k++) {
            int g = glist[k];
            q += (double)gamma[g] * (double)__ldg(m.U + (size_t)g * D * D + i);
        }
        if (i / D == i % D) q += change;
        double v = quad[i] + q;
        quad[i] = v;
        A[i] = v;
    }
    for (int d = tid; d < D; d += kIvThreads) {
        double s = 0.0;
        for (int pg = 0; pg < ngD; pg++) s += (double)lin_part[pg * D + d];
        if (d == 0) s += (double)m.prior_offset * change;
        double v = lin[d] + s;
        lin[d] = v;
        bvec[d] = v;
    }
    __syncthreads();
    if (tid == 0) a.st.num_frames[ch] = nf_new;
    // ---- D. solve quad * x = lin by Cholesky (what the reference's batch path does per chunk) ----
    float *out = a.st.ivec + (size_t)ch * D;
    if (nf_new <= 0.0) {
        for (int d = tid; d < D; d += kIvThreads) out[d] = 0.f;
        return;
    }
    for (int j = 0; j < D; j++) {
        if (tid == 0) {
            double d = A[j * D + j];
            for (int k = 0; k < j; k++) d -= A[j * D + k] * A[j * D + k];
            A[j * D + j] = sqrt(fmax(d, 1e-300));
        }
        __syncthreads();
        for (int i = j + 1 + tid; i < D; i += kIvThreads) {
            double s = A[i * D + j];
            for (int k = 0; k < j; k++) s -= A[i * D + k] * A[j * D + k];
            A[i * D + j] = s / A[j * D + j];
        }
        __syncthreads();
    }
    if (warp == 0) {
        for (int i = 0; i < D; i++) {
            double s = 0.0;
            for (int k = lane; k < i; k += 32) s += A[i * D + k] * bvec[k];
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) bvec[i] = (bvec[i] - s) / A[i * D + i];
            __syncwarp();
        }
        for (int i = D - 1; i >= 0; i--) {
            double s = 0.0;
            for (int k = i + 1 + lane; k < D; k += 32) s += A[k * D + i] * bvec[k];
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) bvec[i] = (bvec[i] - s) / A[i * D + i];
            __syncwarp();
        }
        for (int d = lane; d < D; d += 32) out[d] = (float)(bvec[d] - (d == 0 ? (double)m.prior_offset : 0.0));
    }
}

extern "C" cudaError_t vbk_ivector(const IvecArgs *a, cudaStream_t s) {
    if (a->num_lanes <= 0) return cudaSuccess;
    if (a->m.num_gauss > 1024 || a->m.ivec_dim > kIvThreads || a->m.feat_dim > kIvThreads || a->m.num_gselect > kMaxGselect ||
        a->m.splice_dim != 7 * a->m.feat_dim)
        return cudaErrorInvalidValue;
    int smem = ivec_smem_bytes(a->m.num_gauss, a->m.feat_dim, a->m.ivec_dim);
    static int configured[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && configured[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(ivector_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured[dev] = smem;
    }
    ivector_kernel<<<a->num_lanes, kIvThreads, smem, s>>>(*a);
    return cudaGetLastError();
}

// generic ring-row copy used by the debug capture taps
__global__ void copy_rows_kernel(NodeDesc node, int ch, int t_begin, int n_rows, float *dst) {
    int total = n_rows * node.dim;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i / node.dim, c = i % node.dim;
        dst[i] = ring_row(node, ch, t_begin + r * node.step)[c];
    }
}
extern "C" cudaError_t vbk_copy_rows(NodeDesc node, int ch, int t_begin, int n_rows, float *dst, cudaStream_t s) {
    if (n_rows <= 0) return cudaSuccess;
    int total = n_rows * node.dim;
    copy_rows_kernel<<<(total + 255) / 256, 256, 0, s>>>(node, ch, t_begin, n_rows, dst);
    return cudaGetLastError();
}

}  // namespace vb
