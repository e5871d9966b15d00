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

constexpr int kFeatThreads = 256;
constexpr int kFeatHalfWarps = kFeatThreads / 16;  // a frame is owned by a half-warp: 16 lanes x 16 complex points in registers
constexpr int kFftTile = 16 * 17;                  // float2 per half-warp: the 16 x 16 transpose tile, padded rows
constexpr int kMelStride = kMelMaxLen + 4;         // filter rows staggered over the banks

// dynamic smem layout: fft[kFeatHalfWarps][kFftTile] float2 | logmel[kFeatHalfWarps][40] | window[400] | tw[256] float2 |
//                      dct_t[1600] | lifter[40] | mel_w[40][kMelStride] | mel_start[40] | mel_len[40] | wave[kCarryMax + spc] int16
extern "C" int vbk_feat_smem_bytes(int spc) {
    return (int)sizeof(float) * (kFeatHalfWarps * kFftTile * 2 + kFeatHalfWarps * 40 + 400 + 512 + 1600 + 40 + 40 * kMelStride + 80) +
           (int)sizeof(int16_t) * ((kCarryMax + spc + 7) & ~7);
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// forward 4-point DFT in place: (a, b, c, d) = x[0..3] -> X[0..3]
__device__ __forceinline__ void dft4(float2 &a, float2 &b, float2 &c, float2 &d) {
    const float2 s0 = make_float2(a.x + c.x, a.y + c.y), d0 = make_float2(a.x - c.x, a.y - c.y);
    const float2 s1 = make_float2(b.x + d.x, b.y + d.y), d1 = make_float2(b.x - d.x, b.y - d.y);
    a = make_float2(s0.x + s1.x, s0.y + s1.y);
    c = make_float2(s0.x - s1.x, s0.y - s1.y);
    b = make_float2(d0.x + d1.y, d0.y - d1.x);  // d0 - i d1
    d = make_float2(d0.x - d1.y, d0.y + d1.x);  // d0 + i d1
}

// forward 16-point DFT in registers (4 x 4 Cooley-Tukey).  In: v[n]; out: X[k1 + 4 k2] in v[4 k1 + k2] (fft16_at(k) gives the slot of X[k]).
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
#pragma unroll
    for (int n2 = 0; n2 < 4; n2++) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);  // v[4 k1 + n2] = sum_n1 x[4 n1 + n2] W4^(n1 k1)
    constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, h = 0.70710678118654752f;
    v[5] = cmul(v[5], make_float2(c1, -s1));    // W16^1
    v[6] = cmul(v[6], make_float2(h, -h));      // W16^2
    v[7] = cmul(v[7], make_float2(s1, -c1));    // W16^3
    v[9] = cmul(v[9], make_float2(h, -h));      // W16^2
    v[10] = make_float2(v[10].y, -v[10].x);     // W16^4 = -i
    v[11] = cmul(v[11], make_float2(-h, -h));   // W16^6
    v[13] = cmul(v[13], make_float2(s1, -c1));  // W16^3
    v[14] = cmul(v[14], make_float2(-h, -h));   // W16^6
    v[15] = cmul(v[15], make_float2(-c1, s1));  // W16^9
#pragma unroll
    for (int k1 = 0; k1 < 4; k1++) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}
__device__ __forceinline__ constexpr int fft16_at(int k) { return 4 * (k & 3) + (k >> 2); }

// K1.  One CTA per lane; the lane's samples (carried tail + new chunk) are staged once as int16.  A half-warp owns a frame: the 512
// real points are transformed as a 256-point complex FFT of (even, odd) pairs, 16 x 16: every lane runs a 16-point DFT on
// registers, the tile is transposed through shared memory once, a second 16-point DFT, then the real-input split against
// Z[256 - k] — two shared-memory passes per frame instead of nine radix-2 passes.  Mel bins: four lanes per filter; DCT: lane per
// coefficient.  Rows go straight into the network's input ring.
__global__ void __launch_bounds__(kFeatThreads) mfcc_kernel(FeatArgs a) {
    extern __shared__ __align__(16) float sm[];
    const LaneDesc ln = a.lanes[blockIdx.x];
    float2 *fft = reinterpret_cast<float2 *>(sm);
    float *logmel = sm + kFeatHalfWarps * kFftTile * 2;
    float *window = logmel + kFeatHalfWarps * 40;
    float2 *tw = reinterpret_cast<float2 *>(window + 400);
    float *dct_t = window + 400 + 512;
    float *lifter = dct_t + 1600;
    float *mel_w = lifter + 40;
    int *mel_start = reinterpret_cast<int *>(mel_w + 40 * kMelStride);
    int *mel_len = mel_start + 40;
    int16_t *wave = reinterpret_cast<int16_t *>(mel_len + 40);
    const int tid = threadIdx.x;
    for (int i = tid; i < 400; i += kFeatThreads) window[i] = a.tab.window[i];
    for (int i = tid; i < 256; i += kFeatThreads) tw[i] = reinterpret_cast<const float2 *>(a.tab.twiddle)[i];
    for (int i = tid; i < 1600; i += kFeatThreads) dct_t[i] = a.tab.dct_t[i];
    for (int i = tid; i < 40 * kMelMaxLen; i += kFeatThreads) mel_w[(i / kMelMaxLen) * kMelStride + i % kMelMaxLen] = a.tab.mel_w[i];
    if (tid < 40) {
        lifter[tid] = a.tab.lifter[tid];
        mel_start[tid] = a.tab.mel_start[tid];
        mel_len[tid] = a.tab.mel_len[tid];
    }
    // stage samples: carried tail then the new chunk (int16, values unscaled [REF src/batch_recognizer.cc:153-155])
    int16_t *carry = a.carry + (size_t)ln.channel * kCarryMax;
    for (int i = tid; i < ln.carry; i += kFeatThreads) wave[i] = carry[i];
    {
        const int16_t *src = a.staging + (size_t)ln.src_row * a.src_stride + ln.src_off;
        // 16-byte loads when the row start is aligned, stored as 32-bit words when the carried tail has an even length
        const int n8 = ((reinterpret_cast<uintptr_t>(src) & 15) || (ln.carry & 1)) ? 0 : ln.n_samples >> 3;
        const int4 *src4 = reinterpret_cast<const int4 *>(src);
        int *dst32 = reinterpret_cast<int *>(wave + ln.carry);
        for (int i = tid; i < n8; i += kFeatThreads) {
            const int4 v = __ldg(src4 + i);
            dst32[4 * i + 0] = v.x;
            dst32[4 * i + 1] = v.y;
            dst32[4 * i + 2] = v.z;
            dst32[4 * i + 3] = v.w;
        }
        for (int i = (n8 << 3) + tid; i < ln.n_samples; i += kFeatThreads) wave[ln.carry + i] = src[i];
    }
    __syncthreads();
    const int total = ln.carry + ln.n_samples;
    const int nf = ln.frames_after - ln.frames_before;
    const int hw = tid >> 4, L = tid & 15;
    float2 *tile = fft + hw * kFftTile;
    float *pw = reinterpret_cast<float *>(tile);  // the power spectrum reuses the tile
    float *lm = logmel + hw * 40;
    for (int f0 = 0; f0 < nf; f0 += kFeatHalfWarps) {  // warp-uniform trip count; a half-warp beyond the last frame redoes it and stores nothing
        const int f = min(f0 + hw, nf - 1);
        const bool live = f0 + hw < nf;
        const int16_t *w = wave + f * kFrameShift;
        int isum = 0;  // the sum of 400 int16 samples is exact in an int
        for (int i = L; i < kFrameLen / 2; i += 16) {
            const short2 p = reinterpret_cast<const short2 *>(w)[i];
            isum += (int)p.x + (int)p.y;
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) isum += __shfl_xor_sync(0xffffffffu, isum, o);
        const float mean = (float)isum / kFrameLen;
        // pass 1: lane L takes z[16 n1 + L] = (x[32 n1 + 2 L], x[32 n1 + 2 L + 1]) of the DC-free, pre-emphasised, windowed frame
        float2 v[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const int i = 32 * n1 + 2 * L;
            v[n1] = make_float2(0.f, 0.f);
            if (i < kFrameLen) {  // (n1 <= 12; 400 is even, so a pair is inside or outside as a whole)
                const short2 p = *reinterpret_cast<const short2 *>(w + i);
                const float x0 = (float)p.x - mean, x1 = (float)p.y - mean;
                const float xm = i > 0 ? (float)w[i - 1] - mean : x0;
                const float2 wn = *reinterpret_cast<const float2 *>(window + i);
                v[n1] = make_float2((x0 - 0.97f * xm) * wn.x, (x1 - 0.97f * x0) * wn.y);
            }
        }
        fft16(v);
        // twiddle W256^(L k1) = W512^(2 L k1) (table of 256: the upper half by sign), transpose
#pragma unroll
        for (int k1 = 0; k1 < 16; k1++) {
            float2 x = v[fft16_at(k1)];
            if (k1 > 0) {
                const int j = 2 * L * k1;
                float2 t = tw[j & 255];
                if (j & 256) t = make_float2(-t.x, -t.y);
                x = cmul(x, t);
            }
            tile[k1 * 17 + L] = x;
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++) v[n2] = tile[L * 17 + n2];
        __syncwarp();
        fft16(v);  // Z[L + 16 k2] in v[fft16_at(k2)]
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) tile[L + 16 * k2] = v[fft16_at(k2)];
        __syncwarp();
        // real-input split: X[k] = (Z[k] + conj Z[256-k]) / 2 - i / 2 W512^k (Z[k] - conj Z[256-k]); power spectrum of bins 0..255
        float pk[16];
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const int k = L + 16 * k2;
            const float2 zk = v[fft16_at(k2)], zn = tile[(256 - k) & 255];
            const float2 d = make_float2(zk.x - zn.x, zk.y + zn.y);
            const float2 wd = cmul(tw[k], d);
            const float xr = 0.5f * ((zk.x + zn.x) + wd.y), xi = 0.5f * ((zk.y - zn.y) - wd.x);
            pk[k2] = xr * xr + xi * xi;
        }
        __syncwarp();
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) pw[L + 16 * k2] = pk[k2];
        __syncwarp();
        // mel energies: four lanes per filter (taps sub, sub + 4, ..), ten rounds of four filters
        {
            const int sub = L & 3, grp = L >> 2;
            for (int r = 0; r < kNumMel / 4; r += 2) {  // two rounds at a time: their sums are independent chains
                const int j0 = 4 * r + grp, j1 = j0 + 4;
                const int st0 = mel_start[j0], n0 = mel_len[j0], st1 = mel_start[j1], n1 = mel_len[j1];
                const float *mw0 = mel_w + j0 * kMelStride, *mw1 = mel_w + j1 * kMelStride;
                float e0 = 0.f, e1 = 0.f;
                const int nmax = max(n0, n1);
                for (int i = sub; i < nmax; i += 4) {
                    if (i < n0) e0 = fmaf(mw0[i], pw[st0 + i], e0);
                    if (i < n1) e1 = fmaf(mw1[i], pw[st1 + i], e1);
                }
                e0 += __shfl_xor_sync(0xffffffffu, e0, 1);
                e1 += __shfl_xor_sync(0xffffffffu, e1, 1);
                e0 += __shfl_xor_sync(0xffffffffu, e0, 2);
                e1 += __shfl_xor_sync(0xffffffffu, e1, 2);
                if (sub == 0) {
                    lm[j0] = e0;
                    lm[j1] = e1;
                }
            }
        }
        __syncwarp();
        for (int j = L; j < kNumMel; j += 16) lm[j] = logf(fmaxf(lm[j], FLT_EPSILON));
        __syncwarp();
        if (live) {
            float *out = ring_row(a.in_node, ln.channel, ln.frames_before + f);
            for (int k = L; k < kNumCeps; k += 16) {
                float c0 = 0.f, c1 = 0.f;  // (two chains: even and odd mel bins)
#pragma unroll 10
                for (int j = 0; j < kNumMel; j += 2) {
                    c0 = fmaf(dct_t[j * 40 + k], lm[j], c0);
                    c1 = fmaf(dct_t[(j + 1) * 40 + k], lm[j + 1], c1);
                }
                out[k] = (c0 + c1) * lifter[k];
            }
        }
        __syncwarp();
    }
    __syncthreads();
    // new carry = samples from the start of the next frame
    const int next_start = nf > 0 ? nf * kFrameShift : 0;
    const int keep = (ln.frames_after > 0 || total >= kFrameLen) ? total - next_start : total;
    if (!ln.last)
        for (int i = tid; i < keep && i < kCarryMax; i += kFeatThreads) carry[i] = wave[next_start + i];
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
// i-vector statistics (double accumulation), per-chunk Cholesky solve — three launches per step:
//   ivector_cmn_kernel   : CTA per lane, thread per dimension: the sliding-window recursion over the chunk's new frames
//   ivector_post_kernel  : CTA per (lane, tile of 16 frames): LDA of the raw and the normalised splice (thread = output dim x
//                          three frames, 16-byte operand loads), UBM log-likelihoods (thread = 2 Gaussians x 16 frames in 32
//                          register accumulators), warp per frame: top-N, posteriors -> {Gaussian, weight} lists + raw LDA rows
//   ivector_stats_kernel : CTA per lane: per-Gaussian occupancies as exact fixed-point sums (integer atomics: the order of the
//                          additions cannot matter) and first-order sums per selected Gaussian (frames ascending), projected
//                          once per Gaussian; quadratic term from the packed lower triangles; right-looking Cholesky in
//                          double in shared memory, substitution by one warp
// The frame-parallel part has lanes x tiles CTAs of work instead of one CTA per lane looping over its tiles.
// =====================================================================================================
constexpr int kIvThreads = 256;
constexpr int kIvWarps = kIvThreads / 32;
constexpr int kIvTB = 16;           // frames per tile
constexpr int kMaxGselect = 8;
constexpr int kIvGroup = 64;        // frames per pass of the statistics kernel (one bit each in a Gaussian's frame mask)
constexpr int kIvGaussBatch = 64;   // selected Gaussians whose first-order sums are staged at a time
constexpr float kIvFix = 1099511627776.f;  // 2^40: posterior weights as fixed point (exact for weights >= 2^-17)

extern "C" int vbk_ivector_frames_cap(int samples_per_chunk) { return ((samples_per_chunk / kFrameShift + 8 + kIvTB - 1) / kIvTB) * kIvTB; }

constexpr int kCmnThreads = 256;
constexpr int kCmnTile = 64;  // frames per pass (a chunk has 51)
__global__ void __launch_bounds__(kCmnThreads) ivector_cmn_kernel(IvecArgs a) {
    const LaneDesc ln = a.lanes[blockIdx.x];
    const IvecModel &m = a.m;
    constexpr int F = kNumCeps;
    const int tid = threadIdx.x, ch = ln.channel;
    __shared__ float s_x[kCmnTile][F], s_xo[kCmnTile][F];
    __shared__ double s_sum[kCmnTile][F];
    double *cm_sum = a.st.cmvn_sum + (size_t)ch * F;
    float *nring = a.st.norm_ring + (size_t)ch * kNormRing * F;
    double run = 0.0;  // (threads < F) the running window sum of this dimension
    if (tid < F && !ln.first) run = cm_sum[tid];
    for (int t0 = ln.frames_before; t0 < ln.frames_after; t0 += kCmnTile) {
        const int n = min(kCmnTile, ln.frames_after - t0);
        // the tile's rows (and the rows leaving the window), all threads
        for (int i = tid; i < n * F; i += kCmnThreads) {
            const int j = i / F, d = i - j * F, t = t0 + j;
            s_x[j][d] = ring_row(a.in_node, ch, t)[d];
            s_xo[j][d] = t >= m.cmn_window ? ring_row(a.in_node, ch, t - m.cmn_window)[d] : 0.f;
        }
        __syncthreads();
        // the serial part: the running sums, thread per dimension (two additions per frame)
        if (tid < F) {
            for (int j = 0; j < n; j++) {
                run += (double)s_x[j][tid];
                if (t0 + j >= m.cmn_window) run -= (double)s_xo[j][tid];
                s_sum[j][tid] = run;
            }
        }
        __syncthreads();
        // everything else is per (frame, dimension): smoothing with the global statistics, the mean, the normalised value
        for (int i = tid; i < n * F; i += kCmnThreads) {
            const int j = i / F, d = i - j * F, t = t0 + j;
            const double nn = t + 1 < m.cmn_window ? t + 1 : m.cmn_window;
            const double fg = nn < m.cmn_window ? fmin((double)m.cmn_window - nn, (double)m.global_frames) : 0.0;
            double tot = s_sum[j][d];
            if (fg > 0.0) tot += fg * (m.gcmvn_sum[d] / m.gcmvn_count);
            nring[(t & (kNormRing - 1)) * F + d] = (float)((double)s_x[j][d] - tot / (nn + fg));
        }
        __syncthreads();
    }
    if (tid < F) cm_sum[tid] = run;
}

struct IvPostLayout {  // shared-memory carve-up of the frame kernel, in floats
    int win_raw, win_nrm, fnT, fn2T, ll, total_floats;
};
__host__ __device__ inline IvPostLayout iv_post_layout(int G, int F) {
    IvPostLayout L;
    int o = 0;
    L.win_raw = o; o += (kIvTB + 6) * F;
    L.win_nrm = o; o += (kIvTB + 6) * F;
    o = (o + 3) & ~3;
    L.fnT = o; o += F * kIvTB;
    L.fn2T = o; o += F * kIvTB;
    L.ll = o; o += kIvTB * G;
    L.total_floats = o;
    return L;
}

// top-N of one frame's G log-likelihoods by repeated warp arg-max (ties: the smaller index), PER values per lane
template <int PER>
__device__ __forceinline__ void iv_select(const float *llrow, int G, int ng, int lane, float *best_v, int *best_g) {
    float llv[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
        const int g = lane + 32 * j;
        llv[j] = g < G ? llrow[g] : -FLT_MAX;
    }
#pragma unroll
    for (int r = 0; r < kMaxGselect; r++) {
        best_v[r] = -FLT_MAX;
        best_g[r] = 0;
        if (r >= ng) continue;
        float bv = -FLT_MAX;
        int bg = 0x7fffffff;
#pragma unroll
        for (int j = 0; j < PER; j++)
            if (llv[j] > bv) { bv = llv[j]; bg = lane + 32 * j; }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int og = __shfl_xor_sync(0xffffffffu, bg, o);
            if (ov > bv || (ov == bv && og < bg)) { bv = ov; bg = og; }
        }
        best_v[r] = bv;
        best_g[r] = bg;
#pragma unroll
        for (int j = 0; j < PER; j++)
            if (bg == lane + 32 * j) llv[j] = -FLT_MAX;
    }
}

__global__ void __launch_bounds__(kIvThreads, 2) ivector_post_kernel(IvecArgs a) {
    extern __shared__ __align__(16) float smf[];
    const LaneDesc ln = a.lanes[blockIdx.x];
    const IvecModel &m = a.m;
    constexpr int F = kNumCeps;  // (the loader accepts no other feature dimension: constant strides keep the address arithmetic out of the loops)
    const int G = m.num_gauss, S = m.splice_dim;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const IvPostLayout L = iv_post_layout(G, F);
    float *win_raw = smf + L.win_raw, *win_nrm = smf + L.win_nrm, *fnT = smf + L.fnT, *fn2T = smf + L.fn2T, *ll = smf + L.ll;
    const int ch = ln.channel;
    const float *nring = a.st.norm_ring + (size_t)ch * kNormRing * F;
    const int last_avail = ln.frames_after - 1;
    const int ng = min(m.num_gselect, G);
    const int n_iv = min(ln.iv_end_after - ln.iv_end_before, a.frames_cap);
    for (int f0 = blockIdx.y * kIvTB; f0 < n_iv; f0 += gridDim.y * kIvTB) {
        const int tb = ln.iv_end_before + f0;
        const int nb = min(kIvTB, n_iv - f0);
        __syncthreads();
        for (int i = tid; i < (nb + 6) * F; i += kIvThreads) {
            const int r = i / F, d = i % F;
            const int tt = min(max(tb + r - 3, 0), last_avail);
            win_raw[i] = ring_row(a.in_node, ch, tt)[d];
            win_nrm[i] = nring[(tt & (kNormRing - 1)) * F + d];
        }
        __syncthreads();
        // LDA: out[t][d] = offset[d] + sum_k lda[k][d] * splice(t)[k]; splice(t)[o*F + kd] = win[t + o][kd].  Thread = (d, frame
        // group): frames tg, tg + ngF, tg + 2 ngF; four weights and the six 16-byte operand pieces feed 24 multiply-adds
        const int ngF = kIvThreads / F;
        if (tid < ngF * F) {
            const int d = tid % F, tg = tid / F;
            for (int t0 = tg; t0 < nb; t0 += 3 * ngF) {
                const int t1 = t0 + ngF, t2 = t0 + 2 * ngF;
                const int r1 = min(t1, nb - 1), r2 = min(t2, nb - 1);  // rows beyond the tile repeat the last one (discarded)
                const float off = __ldg(m.lda_t + (size_t)S * F + d);
                float u0 = off, u1 = off, u2 = off, n0 = off, n1 = off, n2 = off;
                for (int o = 0; o < 7; o++) {
                    const float *w = m.lda_t + o * F * F + d;
                    const float4 *xr0 = reinterpret_cast<const float4 *>(win_raw + (t0 + o) * F), *xr1 = reinterpret_cast<const float4 *>(win_raw + (r1 + o) * F),
                                 *xr2 = reinterpret_cast<const float4 *>(win_raw + (r2 + o) * F);
                    const float4 *xn0 = reinterpret_cast<const float4 *>(win_nrm + (t0 + o) * F), *xn1 = reinterpret_cast<const float4 *>(win_nrm + (r1 + o) * F),
                                 *xn2 = reinterpret_cast<const float4 *>(win_nrm + (r2 + o) * F);
#pragma unroll
                    for (int k4 = 0; k4 < F / 4; k4++) {
                        const float w0 = __ldg(w + (4 * k4) * F), w1 = __ldg(w + (4 * k4 + 1) * F), w2 = __ldg(w + (4 * k4 + 2) * F), w3 = __ldg(w + (4 * k4 + 3) * F);
                        const float4 a0 = xr0[k4], a1 = xr1[k4], a2 = xr2[k4], b0 = xn0[k4], b1 = xn1[k4], b2 = xn2[k4];
                        u0 = fmaf(w3, a0.w, fmaf(w2, a0.z, fmaf(w1, a0.y, fmaf(w0, a0.x, u0))));
                        u1 = fmaf(w3, a1.w, fmaf(w2, a1.z, fmaf(w1, a1.y, fmaf(w0, a1.x, u1))));
                        u2 = fmaf(w3, a2.w, fmaf(w2, a2.z, fmaf(w1, a2.y, fmaf(w0, a2.x, u2))));
                        n0 = fmaf(w3, b0.w, fmaf(w2, b0.z, fmaf(w1, b0.y, fmaf(w0, b0.x, n0))));
                        n1 = fmaf(w3, b1.w, fmaf(w2, b1.z, fmaf(w1, b1.y, fmaf(w0, b1.x, n1))));
                        n2 = fmaf(w3, b2.w, fmaf(w2, b2.z, fmaf(w1, b2.y, fmaf(w0, b2.x, n2))));
                    }
                }
                float *fu = a.fu + ((size_t)blockIdx.x * a.frames_cap + f0) * F;
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
        // diag-UBM log-likelihoods of the tile: ll[t][g] = gconst[g] + sum_d (mi[d][g] x - 0.5 iv[d][g] x^2)   (niv_t = -0.5 iv)
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
            const float2 *u0 = m.ubm_t + g0, *u1 = m.ubm_t + g1c;  // {mean * inv_var, -0.5 inv_var} of a Gaussian: one 8-byte load
#pragma unroll 4
            for (int d = 0; d < F; d++) {
                const float2 p0 = __ldg(u0), p1 = __ldg(u1);
                u0 += G;
                u1 += G;
                const float mi0 = p0.x, iv0 = p0.y, mi1 = p1.x, iv1 = p1.y;
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
        // per frame (warp): top-N Gaussians, posteriors (min_post pruning, posterior_scale)
        for (int t = warp; t < nb; t += kIvWarps) {
            float best_v[kMaxGselect];
            int best_g[kMaxGselect];
            if (G <= 512) iv_select<16>(ll + t * G, G, ng, lane, best_v, best_g);
            else iv_select<32>(ll + t * G, G, ng, lane, best_v, best_g);
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
            const size_t row = ((size_t)blockIdx.x * a.frames_cap + f0 + t) * kMaxGselect;
#pragma unroll
            for (int r = 0; r < kMaxGselect; r++) {
                const float w = r < ng ? post[r] / kept * m.posterior_scale : 0.f;
                if (lane == r) {
                    a.sel_g[row + r] = r < ng ? best_g[r] : -1;
                    a.sel_w[row + r] = w;
                }
            }
        }
    }
}

__host__ __device__ __forceinline__ int tri(int i) { return i * (i + 1) / 2; }

struct IvStatLayout {  // shared-memory carve-up of the statistics kernel, in bytes
    int gsum, gmask, A, bvec, gam, lin_part, fu, vbuf, sel_g, sel_w, glist, total;
};
__host__ __device__ inline IvStatLayout iv_stat_layout(int G, int F, int D) {
    IvStatLayout L;
    int o = 0;
    L.gsum = o; o += G * 8;
    L.gmask = o; o += G * 8;
    L.A = o; o += tri(D) * 8;
    L.bvec = o; o += D * 8;
    L.gam = o; o += G * 8;  // occupancies of the selected Gaussians, in list order
    const int ngD = kIvThreads / (D / 4);
    L.lin_part = o; o += ngD * D * 8;
    L.fu = o; o += kIvGroup * F * 4;
    L.vbuf = o; o += kIvGaussBatch * F * 4;
    L.sel_g = o; o += kIvGroup * kMaxGselect * 4;
    L.sel_w = o; o += kIvGroup * kMaxGselect * 4;
    L.glist = o; o += (G + 8) * 4;
    L.total = (o + 15) & ~15;
    return L;
}

// ascending list of the Gaussians with a non-zero word in v[0..G): block-wide stable compaction.  Returns the count.
__device__ int iv_compact(const unsigned long long *v, int G, int *glist, int *s_warp, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    int base = 0;
    for (int g0 = 0; g0 < G; g0 += kIvThreads) {
        const int g = g0 + tid;
        const bool f = g < G && v[g] != 0ull;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int pre = 0, tot = 0;
        for (int w = 0; w < kIvWarps; w++) {
            const int c = s_warp[w];
            pre += w < warp ? c : 0;
            tot += c;
        }
        if (f) glist[base + pre + __popc(bal & ((1u << lane) - 1u))] = g;
        base += tot;
        __syncthreads();
    }
    return base;
}

__global__ void __launch_bounds__(kIvThreads) ivector_stats_kernel(IvecArgs a) {
    extern __shared__ __align__(16) unsigned char smb[];
    const LaneDesc ln = a.lanes[blockIdx.x];
    const IvecModel &m = a.m;
    const int F = m.feat_dim, D = m.ivec_dim, G = m.num_gauss, NT = tri(D);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const IvStatLayout L = iv_stat_layout(G, F, D);
    unsigned long long *gsum = reinterpret_cast<unsigned long long *>(smb + L.gsum), *gmask = reinterpret_cast<unsigned long long *>(smb + L.gmask);
    double *A = reinterpret_cast<double *>(smb + L.A), *bvec = reinterpret_cast<double *>(smb + L.bvec), *gam = reinterpret_cast<double *>(smb + L.gam);
    double *lin_part = reinterpret_cast<double *>(smb + L.lin_part);
    float *fu = reinterpret_cast<float *>(smb + L.fu), *vbuf = reinterpret_cast<float *>(smb + L.vbuf), *sel_w = reinterpret_cast<float *>(smb + L.sel_w);
    int *sel_g = reinterpret_cast<int *>(smb + L.sel_g), *glist = reinterpret_cast<int *>(smb + L.glist);
    __shared__ int s_warp[kIvWarps];
    const int ch = ln.channel;
    double *lin = a.st.lin + (size_t)ch * D;
    double *quad = a.st.quad + (size_t)ch * NT;  // packed lower triangle
    const int ng = min(m.num_gselect, G);
    const int n_iv = min(ln.iv_end_after - ln.iv_end_before, a.frames_cap);
    const int D4 = D / 4, ngD = kIvThreads / D4;  // (D is a multiple of 4: the loader pads the extractor)
    for (int g = tid; g < G; g += kIvThreads) gsum[g] = 0ull;
    double lacc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int fb = 0; fb < n_iv; fb += kIvGroup) {
        const int nbf = min(kIvGroup, n_iv - fb);
        __syncthreads();
        for (int g = tid; g < G; g += kIvThreads) gmask[g] = 0ull;
        {
            const size_t row0 = (size_t)blockIdx.x * a.frames_cap + fb;
            for (int i = tid; i < nbf * kMaxGselect; i += kIvThreads) {
                sel_g[i] = a.sel_g[row0 * kMaxGselect + i];
                sel_w[i] = a.sel_w[row0 * kMaxGselect + i];
            }
            for (int i = tid; i < nbf * F; i += kIvThreads) fu[i] = a.fu[row0 * F + i];
        }
        __syncthreads();
        for (int e = tid; e < nbf * ng; e += kIvThreads) {
            const int t = e / ng, r = e % ng;
            const float w = sel_w[t * kMaxGselect + r];
            if (w != 0.f) {
                const int g = sel_g[t * kMaxGselect + r];
                atomicAdd(gsum + g, __float2ull_rn(w * kIvFix));
                atomicOr(gmask + g, 1ull << t);
            }
        }
        __syncthreads();
        const int nl = iv_compact(gmask, G, glist, s_warp, tid);
        // first-order sums per selected Gaussian (its frames ascending), then one projection per Gaussian:
        // lin[d] += sum_a SiM[g][a][d] * (sum_t w[t][g] fu[t][a])
        for (int kb = 0; kb < nl; kb += kIvGaussBatch) {
            const int nk = min(kIvGaussBatch, nl - kb);
            for (int i = tid; i < nk * F; i += kIvThreads) {
                const int k = i / F, f = i % F;
                const int g = glist[kb + k];
                unsigned long long mk = gmask[g];
                float s = 0.f;
                while (mk) {
                    const int t = __ffsll((long long)mk) - 1;
                    mk &= mk - 1;
                    float w = 0.f;
                    for (int r = 0; r < ng; r++)
                        if (sel_g[t * kMaxGselect + r] == g) {
                            w = sel_w[t * kMaxGselect + r];
                            break;
                        }
                    s = fmaf(w, fu[t * F + f], s);
                }
                vbuf[i] = s;
            }
            __syncthreads();
            if (tid < ngD * D4) {
                const int d4 = tid % D4, pg = tid / D4;
                for (int k = pg; k < nk; k += ngD) {
                    const float4 *sm_g = reinterpret_cast<const float4 *>(m.sim + (size_t)glist[kb + k] * F * D) + d4;
                    const float *v = vbuf + k * F;
                    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
                    for (int q = 0; q < F; q++) {
                        const float4 c = __ldg(sm_g + (size_t)q * D4);
                        const float x = v[q];
                        s0.x = fmaf(c.x, x, s0.x);
                        s0.y = fmaf(c.y, x, s0.y);
                        s0.z = fmaf(c.z, x, s0.z);
                        s0.w = fmaf(c.w, x, s0.w);
                    }
                    lacc[0] += (double)s0.x;
                    lacc[1] += (double)s0.y;
                    lacc[2] += (double)s0.z;
                    lacc[3] += (double)s0.w;
                }
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (tid < ngD * D4) {
#pragma unroll
        for (int j = 0; j < 4; j++) lin_part[(tid / D4) * D + (tid % D4) * 4 + j] = lacc[j];
    }
    // ---- fold the chunk's statistics into the channel state (fixed summation order) ----
    const int nl = iv_compact(gsum, G, glist, s_warp, tid);
    for (int k = tid; k < nl; k += kIvThreads) gam[k] = (double)gsum[glist[k]] * (1.0 / (double)kIvFix);
    __syncthreads();
    double tw = 0.0;
    for (int k = 0; k < nl; k++) tw += gam[k];
    const double nf_old = ln.first ? 0.0 : a.st.num_frames[ch], nf_new = nf_old + tw;
    double change = 0.0;
    if (m.max_count > 0.f)
        change = fmax(nf_new, (double)m.max_count) / m.max_count - fmax(nf_old, (double)m.max_count) / m.max_count;
    for (int i = tid; i < NT; i += kIvThreads) {
        double q = 0.0;
        const float *u = m.U_tri + i;
        int k = 0;
        for (; k + 4 <= nl; k += 4) {  // four independent loads in flight, added in list order
            const float u0 = __ldg(u + (size_t)glist[k] * NT), u1 = __ldg(u + (size_t)glist[k + 1] * NT), u2 = __ldg(u + (size_t)glist[k + 2] * NT),
                        u3 = __ldg(u + (size_t)glist[k + 3] * NT);
            q += gam[k] * (double)u0;
            q += gam[k + 1] * (double)u1;
            q += gam[k + 2] * (double)u2;
            q += gam[k + 3] * (double)u3;
        }
        for (; k < nl; k++) q += gam[k] * (double)__ldg(u + (size_t)glist[k] * NT);
        // row r of the packed triangle: the largest r with r (r + 1) / 2 <= i; i is a diagonal entry iff i == tri(r) + r
        int r = (int)((sqrtf(8.f * (float)i + 1.f) - 1.f) * 0.5f);
        while (tri(r + 1) <= i) r++;
        while (tri(r) > i) r--;
        const bool diag = i == tri(r) + r;
        if (diag) q += change;
        const double prev = ln.first ? (diag ? 1.0 : 0.0) : quad[i];  // stream start: prior (OnlineIvectorEstimationStats ctor)
        const double v = prev + q;
        quad[i] = v;
        A[i] = v;
    }
    for (int d = tid; d < D; d += kIvThreads) {
        double s = 0.0;
        for (int pg = 0; pg < ngD; pg++) s += lin_part[pg * D + d];
        if (d == 0) s += (double)m.prior_offset * change;
        const double prev = ln.first ? (d == 0 ? (double)m.prior_offset : 0.0) : lin[d];
        const double v = prev + s;
        lin[d] = v;
        bvec[d] = v;
    }
    __syncthreads();
    if (tid == 0) a.st.num_frames[ch] = nf_new;
    // ---- solve quad * x = lin by Cholesky (what the reference's batch path does per chunk) ----
    float *out = a.st.ivec + (size_t)ch * D;
    if (nf_new <= 0.0) {
        for (int d = tid; d < D; d += kIvThreads) out[d] = 0.f;
        return;
    }
    for (int j = 0; j < D; j++) {  // right-looking: scale column j, then the rank-1 update of the trailing triangle (warp per row)
        const double rinv = rsqrt(fmax(A[tri(j) + j], 1e-300));
        for (int i = j + 1 + tid; i < D; i += kIvThreads) A[tri(i) + j] *= rinv;
        __syncthreads();
        if (tid == 0) A[tri(j) + j] = rinv;  // the diagonal keeps 1 / L_jj: the substitution multiplies (nobody reads it before)
        for (int i = j + 1 + warp; i < D; i += kIvWarps) {
            const double lij = A[tri(i) + j];
            for (int k = j + 1 + lane; k <= i; k += 32) A[tri(i) + k] -= lij * A[tri(k) + j];
        }
        __syncthreads();
    }
    if (warp == 0) {
        for (int i = 0; i < D; i++) {
            double s = 0.0;
            for (int k = lane; k < i; k += 32) s += A[tri(i) + k] * bvec[k];
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) bvec[i] = (bvec[i] - s) * A[tri(i) + i];
            __syncwarp();
        }
        for (int i = D - 1; i >= 0; i--) {
            double s = 0.0;
            for (int k = i + 1 + lane; k < D; k += 32) s += A[tri(k) + i] * bvec[k];
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) bvec[i] = (bvec[i] - s) * A[tri(i) + i];
            __syncwarp();
        }
        for (int d = lane; d < D; d += 32) out[d] = (float)(bvec[d] - (d == 0 ? (double)m.prior_offset : 0.0));
    }
}

extern "C" cudaError_t vbk_ivector(const IvecArgs *a, cudaStream_t s) {
    if (a->num_lanes <= 0) return cudaSuccess;
    if (a->m.num_gauss > 1024 || a->m.ivec_dim > kIvThreads || a->m.ivec_dim % 4 || a->m.feat_dim != kNumCeps ||
        a->m.num_gselect > kMaxGselect || a->m.splice_dim != 7 * a->m.feat_dim || !a->sel_g || !a->sel_w || !a->fu || a->frames_cap < kIvTB)
        return cudaErrorInvalidValue;
    const int smem_post = iv_post_layout(a->m.num_gauss, a->m.feat_dim).total_floats * 4;
    const int smem_stat = iv_stat_layout(a->m.num_gauss, a->m.feat_dim, a->m.ivec_dim).total;
    static int configured[16][2] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && (configured[dev][0] < smem_post || configured[dev][1] < smem_stat)) {
        cudaError_t e = cudaFuncSetAttribute(ivector_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_post);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(ivector_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_stat);
        if (e != cudaSuccess) return e;
        configured[dev][0] = smem_post;
        configured[dev][1] = smem_stat;
    }
    ivector_cmn_kernel<<<a->num_lanes, kCmnThreads, 0, s>>>(*a);
    ivector_post_kernel<<<dim3(a->num_lanes, a->frames_cap / kIvTB), kIvThreads, smem_post, s>>>(*a);
    ivector_stats_kernel<<<a->num_lanes, kIvThreads, smem_stat, s>>>(*a);
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
