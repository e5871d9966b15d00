// kernels_nnet.cu — K2: TDNN-F forward for all lanes of a step.
//
// Replaces BatchedStaticNnet3::RunBatch + nnet3 components behind the reference's batch pipeline
// [REF src/batch_model.cc:39-49,81-91]; architecture [REF training/local/chain/run_tdnn.sh:98-129].
//
// Streaming without context recompute: every layer output ("node") lives in a per-channel ring indexed by
// absolute time, so a layer advances exactly by the frames that became computable in this step and the
// time-spliced input of a TDNN layer ([x(t-s), x(t)] etc.) is gathered directly from the producer's ring.
//   * nnet_plan_kernel: per node and lane, which rows are new this step + packed row offsets over lanes
//   * gemm_fp32_kernel: fp32 FFMA tile GEMM with the gather in the A-tile load and bias/ReLU/batchnorm/
//     bypass fused in the epilogue (bring-up + cross-check path)
//   * gemm_tc (kernels_nnet_tc.cu): the tcgen05/TMEM version of the same contract (product path)
#include "vb_kernels.h"

namespace vb {

__device__ __forceinline__ int ring_slot(const NodeDesc &n, int t) { return ((t - n.t_start) / n.step) & (n.ring - 1); }

// one block per node, one thread per lane
__global__ void __launch_bounds__(1024) nnet_plan_kernel(NnetPlanArgs a) {
    __shared__ int s_scan[1024];
    const int n = blockIdx.x, l = threadIdx.x;
    const NodeDesc nd = a.nodes[n];
    int rows = 0, t0 = 0;
    if (l < a.num_lanes) {
        const LaneDesc ln = a.lanes[l];
        int *endp = a.node_end + (size_t)ln.channel * kMaxNodes + n;
        int prev = ln.first ? nd.t_start : *endp;
        int avail = ln.in_end_after - nd.cum_right;  // exclusive
        if (ln.in_end_after <= ln.in_end_before && !ln.first) avail = prev;
        if (avail > prev) rows = (avail - prev + nd.step - 1) / nd.step;
        t0 = prev;
        *endp = prev + rows * nd.step;
        a.table[(size_t)n * a.max_lanes + l] = NodeLane{t0, rows};
    }
    s_scan[l] = rows;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int v = l >= o ? s_scan[l - o] : 0;
        __syncthreads();
        s_scan[l] += v;
        __syncthreads();
    }
    if (l < a.num_lanes) a.rowoff[(size_t)n * (a.max_lanes + 1) + l + 1] = s_scan[l];
    if (l == 0) a.rowoff[(size_t)n * (a.max_lanes + 1)] = 0;
    if (l < a.num_lanes && a.rows) {
        const int ch = a.lanes[l].channel, base = s_scan[l] - rows;
        int2 *dst = a.rows + (size_t)n * a.rows_cap;
        for (int i = 0; i < rows && base + i < a.rows_cap; i++) dst[base + i] = make_int2(ch, t0 + i * nd.step);
    }
}

extern "C" cudaError_t vbk_nnet_plan(const NnetPlanArgs *a, cudaStream_t s) {
    if (a->num_lanes <= 0) return cudaSuccess;
    if (a->num_lanes > 1024) return cudaErrorInvalidValue;
    nnet_plan_kernel<<<a->num_nodes, 1024, 0, s>>>(*a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256) gemm_fp32_kernel(GemmArgs a) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ int s_ch[BM], s_t[BM];
    const int total = a.rowoff[a.num_lanes];
    const int row0 = blockIdx.x * BM;
    if (row0 >= total) return;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const OpDesc &op = a.op;
    if (tid < BM) {
        int r = row0 + tid, ch = -1, t = 0;
        if (r < total) {
            int lo = 0, hi = a.num_lanes;  // last lane with rowoff[lane] <= r
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (a.rowoff[mid] <= r) lo = mid; else hi = mid;
            }
            ch = a.lanes[lo].channel;
            t = a.table[lo].t_begin + (r - a.rowoff[lo]) * a.out.step;
        }
        s_ch[tid] = ch;
        s_t[tid] = t;
    }
    __syncthreads();
    const int in_dim = a.in.dim;
    const int spliced = op.n_off * in_dim;
    const int lr = tid >> 2, lk = (tid & 3) * 4;  // A/B tile load coordinates: row lr (0..63), k offset lk
    const int tx = tid & 15, ty = tid >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < op.K; k0 += BK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
        const int k = k0 + lk;
        if (k < op.K) {
            int ch = s_ch[lr];
            if (ch >= 0) {
                if (k < spliced) {
                    int j = k / in_dim, c = k - j * in_dim;
                    int t = s_t[lr] + op.offs[j];
                    av = *reinterpret_cast<const float4 *>(a.in.buf + ((size_t)ch * a.in.ring + ring_slot(a.in, t)) * in_dim + c);
                } else {
                    av = *reinterpret_cast<const float4 *>(a.ivec + (size_t)ch * a.ivec_dim + (k - spliced));
                }
            }
            int n = n0 + lr;
            if (n < op.N) bv = *reinterpret_cast<const float4 *>(op.W + (size_t)n * op.K + k);
        }
        As[lk + 0][lr] = av.x; As[lk + 1][lr] = av.y; As[lk + 2][lr] = av.z; As[lk + 3][lr] = av.w;
        Bs[lk + 0][lr] = bv.x; Bs[lk + 1][lr] = bv.y; Bs[lk + 2][lr] = bv.z; Bs[lk + 3][lr] = bv.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; kk++) {
            float4 x = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            float4 w = *reinterpret_cast<const float4 *>(&Bs[kk][tx * 4]);
            const float xa[4] = {x.x, x.y, x.z, x.w}, wa[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(xa[i], wa[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int r = ty * 4 + i, ch = s_ch[r];
        if (ch < 0) continue;
        int t = s_t[r];
        float *orow = a.out.buf + ((size_t)ch * a.out.ring + ring_slot(a.out, t)) * a.out.dim;
        const float *brow = op.byp_node >= 0 ? a.byp.buf + ((size_t)ch * a.byp.ring + ring_slot(a.byp, t)) * a.byp.dim : nullptr;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int n = n0 + tx * 4 + j;
            if (n >= op.N) continue;
            float z = acc[i][j];
            if (op.bias) z += op.bias[n];
            if (op.relu) z = fmaxf(z, 0.f);
            if (op.has_bn) z = fmaf(z, op.bn_scale[n], op.bn_offset[n]);
            if (brow) z = fmaf(op.bypass_scale, brow[n], z);
            orow[n] = z;
        }
    }
}

extern "C" cudaError_t vbk_gemm_fp32(const GemmArgs *a, cudaStream_t s) {
    if (a->num_lanes <= 0 || a->max_rows <= 0) return cudaSuccess;
    dim3 grid((a->max_rows + BM - 1) / BM, (a->op.N + BN - 1) / BN);
    gemm_fp32_kernel<<<grid, 256, 0, s>>>(*a);
    return cudaGetLastError();
}

__global__ void split_tf32_kernel(const float *w, float *hi, float *lo, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float v = w[i];
        float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        hi[i] = h;
        lo[i] = v - h;
    }
}
extern "C" cudaError_t vbk_split_tf32(const float *w, float *hi, float *lo, long long n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    split_tf32_kernel<<<(int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, s>>>(w, hi, lo, n);
    return cudaGetLastError();
}

}  // namespace vb
