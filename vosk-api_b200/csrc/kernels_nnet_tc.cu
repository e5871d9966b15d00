// kernels_nnet_tc.cu — K2 product path: TDNN-F layer GEMM on the 5th-generation tensor cores.
//
//   C[rows, N] = epilogue( A[rows, K] * W[N, K]^T ),  A gathered from the producer layer's time ring
//   (the TDNN splice [x(t+o0), x(t+o1), ...] (+ i-vector) is materialised only in shared memory).
//
// Precision: the reference runs this stage as fp32 SGEMM (tensor cores off, [REF src/vosk_api.cc:184-185]) and
// north_star asks for log-likelihoods within 1e-3, which plain TF32/BF16/FP16 inputs cannot hold through 30 chained
// GEMMs.  Each fp32 operand is therefore split x = hi + lo and the tile is accumulated as
// A_hi*W_hi + A_lo*W_hi + A_hi*W_lo in fp32 TMEM accumulators, dropping only the lo*lo term (~2^-22 relative):
//   tensor-cores=1 (default): hi = fp16(x), lo = fp16((x - hi) * 2^11) — x - hi is exact in fp32, the scaling keeps lo a
//     normal fp16 number and is undone (exactly) in the epilogue; kind::f16, K = 16 per MMA, 64 K-elements per 128-byte
//     operand row.  |x| is clamped to the fp16 range (65504), far above what batch-normalised activations reach.
//   tensor-cores=2: hi = x & 0xffffe000 (a TF32 value), lo = x - hi; kind::tf32, K = 8 per MMA, 32 K-elements per row.
// Measured on B200: the tensor core adds into the fp32 accumulator with truncation, so the error grows with the NUMBER
// of MMAs accumulated.  The two cross terms are ~2^-11 of the result, so they get their own TMEM accumulator (their
// truncation error is relative to their own size) and the epilogue adds them in fp32 round-to-nearest; the hi*hi k-steps
// are dealt round-robin over several accumulators.  The fp16 split needs half as many MMAs per K: with two accumulators it
// measures 4.4e-4 end to end on the small architecture (TF32 split with four: 7.1e-4) and, with four from K = 2048,
// 7.9e-4 on the large one (TF32 split: 1.6e-3), while the tile grows from 96 to 128-160 columns and the TDNN-F chain runs
// 26 % faster.  Tensor-pipe FLOPs are 3x the algorithmic ones; bench.py counts the algorithmic ones.
//
// Structure (one 128 x BN output tile per CTA, BN = min(N, 256)):
//   warps 0-7  A producers: coalesced 16-byte gathers from the ring rows -> hi/lo split in registers -> 128B-
//              swizzled K-major smem tiles (generic-proxy stores + fence.proxy.async + mbarrier arrive);
//              afterwards the same warps are the epilogue (warps w and w + 4 share the 32 TMEM lanes of quarter w % 4 and
//              take alternate 16-column chunks: tcgen05.ld, bias / ReLU / batchnorm / bypass fused, 64-byte row-segment
//              stores; measured bound: the TMEM read path, a per-warp shared-memory transpose for coalesced stores
//              changed nothing).
//   warp 8     TMA: cp.async.bulk.tensor of the W_hi / W_lo [BN x 32] boxes (SWIZZLE_128B) per K-block.
//   warp 9     allocates TMEM, issues tcgen05.mma.cta_group::1.kind::tf32 (3 per 8-wide k-step), commits the
//              smem stage back to the producers and finally the accumulator to the epilogue.
#include <cuda.h>
#include <cuda_fp16.h>

#include "vb_kernels.h"

namespace vb {

namespace {
constexpr int TM = 128;       // tile rows (UMMA_M)
constexpr int TK = 32;        // fp32 elements per K-block = one 128-byte swizzle row
constexpr int kTcThreads = 320;        // 8 producer warps (also the epilogue), 1 TMA warp, 1 MMA warp
constexpr int kProducerThreads = 256;
constexpr int kTmaWarp = kProducerThreads / 32, kMmaWarp = kTmaWarp + 1;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    long long spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (!done && ++spins > (1ll << 26)) __trap();  // a protocol bug must fail loudly, never hang the GPU
    }
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row (1024 B) swizzle atoms
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffff) >> 4);        // start address
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}

struct TcSmem {
    uint64_t full[4], empty[4], accum;
    uint32_t tmem_base;
    int row_ch[TM], row_t[TM];
    float bias[256], bn_scale[256], bn_offset[256];  // the tile's columns of the epilogue constants (BN <= 256)
};
}  // namespace

// per-phase cycle counters of the tile pipeline (profiling aid, VB_TC_PROF=1 all layers / 2 the K=192 N=512 affines /
// 3 the K>=1024 linears): tiles, prologue, producer loop, wait for the accumulator, epilogue, MMA warp: first operands
// ready, MMA warp: last commit (both counted from the end of the prologue), MMA warp: cycles waiting for operands / issuing MMAs
__device__ unsigned long long g_tc_prof[16];

struct alignas(64) TensorMapBlob {
    unsigned char b[128];
};

// F16 = false: operands split into TF32 hi/lo (32 K-elements per 128-byte row);  F16 = true: operands split into fp16 hi and
// 2^11-scaled fp16 lo (64 K-elements per 128-byte row, kind::f16, K = 16 per MMA) — half the shared-memory operand bytes
// per multiply-add, which is what paces the main loop.
template <bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
gemm_tc_kernel(GemmArgs a, const __grid_constant__ TensorMapBlob map_hi, const __grid_constant__ TensorMapBlob map_lo, int BN,
               int stages, int tmem_cols, int terms, int n_main, int prof) {
    const long long pt0 = prof ? clock64() : 0;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // dynamic smem: [stage][A_hi 16K | A_lo 16K | B_hi BN*128 | B_lo BN*128], 1024-byte aligned
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ TcSmem ts;
    const int total = a.rowoff[a.num_lanes];
    const int row0 = blockIdx.x * TM;
    if (row0 >= total) return;
    const int n0 = blockIdx.y * BN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const OpDesc &op = a.op;
    const uint32_t stage_bytes = 2u * TM * 128u + 2u * (uint32_t)BN * 128u;
    constexpr int TKE = F16 ? 64 : TK;  // K-elements per K-block (one 128-byte swizzle row)
    const int nkb = (op.K + TKE - 1) / TKE;

    if (tid < TM) {
        int r = row0 + tid, ch = -1, t = 0;
        if (r < total) {
            const int2 rc = a.rows[r];  // written by the plan kernel
            ch = rc.x;
            t = rc.y;
        }
        ts.row_ch[tid] = ch;
        ts.row_t[tid] = t;
    }
    for (int j = tid; j < BN; j += kTcThreads) {
        const int n = n0 + j;
        const bool in_n = n < op.N;
        ts.bias[j] = op.bias && in_n ? __ldg(op.bias + n) : 0.f;
        ts.bn_scale[j] = op.has_bn && in_n ? __ldg(op.bn_scale + n) : 1.f;
        ts.bn_offset[j] = op.has_bn && in_n ? __ldg(op.bn_offset + n) : 0.f;
    }
    if (tid == kProducerThreads) {
        for (int s = 0; s < stages; s++) {
            mbar_init(&ts.full[s], kProducerThreads + 1);
            mbar_init(&ts.empty[s], 1);
        }
        mbar_init(&ts.accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&ts.tmem_base)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = ts.tmem_base;
    const long long pt1 = prof ? clock64() : 0;

    if (warp < kTmaWarp) {
        // =========================== A producers ===========================
        // Thread (c, rsub) owns 16-byte chunk c of rows rsub, rsub+32, ... of the tile.  Row slots are resolved once; the
        // gathers of a K-block are issued back to back into registers and three K-blocks are kept in flight per thread
        // (48 KB of independent L2 requests per SM), so the loop is bandwidth- rather than latency-paced.
        const int in_dim = a.in.dim, spliced = op.n_off * in_dim;
        const int c = tid & 7;          // 16-byte chunk within the 128-byte row
        const int rsub = tid >> 3;      // 0..31: row within a group of 32
        constexpr int RG = kProducerThreads / 8;  // rows per group
        constexpr int RPT = TM / RG;    // rows per thread
        const float *row_base[RPT];     // ring base of the row's channel (nullptr = padding row)
        const float *iv_base[RPT];
        int row_slot[RPT];              // (t - t_start) / step of the row in the input ring
#pragma unroll
        for (int it = 0; it < RPT; it++) {
            const int r = it * RG + rsub;
            const int ch = ts.row_ch[r];
            row_base[it] = ch >= 0 ? a.in.buf + (size_t)ch * a.in.ring * in_dim : nullptr;
            iv_base[it] = ch >= 0 ? a.ivec + (size_t)ch * a.ivec_dim : nullptr;
            row_slot[it] = (ts.row_t[r] - a.in.t_start) / a.in.step;
        }
        const int ring_mask = a.in.ring - 1;
        if constexpr (F16) {
            // chunk c of a row = 8 K-elements = two 16-byte fp32 pieces -> one 16-byte piece of 8 halves in each operand tile
            auto gather16 = [&](int kb, float4 *v) {
#pragma unroll
                for (int pc = 0; pc < 2; pc++) {
                    const int k = kb * TKE + c * 8 + pc * 4;
                    const bool in_k = k < op.K, is_iv = k >= spliced;
                    int seg = 0, col = 0;
                    if (in_k && !is_iv) {
                        seg = k / in_dim;
                        col = k - seg * in_dim;
                    }
                    const int off_rows = in_k && !is_iv ? op.offs[seg] / a.in.step : 0;
#pragma unroll
                    for (int it = 0; it < RPT; it++) {
                        v[it * 2 + pc] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row_base[it] && in_k) {
                            const float *src = is_iv ? iv_base[it] + (k - spliced)
                                                     : row_base[it] + (size_t)((row_slot[it] + off_rows) & ring_mask) * in_dim + col;
                            v[it * 2 + pc] = __ldg(reinterpret_cast<const float4 *>(src));
                        }
                    }
                }
            };
            auto publish16 = [&](int kb, const float4 *v) {
                const int s = kb % stages;
                const uint32_t par = (uint32_t)((kb / stages) & 1);
                mbar_wait(&ts.empty[s], par ^ 1);
                const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + TM * 128;
#pragma unroll
                for (int it = 0; it < RPT; it++) {
                    const int r = it * RG + rsub;
                    const float x[8] = {v[it * 2].x, v[it * 2].y, v[it * 2].z, v[it * 2].w, v[it * 2 + 1].x, v[it * 2 + 1].y, v[it * 2 + 1].z, v[it * 2 + 1].w};
                    uint32_t hp[4], lp[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // hi = fp16(x) (|x| clamped to the fp16 range), lo = fp16((x - hi) * 2^11): x - hi is exact in fp32
                        const float x0 = fminf(fmaxf(x[2 * j], -65504.f), 65504.f), x1 = fminf(fmaxf(x[2 * j + 1], -65504.f), 65504.f);
                        const __half h0 = __float2half_rn(x0), h1 = __float2half_rn(x1);
                        const __half l0 = __float2half_rn((x0 - __half2float(h0)) * 2048.f), l1 = __float2half_rn((x1 - __half2float(h1)) * 2048.f);
                        hp[j] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                        lp[j] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                    }
                    const uint32_t o = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + o), "r"(hp[0]), "r"(hp[1]), "r"(hp[2]), "r"(hp[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + o), "r"(lp[0]), "r"(lp[1]), "r"(lp[2]), "r"(lp[3]) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&ts.full[s]);
            };
            // two K-blocks (2 x 64 K-elements) in flight per thread, buffers rotating by name
            float4 v0[RPT * 2], v1[RPT * 2];
            gather16(0, v0);
            if (nkb > 1) gather16(1, v1);
            for (int kb = 0; kb < nkb; kb += 2) {
                publish16(kb, v0);
                if (kb + 2 < nkb) gather16(kb + 2, v0);
                if (kb + 1 < nkb) {
                    publish16(kb + 1, v1);
                    if (kb + 3 < nkb) gather16(kb + 3, v1);
                }
            }
        } else {
            // gather of one K-block into registers (8 independent 16-byte requests per thread)
            auto gather = [&](int kb, float4 *v) {
                const int k = kb * TK + c * 4;
                const bool in_k = k < op.K, is_iv = k >= spliced;
                int seg = 0, col = 0;
                if (in_k && !is_iv) {
                    seg = k / in_dim;
                    col = k - seg * in_dim;
                }
                const int off_rows = in_k && !is_iv ? op.offs[seg] / a.in.step : 0;  // offsets are multiples of the input step
    #pragma unroll
                for (int it = 0; it < RPT; it++) {
                    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row_base[it] && in_k) {
                        const float *src = is_iv ? iv_base[it] + (k - spliced)
                                                 : row_base[it] + (size_t)((row_slot[it] + off_rows) & ring_mask) * in_dim + col;
                        v[it] = __ldg(reinterpret_cast<const float4 *>(src));
                    }
                }
            };
            // split one gathered K-block into hi/lo and publish it to the MMA warp
            auto publish = [&](int kb, const float4 *v) {
                const int s = kb % stages;
                const uint32_t par = (uint32_t)((kb / stages) & 1);
                mbar_wait(&ts.empty[s], par ^ 1);
                const uint32_t a_hi = smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + TM * 128;
    #pragma unroll
                for (int it = 0; it < RPT; it++) {
                    const int r = it * RG + rsub;
                    float4 h, l;
                    h.x = __uint_as_float(__float_as_uint(v[it].x) & 0xffffe000u);
                    h.y = __uint_as_float(__float_as_uint(v[it].y) & 0xffffe000u);
                    h.z = __uint_as_float(__float_as_uint(v[it].z) & 0xffffe000u);
                    h.w = __uint_as_float(__float_as_uint(v[it].w) & 0xffffe000u);
                    l.x = v[it].x - h.x; l.y = v[it].y - h.y; l.z = v[it].z - h.z; l.w = v[it].w - h.w;
                    const uint32_t o = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + o), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a_lo + o), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the MMA (async proxy)
                mbar_arrive(&ts.full[s]);
            };
            // software pipeline, unrolled by three so that the buffers rotate by name (no register moves that would wait for the
            // loads just issued): while block kb is split and stored, the loads of kb + 1 and kb + 2 are in flight
            float4 v0[RPT], v1[RPT], v2[RPT];
            gather(0, v0);
            if (nkb > 1) gather(1, v1);
            for (int kb = 0; kb < nkb; kb += 3) {
                if (kb + 2 < nkb) gather(kb + 2, v2);
                publish(kb, v0);
                if (kb + 1 < nkb) {
                    if (kb + 3 < nkb) gather(kb + 3, v0);
                    publish(kb + 1, v1);
                }
                if (kb + 2 < nkb) {
                    if (kb + 4 < nkb) gather(kb + 4, v1);
                    publish(kb + 2, v2);
                }
            }
        }
      {  // all eight producer warps: warps w and w + 4 share the 32 TMEM lanes of quarter w % 4 and take alternate 16-column chunks
        // =========================== epilogue ===========================
        const long long pt2 = prof ? clock64() : 0;
        mbar_wait(&ts.accum, 0);
        const long long pt3 = prof ? clock64() : 0;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int ch = ts.row_ch[r], t = ts.row_t[r];
        float *orow = nullptr;
        const float *brow = nullptr;
        if (ch >= 0) {
            orow = a.out.buf + ((size_t)ch * a.out.ring + (((t - a.out.t_start) / a.out.step) & (a.out.ring - 1))) * a.out.dim;
            if (op.byp_node >= 0) brow = a.byp.buf + ((size_t)ch * a.byp.ring + (((t - a.byp.t_start) / a.byp.step) & (a.byp.ring - 1))) * a.byp.dim;
        }
        const uint32_t taddr_row = tmem + ((uint32_t)(quarter * 32) << 16);
        for (int cb = (warp >> 2) * 16; cb < BN; cb += 32) {
            // the bypass row segment (4 independent 16-byte loads) is requested first, so it arrives under the TMEM loads
            const int n = n0 + cb;
            float4 byp[4];
#pragma unroll
            for (int j = 0; j < 4; j++) byp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (brow && n < op.N) {
#pragma unroll
                for (int j = 0; j < 4; j++) byp[j] = *reinterpret_cast<const float4 *>(brow + n + 4 * j);
            }
            // all partial accumulators of this 16-column chunk are requested before the single wait (the loads pipeline);
            // they are summed in fp32 round-to-nearest in a fixed order
            const int nsteps = nkb * 4;
            uint32_t v[5][16];
#pragma unroll
            for (int q = 0; q < 5; q++) {
                const bool live = q <= n_main && !(q < n_main && q >= nsteps);  // (an accumulator K never reached stays out)
                if (live) {
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                        : "=r"(v[q][0]), "=r"(v[q][1]), "=r"(v[q][2]), "=r"(v[q][3]), "=r"(v[q][4]), "=r"(v[q][5]), "=r"(v[q][6]), "=r"(v[q][7]),
                          "=r"(v[q][8]), "=r"(v[q][9]), "=r"(v[q][10]), "=r"(v[q][11]), "=r"(v[q][12]), "=r"(v[q][13]), "=r"(v[q][14]), "=r"(v[q][15])
                        : "r"(taddr_row + (uint32_t)(q * BN + cb)));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; j++) v[q][j] = 0u;
                }
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float acc[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                float x = 0.f;
                if constexpr (F16) {  // accumulator n_main holds 2^11 x the cross terms: the lo operands are scaled to stay normal in fp16
                    float cr = 0.f;
#pragma unroll
                    for (int q = 0; q < 5; q++) {
                        const float t = __uint_as_float(v[q][j]);
                        x += q < n_main ? t : 0.f;
                        cr = q == n_main ? t : cr;
                    }
                    x = fmaf(cr, 1.f / 2048.f, x);
                } else {
#pragma unroll
                    for (int q = 0; q < 5; q++) x += __uint_as_float(v[q][j]);
                }
                acc[j] = x;
            }
            if (orow && n < op.N) {
                float z[16];
                const float *bv = reinterpret_cast<const float *>(byp);
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    float x = acc[j];
                    if (op.bias) x += ts.bias[cb + j];
                    if (op.relu) x = fmaxf(x, 0.f);
                    if (op.has_bn) x = fmaf(x, ts.bn_scale[cb + j], ts.bn_offset[cb + j]);
                    if (brow) x = fmaf(op.bypass_scale, bv[j], x);
                    z[j] = x;
                }
#pragma unroll
                for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4 *>(orow + n + j) = make_float4(z[j], z[j + 1], z[j + 2], z[j + 3]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (prof && tid == 0) {
            const long long pt4 = clock64();
            atomicAdd(&g_tc_prof[0], 1ull);
            atomicAdd(&g_tc_prof[1], (unsigned long long)(pt1 - pt0));
            atomicAdd(&g_tc_prof[2], (unsigned long long)(pt2 - pt1));
            atomicAdd(&g_tc_prof[3], (unsigned long long)(pt3 - pt2));
            atomicAdd(&g_tc_prof[4], (unsigned long long)(pt4 - pt3));
        }
      }
    } else if (warp == kTmaWarp) {
        // =========================== TMA: weight boxes ===========================
        if (lane == 0) {
            for (int kb = 0; kb < nkb; kb++) {
                const int s = kb % stages;
                const uint32_t par = (uint32_t)((kb / stages) & 1);
                mbar_wait(&ts.empty[s], par ^ 1);
                unsigned char *B_hi = smem + (size_t)s * stage_bytes + 2 * TM * 128, *B_lo = B_hi + (size_t)BN * 128;
                mbar_arrive_expect_tx(&ts.full[s], 2u * (uint32_t)BN * 128u);
                const int k0 = kb * TKE;
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(B_hi)),
                             "l"(&map_hi), "r"(k0), "r"(n0), "r"(smem_u32(&ts.full[s]))
                             : "memory");
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(B_lo)),
                             "l"(&map_lo), "r"(k0), "r"(n0), "r"(smem_u32(&ts.full[s]))
                             : "memory");
            }
        }
    } else {
        // =========================== MMA issuer ===========================
        // instruction descriptor: D=f32, A=B=tf32, both K-major, N = BN, M = 128
        // (A/B format field: 2 = TF32 for kind::tf32, 0 = F16 for kind::f16)
        const uint32_t idesc = (1u << 4) | ((F16 ? 0u : 2u) << 7) | ((F16 ? 0u : 2u) << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
        for (int kb = 0; kb < nkb; kb++) {
            const int s = kb % stages;
            const uint32_t par = (uint32_t)((kb / stages) & 1);
            const long long q0 = prof ? clock64() : 0;
            mbar_wait(&ts.full[s], par);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (prof && kb == 0 && lane == 0) atomicAdd(&g_tc_prof[5], (unsigned long long)(clock64() - pt1));
            const long long q1 = prof ? clock64() : 0;
            if (prof && lane == 0) atomicAdd(&g_tc_prof[8], (unsigned long long)(q1 - q0));
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                const uint64_t dAh = umma_desc(sa), dAl = umma_desc(sa + TM * 128);
                const uint64_t dBh = umma_desc(sa + 2 * TM * 128), dBl = umma_desc(sa + 2 * TM * 128 + BN * 128);
                auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc) {
                    if constexpr (F16)
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                            "l"(da), "l"(db), "r"(idesc), "r"(acc)
                            : "memory");
                    else
                        asm volatile(
                            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                            "l"(da), "l"(db), "r"(idesc), "r"(acc)
                            : "memory");
                };
#pragma unroll
                for (int k8 = 0; k8 < 4; k8++) {
                    const uint64_t adv = (uint64_t)((k8 * 32) >> 4);  // one MMA = 32 bytes along K inside the swizzle row (8 tf32 / 16 f16)
                    const int step = kb * 4 + k8;
                    const uint32_t d_main = tmem + (uint32_t)((step % n_main) * BN);   // hi*hi: round-robin over n_main accumulators
                    const uint32_t d_cross = tmem + (uint32_t)(n_main * BN);            // cross terms: their own accumulator
                    mma(d_main, dAh + adv, dBh + adv, step >= n_main ? 1u : 0u);
                    mma(d_cross, dAl + adv, dBh + adv, step ? 1u : 0u);
                    mma(d_cross, dAh + adv, dBl + adv, 1u);
                    if (terms > 3) mma(d_cross, dAl + adv, dBl + adv, 1u);
                }
                if (prof) atomicAdd(&g_tc_prof[9], (unsigned long long)(clock64() - q1));  // issuing the K-block's MMAs (the issue blocks while the pipe is busy)
                // tcgen05.commit: arrives on the barrier when the MMAs issued so far have read their operands
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ts.empty[s])) : "memory");
                if (kb == nkb - 1) {
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&ts.accum)) : "memory");
                    if (prof) atomicAdd(&g_tc_prof[6], (unsigned long long)(clock64() - pt1));
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols));
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// Number of hi*hi accumulators (fewer truncating adds per accumulator) and the N tile that fits
// (n_main + 1) * BN fp32 columns into the 512 TMEM columns.
// TF32 split, measured on the small architecture (end-to-end log-likelihood error against the oracle, tolerance 1e-3): 4 main
// accumulators at K = 192 (6 truncating adds each, 96-column tiles) 7e-4; 3 (8 adds, 128-column tiles) 1.05e-3; 1 (24 adds,
// 256-column tiles) 1.3e-3.  The tolerance decides: 4.  The fp16 split issues half as many MMAs per K (K = 16 each), so two
// main accumulators give the same number of adds per accumulator, and the tile can be up to 160 columns wide.
static int main_accs(int K, bool f16) {
    return f16 ? (K >= 2048 ? 4 : K >= 128 ? 2 : 1) : (K >= 128 ? 4 : K >= 64 ? 2 : 1);
}
static int tile_n(int N, int K, bool f16) {
    int limit = (512 / (main_accs(K, f16) + 1)) & ~15;
    if (!f16) return N < limit ? N : limit;  // (a partial last tile is fine: TMA zero-fills beyond N and the epilogue guards its stores)
    const int ntiles = (N + limit - 1) / limit;  // equal tiles: 512 columns -> 4 x 128 rather than 3 x 160 + 32
    return (((N + ntiles - 1) / ntiles) + 15) & ~15;
}
static bool tc_f16(int mode) { return mode != 2; }

__global__ void split_f16_kernel(const float *w, __half *hi, __half *lo, int N, int K, int Kp) {
    const long long n = (long long)N * Kp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / Kp), k = (int)(i - (long long)r * Kp);
        float v = k < K ? w[(size_t)r * K + k] : 0.f;
        v = fminf(fmaxf(v, -65504.f), 65504.f);
        const __half h = __float2half_rn(v);
        hi[i] = h;
        lo[i] = __float2half_rn((v - __half2float(h)) * 2048.f);
    }
}

// Weight operands of the tensor-core path.  mode 2: W_hi = W & 0xffffe000, W_lo = W - W_hi as fp32 [N][K];
// otherwise fp16 hi and 2^11-scaled fp16 lo as [N][Kp], Kp = K rounded up to 8 (zero columns).  hi / lo need N * K * 4 bytes each.
extern "C" cudaError_t vbk_split_weights(const float *w, int N, int K, int mode, void *hi, void *lo, cudaStream_t s) {
    if (N <= 0 || K <= 0) return cudaSuccess;
    if (!tc_f16(mode)) return vbk_split_tf32(w, (float *)hi, (float *)lo, (long long)N * K, s);
    if (K < 8) return cudaErrorInvalidValue;
    const int Kp = (K + 7) & ~7;
    const long long n = (long long)N * Kp;
    split_f16_kernel<<<(int)((n + 255) / 256 > 4096 ? 4096 : (n + 255) / 256), 256, 0, s>>>(w, (__half *)hi, (__half *)lo, N, K, Kp);
    return cudaGetLastError();
}

extern "C" cudaError_t vbk_make_weight_map(const void *w, int N, int K, int mode, void *out128) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return cudaErrorNotSupported;
    const bool f16 = tc_f16(mode);
    CUtensorMap m;
    const int Kp = f16 ? (K + 7) & ~7 : K;
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * (f16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)(f16 ? 64 : TK), (cuuint32_t)tile_n(N, K, f16)};  // one 128-byte swizzle row per weight row
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void *)w, dims, strides, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    memcpy(out128, &m, 128);
    return cudaSuccess;
}

extern "C" cudaError_t vbk_gemm_tc(const GemmArgs *a, cudaStream_t s) {
    if (a->num_lanes <= 0 || a->max_rows <= 0) return cudaSuccess;
    const OpDesc &op = a->op;
    if (op.N % 16 || (op.K * 4) % 16 || !a->map_hi || !a->map_lo) return cudaErrorInvalidValue;
    const bool f16 = tc_f16(a->tc_mode);
    const int BN = tile_n(op.N, op.K, f16), n_main = main_accs(op.K, f16);
    const uint32_t stage_bytes = 2u * TM * 128u + 2u * (uint32_t)BN * 128u;
    int stages = (int)((200u * 1024u) / stage_bytes);
    if (stages > 4) stages = 4;
    const int nkb = (op.K + (f16 ? 64 : TK) - 1) / (f16 ? 64 : TK);
    if (stages > nkb) stages = nkb;
    if (stages < 1) stages = 1;
    int tmem_cols = 32;
    while (tmem_cols < (n_main + 1) * BN) tmem_cols <<= 1;  // n_main hi*hi accumulators + one for the small cross terms
    const int smem = (int)(stage_bytes * stages + 1024);
    static int done[2][16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 16 && done[f16][dev] < smem) {  // per-device function attribute: raise it only when a larger tile set is needed
        cudaError_t e = f16 ? cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                            : cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        done[f16][dev] = smem;
    }
    dim3 grid((a->max_rows + TM - 1) / TM, (op.N + BN - 1) / BN);
    TensorMapBlob mh, ml;
    memcpy(mh.b, a->map_hi, 128);
    memcpy(ml.b, a->map_lo, 128);
    const int terms = 3;  // hi*hi + lo*hi + hi*lo
    static int prof = getenv("VB_TC_PROF") ? atoi(getenv("VB_TC_PROF")) : 0;
    const bool prof_this = prof == 1 || (prof == 2 && op.K == 192 && op.N == 512) || (prof == 3 && op.K >= 1024);
    if (prof_this) {
        static bool hooked = false;
        if (!hooked) {
            hooked = true;
            atexit([] {
                unsigned long long h[16] = {};
                if (cudaMemcpyFromSymbol(h, g_tc_prof, sizeof h) != cudaSuccess || !h[0]) return;
                const double n = (double)h[0];
                fprintf(stderr, "[gemm_tc prof] tiles %.0f; cycles per tile: prologue %.0f, producer loop %.0f, wait accumulator %.0f, epilogue %.0f; MMA warp: first operands after %.0f, last commit after %.0f\n",
                        n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n);
                fprintf(stderr, "[gemm_tc prof] MMA warp per tile: waiting for operands %.0f, issuing the MMAs %.0f\n", h[8] / n, h[9] / n);
            });
        }
    }
    if (f16) gemm_tc_kernel<true><<<grid, kTcThreads, smem, s>>>(*a, mh, ml, BN, stages, tmem_cols, terms, n_main, prof_this ? 1 : 0);
    else gemm_tc_kernel<false><<<grid, kTcThreads, smem, s>>>(*a, mh, ml, BN, stages, tmem_cols, terms, n_main, prof_this ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace vb
