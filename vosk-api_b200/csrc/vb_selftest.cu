// vb_selftest.cu — GEMM self-test entry (diagnostics): runs one synthetic TDNN op through the fp32 and the
// tensor-core kernel and reports their errors against a double-precision host product.
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "vb_kernels.h"

using namespace vb;

extern "C" int vosk_b200_gemm_selftest(int M, int N, int K, int bits, double *out /* [4] */) {
    // out: max|fp32-ref|, max|tc-ref|, rms(tc-ref), rms(ref)
    std::mt19937 rng(123);
    std::normal_distribution<float> nd(0.f, 1.f);
    auto quant = [&](float v) {
        if (bits <= 0) return v;
        int e;
        float m = std::frexp(v, &e);
        float s = std::ldexp(1.f, bits);
        return std::ldexp(std::round(m * s) / s, e);
    };
    std::vector<float> A((size_t)M * K), W((size_t)N * K);
    for (auto &v : A) v = quant(nd(rng));
    for (auto &v : W) v = quant(nd(rng) / std::sqrt((float)K));
    int ring = 1;
    while (ring < M) ring <<= 1;
    float *dA, *dW, *dHi, *dLo, *dC0, *dC1;
    cudaMalloc(&dA, (size_t)ring * K * 4);
    cudaMemset(dA, 0, (size_t)ring * K * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&dW, W.size() * 4);
    cudaMalloc(&dHi, W.size() * 4);
    cudaMalloc(&dLo, W.size() * 4);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    const int tc_mode = getenv("VB_TC_MODE") ? atoi(getenv("VB_TC_MODE")) : 1;
    vbk_split_weights(dW, N, K, tc_mode, dHi, dLo, 0);
    cudaMalloc(&dC0, (size_t)ring * N * 4);
    cudaMalloc(&dC1, (size_t)ring * N * 4);
    LaneDesc ln{};
    ln.channel = 0;
    NodeLane nl{0, M};
    int rowoff[2] = {0, M};
    LaneDesc *dL;
    NodeLane *dT;
    int *dR;
    cudaMalloc(&dL, sizeof ln);
    cudaMalloc(&dT, sizeof nl);
    cudaMalloc(&dR, sizeof rowoff);
    cudaMemcpy(dL, &ln, sizeof ln, cudaMemcpyHostToDevice);
    cudaMemcpy(dT, &nl, sizeof nl, cudaMemcpyHostToDevice);
    cudaMemcpy(dR, rowoff, sizeof rowoff, cudaMemcpyHostToDevice);
    std::vector<int2> rows(M);
    for (int m = 0; m < M; m++) rows[m] = make_int2(0, m);
    int2 *dRows;
    cudaMalloc(&dRows, (size_t)M * sizeof(int2));
    cudaMemcpy(dRows, rows.data(), (size_t)M * sizeof(int2), cudaMemcpyHostToDevice);
    alignas(64) unsigned char mh[128], ml[128];
    if (vbk_make_weight_map(dHi, N, K, tc_mode, mh) != cudaSuccess || vbk_make_weight_map(dLo, N, K, tc_mode, ml) != cudaSuccess) return -1;
    GemmArgs g{};
    g.op.in_node = 0; g.op.out_node = 1; g.op.byp_node = -1; g.op.n_off = 1; g.op.offs[0] = 0; g.op.K = K; g.op.N = N;
    g.op.W = dW; g.op.W_hi = dHi; g.op.W_lo = dLo;
    g.in = NodeDesc{K, 1, ring, 0, 0, dA};
    g.out = NodeDesc{N, 1, ring, 0, 0, dC0};
    g.byp = g.in;
    g.lanes = dL; g.num_lanes = 1; g.table = dT; g.rowoff = dR; g.rows = dRows; g.ivec = dA; g.ivec_dim = 4; g.max_rows = M;
    g.map_hi = mh; g.map_lo = ml; g.tc_mode = tc_mode;
    if (vbk_gemm_fp32(&g, 0) != cudaSuccess) return -2;
    g.out.buf = dC1;
    if (vbk_gemm_tc(&g, 0) != cudaSuccess) return -3;
    if (cudaDeviceSynchronize() != cudaSuccess) return -4;
    std::vector<float> C0((size_t)M * N), C1((size_t)M * N);
    cudaMemcpy(C0.data(), dC0, C0.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(C1.data(), dC1, C1.size() * 4, cudaMemcpyDeviceToHost);
    double e0 = 0, e1 = 0, s1 = 0, sr = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            double r = 0;
            for (int k = 0; k < K; k++) r += (double)A[(size_t)m * K + k] * W[(size_t)n * K + k];
            double d0 = C0[(size_t)m * N + n] - r, d1 = C1[(size_t)m * N + n] - r;
            e0 = std::fmax(e0, std::fabs(d0));
            e1 = std::fmax(e1, std::fabs(d1));
            s1 += d1 * d1;
            sr += r * r;
        }
    out[0] = e0; out[1] = e1; out[2] = std::sqrt(s1 / ((double)M * N)); out[3] = std::sqrt(sr / ((double)M * N));
    cudaFree(dA); cudaFree(dW); cudaFree(dHi); cudaFree(dLo); cudaFree(dC0); cudaFree(dC1); cudaFree(dL); cudaFree(dT); cudaFree(dR); cudaFree(dRows);
    return 0;
}
