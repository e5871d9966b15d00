// kernels_decode.cu — K3: WFST token-passing beam search over the CSR HCLG, one persistent CTA per lane.
//
// Replaces Kaldi CudaDecoder::AdvanceDecoding as configured by the reference (max_active 7000, beam 13,
// lattice_beam 6 — [REF src/batch_model.cc:78-80]) with the canonical, order-independent semantics of
// DESIGN.md (identical to oracle/orc_decode.cc):
//   cutoff   : best+beam, exact max_active-th / min_active-th order statistic (radix select), adaptive beam
//   emitting : tot = (tok + (cost_offset - loglike[pdf])) + arc.w ; next_cutoff = min(tot) + adaptive_beam
//   recombine: one token per state = atomicMin over the 64-bit word (ordered-float(cost) << 32 | arc id)
//   closure  : epsilon arcs relaxed to the fixed point below next_cutoff
//   log      : tokens surviving the NEXT frame's cutoff are appended to the per-channel token log
//              {prev token, arc, cost}; the best path is traced on the device at stream end.
//
// B200 mapping: lanes are independent, so each CTA owns one lane for all frames of the chunk and the whole
// frame loop runs inside one launch with block-level barriers only (no grid sync, no per-frame launches,
// no host round trips).  Arc expansion is warp-cooperative: a warp takes 32 tokens, prefix-sums their
// out-degrees in shared memory and then walks the concatenated arc list 32 arcs at a time with one 16-byte
// load per arc; the frame's log-likelihood row is staged in shared memory; per-state recombination is a
// 64-bit atomicMin on an open-addressing table private to the CTA.
#include <cfloat>

#include "vb_kernels.h"

namespace vb {

namespace {
constexpr int kDecThreads = 512;
constexpr int kDecWarps = kDecThreads / 32;
constexpr unsigned long long kValMax = ~0ull;
constexpr int kEmpty = -1;

__device__ __forceinline__ unsigned ford(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord(unsigned u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned long long pack(float c, int arc) {
    return ((unsigned long long)ford(c) << 32) | (unsigned)arc;
}
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

struct Shared {
    unsigned min_ord;  // running minimum of candidate costs (ordered)
    int n_cand, n_next, error;
    int warp_cnt[kDecWarps];
    unsigned hist[256];
    unsigned sel_prefix, sel_mask;
    int sel_k;
    unsigned red_u[kDecWarps];
    unsigned long long red_ull[kDecWarps];
    // per-warp expansion staging
    int st_pref[kDecWarps][33];
    int st_a0[kDecWarps][32];
    float st_cost[kDecWarps][32];
    int st_src[kDecWarps][32];
    unsigned long long cnt_tok, cnt_arc_e, cnt_arc_eps, cnt_new;
};

struct Ctx {
    const DecArgs &a;
    Shared &sh;
    float *ll;  // shared log-likelihood row
    int *hkey;
    unsigned long long *hval;
    int *htok;
    unsigned long long *cpk;
    int *cslot, *csrc, *rank;
    int tid, warp, lane;
};

__device__ __forceinline__ int agg_inc(int *counter) {
    unsigned m = __activemask();
    int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & lanemask_lt());
}

// insert (state, packed) ; records a candidate when it improved the state's best word
__device__ __forceinline__ void relax(Ctx &c, int state, unsigned long long pk, int src) {
    const unsigned mask = (unsigned)c.a.hash_size - 1;
    unsigned h = ((unsigned)state * 2654435761u) >> 7 & mask;
    int probes = 0;
    for (;;) {
        int cur = __ldcg(c.hkey + h);
        if (cur == state) break;
        if (cur == kEmpty) {
            int prev = atomicCAS(c.hkey + h, kEmpty, state);
            if (prev == kEmpty || prev == state) break;
        }
        h = (h + 1) & mask;
        if (++probes > c.a.hash_size) {
            c.sh.error = 1;
            return;
        }
    }
    unsigned long long old = atomicMin(c.hval + h, pk);
    if (pk < old) {
        int idx = agg_inc(&c.sh.n_cand);
        if (idx < c.a.cand_cap) {
            c.cpk[idx] = pk;
            c.cslot[idx] = (int)h;
            c.csrc[idx] = src;
        } else {
            c.sh.error = 2;
        }
    }
}

__device__ float block_min(Ctx &c, const float *cost, int n) {
    unsigned m = 0xffffffffu;
    for (int i = c.tid; i < n; i += kDecThreads) m = min(m, ford(cost[i]));
    for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (c.lane == 0) c.sh.red_u[c.warp] = m;
    __syncthreads();
    m = c.sh.red_u[0];
    for (int w = 1; w < kDecWarps; w++) m = min(m, c.sh.red_u[w]);
    __syncthreads();
    return unord(m);
}

// exact k-th smallest (0-based) of cost[0..n) by 4-pass radix select on the ordered key
__device__ float block_select(Ctx &c, const float *cost, int n, int k) {
    if (c.tid == 0) {
        c.sh.sel_prefix = 0;
        c.sh.sel_mask = 0;
        c.sh.sel_k = k;
    }
    for (int pass = 3; pass >= 0; pass--) {
        const int shift = pass * 8;
        if (c.tid < 256) c.sh.hist[c.tid] = 0;
        __syncthreads();
        const unsigned prefix = c.sh.sel_prefix, mask = c.sh.sel_mask;
        for (int i = c.tid; i < n; i += kDecThreads) {
            unsigned key = ford(cost[i]);
            if ((key & mask) == prefix) {
                unsigned bin = (key >> shift) & 255u;
                unsigned grp = __match_any_sync(__activemask(), bin);
                if ((int)c.lane == __ffs(grp) - 1) atomicAdd(&c.sh.hist[bin], (unsigned)__popc(grp));
            }
        }
        __syncthreads();
        if (c.warp == 0) {
            unsigned loc[8], s = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                loc[j] = c.sh.hist[c.lane * 8 + j];
                s += loc[j];
            }
            unsigned incl = s;
            for (int o = 1; o < 32; o <<= 1) {
                unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if ((int)c.lane >= o) incl += v;
            }
            unsigned excl = incl - s;
            const unsigned kk = (unsigned)c.sh.sel_k;
            bool mine = kk >= excl && kk < incl;
            if (mine) {
                unsigned cum = excl;
                int b = 0;
                for (; b < 8; b++) {
                    if (cum + loc[b] > kk) break;
                    cum += loc[b];
                }
                c.sh.sel_k = (int)(kk - cum);
                c.sh.sel_prefix = prefix | ((unsigned)(c.lane * 8 + b) << shift);
                c.sh.sel_mask = mask | (255u << shift);
            }
        }
        __syncthreads();
    }
    float r = unord(c.sh.sel_prefix);
    __syncthreads();
    return r;
}

// GetCutoff of LatticeFasterDecoder (see oracle/orc_decode.cc get_cutoff)
__device__ float get_cutoff(Ctx &c, const float *cost, int n, float *adaptive_beam, float *best_out) {
    const DecArgs &a = c.a;
    float best = block_min(c, cost, n);
    *best_out = best;
    float beam_cutoff = best + a.beam, min_active_cutoff = INFINITY, max_active_cutoff = INFINITY;
    if (n > a.max_active) max_active_cutoff = block_select(c, cost, n, a.max_active);
    if (max_active_cutoff < beam_cutoff) {
        *adaptive_beam = max_active_cutoff - best + a.beam_delta;
        return max_active_cutoff;
    }
    if (n > a.min_active) {
        if (a.min_active == 0) min_active_cutoff = best;
        else min_active_cutoff = block_select(c, cost, n, a.min_active);
    }
    if (min_active_cutoff > beam_cutoff) {
        *adaptive_beam = min_active_cutoff - best + a.beam_delta;
        return min_active_cutoff;
    }
    *adaptive_beam = a.beam;
    return beam_cutoff;
}

// Warp-cooperative arc walk.  Each lane contributes one token (a0, deg, cost, src); the warp then visits the
// concatenated arc list 32 arcs at a time.  EMIT: acoustic cost from the shared log-likelihood row.
template <bool EMIT>
__device__ __forceinline__ void warp_expand(Ctx &c, int a0, int deg, float cost, int src, float cost_offset,
                                            float adaptive_beam, float hard_cutoff) {
    Shared &sh = c.sh;
    const int w = c.warp, lane = c.lane;
    int incl = deg;
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) return;
    sh.st_pref[w][lane] = incl - deg;
    sh.st_a0[w][lane] = a0;
    sh.st_cost[w][lane] = cost;
    sh.st_src[w][lane] = src;
    if (lane == 0) sh.st_pref[w][32] = total;
    __syncwarp();
    for (int j0 = 0; j0 < total; j0 += 32) {
        const int j = j0 + lane;
        const bool valid = j < total;
        float tot = INFINITY;
        int arc = 0, next = 0, q = 0;
        if (valid) {
            int lo = 0, hi = 32;  // largest q with pref[q] <= j
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (sh.st_pref[w][mid] <= j) lo = mid; else hi = mid;
            }
            q = lo;
            arc = sh.st_a0[w][q] + (j - sh.st_pref[w][q]);
            const int4 av = __ldg(c.a.g.arcs + arc);
            next = av.y;
            const float wgt = __int_as_float(av.x);
            if (EMIT) {
                float ac = cost_offset - c.a.acoustic_scale * c.ll[av.z];
                tot = sh.st_cost[w][q] + ac + wgt;
            } else {
                tot = sh.st_cost[w][q] + wgt;
            }
        }
        float cut = hard_cutoff;
        if (EMIT) {
            unsigned m = ford(tot);
            for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) atomicMin(&sh.min_ord, m);
            __syncwarp();
            cut = unord(*(volatile unsigned *)&sh.min_ord) + adaptive_beam;  // loose (>= final) cutoff
        }
        if (valid && tot < cut) relax(c, next, pack(tot, arc), sh.st_src[w][q]);
    }
    __syncwarp();
}

// epsilon closure over candidates [lo, hi) until no candidate is added; returns total candidate count
__device__ int closure(Ctx &c, int lo, int hi, float cutoff) {
    const DecArgs &a = c.a;
    unsigned long long arcs_seen = 0;
    while (lo < hi) {
        for (int base = lo + c.warp * 32; base < hi; base += kDecWarps * 32) {
            const int i = base + c.lane;
            int a0 = 0, deg = 0, slot = 0;
            float cost = 0.f;
            if (i < hi) {
                unsigned long long pk = c.cpk[i];
                slot = c.cslot[i];
                cost = unord((unsigned)(pk >> 32));
                if (__ldcg(c.hval + slot) == pk && cost < cutoff) {
                    int s = __ldcg(c.hkey + slot);
                    a0 = __ldg(a.g.eps_begin + s);
                    deg = __ldg(a.g.e_begin + s + 1) - a0;
                }
            }
            arcs_seen += (unsigned)deg;
            warp_expand<false>(c, a0, deg, cost, slot, 0.f, 0.f, cutoff);
        }
        __syncthreads();
        lo = hi;
        hi = min(c.sh.n_cand, a.cand_cap);
        __syncthreads();
    }
    if (arcs_seen) atomicAdd(&c.sh.cnt_arc_eps, arcs_seen);
    return hi;
}

// turn the winning candidates into the next frame's token list; clears the hash table
__device__ void finalize_tokens(Ctx &c, int n_emit, int n_cand, float cutoff, int *t_state, float *t_cost, int *t_arc,
                                int *t_prev) {
    const DecArgs &a = c.a;
    for (int i = c.tid; i < n_cand; i += kDecThreads) {
        unsigned long long pk = c.cpk[i];
        int slot = c.cslot[i];
        float cost = unord((unsigned)(pk >> 32));
        if (__ldcg(c.hval + slot) == pk && cost < cutoff) {
            int idx = agg_inc(&c.sh.n_next);
            if (idx < a.tok_cap) {
                t_state[idx] = __ldcg(c.hkey + slot);
                t_cost[idx] = cost;
                t_arc[idx] = (int)(unsigned)pk;
                c.htok[slot] = idx;
                if (i < n_emit) t_prev[idx] = c.csrc[i];
            } else {
                c.sh.error = 3;
            }
        }
    }
    __syncthreads();
    for (int i = n_emit + c.tid; i < n_cand; i += kDecThreads) {
        unsigned long long pk = c.cpk[i];
        int slot = c.cslot[i];
        float cost = unord((unsigned)(pk >> 32));
        if (__ldcg(c.hval + slot) == pk && cost < cutoff) {
            int idx = __ldcg(c.htok + slot);
            if (idx < a.tok_cap) t_prev[idx] = -2 - __ldcg(c.htok + c.csrc[i]);
        }
    }
    __syncthreads();
    if (c.sh.error == 1 || c.sh.error == 2) {
        // a candidate was dropped (table/candidate overflow): entries may exist that no candidate points at
        for (int i = c.tid; i < a.hash_size; i += kDecThreads) {
            c.hkey[i] = kEmpty;
            c.hval[i] = kValMax;
        }
    } else {
        for (int i = c.tid; i < n_cand; i += kDecThreads) {
            int slot = c.cslot[i];
            c.hkey[slot] = kEmpty;
            c.hval[slot] = kValMax;
        }
    }
    __syncthreads();
}
}  // namespace

__global__ void __launch_bounds__(kDecThreads, 2) decode_kernel(DecArgs a) {
    extern __shared__ __align__(16) float s_ll[];
    __shared__ Shared sh;
    const int tid = threadIdx.x;
    const size_t g = blockIdx.x;
    Ctx c{a, sh, s_ll,
          a.hash_key + g * a.hash_size, a.hash_val + g * a.hash_size, a.hash_tok + g * a.hash_size,
          a.cand_packed + g * a.cand_cap, a.cand_slot + g * a.cand_cap, a.cand_src + g * a.cand_cap,
          a.rank + g * a.tok_cap, tid, tid >> 5, tid & 31};
    if (tid == 0) {
        sh.cnt_tok = sh.cnt_arc_e = sh.cnt_arc_eps = sh.cnt_new = 0;
    }
    const int npdf = a.out_node.dim;
    for (int l = blockIdx.x; l < a.num_lanes; l += gridDim.x) {
        const LaneDesc ln = a.lanes[l];
        const int ch = ln.channel;
        DecChannelState *cs = a.cs + ch;
        const size_t tbase = (size_t)ch * 2 * a.tok_cap;
        int *log_prev = a.log_prev + (size_t)ch * a.log_cap;
        int *log_arc = a.log_arc + (size_t)ch * a.log_cap;
        float *log_cost = a.log_cost + (size_t)ch * a.log_cap;
        int *log_state = a.log_state ? a.log_state + (size_t)ch * a.log_cap : nullptr;
        int *frame_off = a.log_frame_off + (size_t)ch * (a.max_frames + 2);
        __syncthreads();
        if (tid == 0) sh.error = ln.first ? 0 : cs->error;
        int n_cur, parity, frame, log_count;
        if (ln.first) {
            // InitDecoding: start token + epsilon closure with cutoff = beam
            if (tid == 0) {
                sh.n_cand = 0;
                sh.n_next = 0;
            }
            __syncthreads();
            if (tid == 0) relax(c, a.g.start, pack(0.f, -1), -1);
            __syncthreads();
            int nc = closure(c, 0, min(sh.n_cand, a.cand_cap), a.beam);
            finalize_tokens(c, 1, nc, INFINITY, a.tok_state + tbase, a.tok_cost + tbase, a.tok_arc + tbase, a.tok_prev + tbase);
            n_cur = min(sh.n_next, a.tok_cap);
            parity = 0;
            frame = 0;
            log_count = 0;
        } else {
            n_cur = cs->n_cur;
            parity = cs->parity;
            frame = cs->frame;
            log_count = cs->log_count;
        }
        const int nf = a.out_table[l].n_rows;
        const int t_first = a.out_table[l].t_begin;
        const int total_frames = nf + (ln.last ? 1 : 0);  // the extra pass logs the final frame's tokens
        for (int fi = 0; fi < total_frames; fi++) {
            const bool final_pass = fi == nf;
            const int *t_state = a.tok_state + tbase + (size_t)parity * a.tok_cap;
            const float *t_cost = a.tok_cost + tbase + (size_t)parity * a.tok_cap;
            const int *t_arc = a.tok_arc + tbase + (size_t)parity * a.tok_cap;
            const int *t_prev = a.tok_prev + tbase + (size_t)parity * a.tok_cap;
            int *n_state = a.tok_state + tbase + (size_t)(parity ^ 1) * a.tok_cap;
            float *n_cost = a.tok_cost + tbase + (size_t)(parity ^ 1) * a.tok_cap;
            int *n_arc = a.tok_arc + tbase + (size_t)(parity ^ 1) * a.tok_cap;
            int *n_prev = a.tok_prev + tbase + (size_t)(parity ^ 1) * a.tok_cap;
            if (n_cur == 0 && !final_pass) { frame++; continue; }  // search died: nothing to expand
            float adaptive_beam = a.beam, best = 0.f, cur_cutoff = INFINITY;
            if (!final_pass) {
                // stage this frame's log-likelihood row
                const float *row = a.out_node.buf + ((size_t)ch * a.out_node.ring +
                                                     (((t_first + fi * a.out_node.step) - a.out_node.t_start) / a.out_node.step & (a.out_node.ring - 1))) * npdf;
                for (int i = tid * 4; i < npdf; i += kDecThreads * 4) *reinterpret_cast<float4 *>(s_ll + i) = *reinterpret_cast<const float4 *>(row + i);
                cur_cutoff = get_cutoff(c, t_cost, n_cur, &adaptive_beam, &best);
            }
            if (tid == 0) {
                sh.min_ord = 0xffffffffu;
                sh.n_cand = 0;
                sh.n_next = 0;
            }
            // ---- survivors: rank (stable), log, expand ----
            const int span = ((n_cur + kDecWarps - 1) / kDecWarps + 31) & ~31;
            const int wbeg = min(c.warp * span, n_cur), wend = min(wbeg + span, n_cur);
            int cnt = 0;
            for (int i = wbeg + c.lane; i < wend + ((32 - (wend - wbeg) % 32) % 32); i += 32) {
                bool f = i < wend && t_cost[i] <= cur_cutoff;
                cnt += __popc(__ballot_sync(0xffffffffu, f));
            }
            if (c.lane == 0) sh.warp_cnt[c.warp] = cnt;
            __syncthreads();
            int base = 0, n_surv = 0;
            for (int w = 0; w < kDecWarps; w++) {
                if (w < c.warp) base += sh.warp_cnt[w];
                n_surv += sh.warp_cnt[w];
            }
            const bool log_ok = frame <= a.max_frames && log_count + n_surv <= a.log_cap;
            if (!log_ok && tid == 0) sh.error = 4;
            // pass 2a: ranks (needed by every token's epsilon back-reference before logging)
            {
                int b = base;
                for (int i0 = wbeg; i0 < wend; i0 += 32) {
                    int i = i0 + c.lane;
                    bool f = i < wend && t_cost[i] <= cur_cutoff;
                    unsigned bal = __ballot_sync(0xffffffffu, f);
                    if (i < wend) c.rank[i] = f ? b + __popc(bal & lanemask_lt()) : -1;
                    b += __popc(bal);
                }
            }
            __syncthreads();
            const float cost_offset = -best;
            unsigned long long arcs_seen = 0;
            for (int i0 = wbeg; i0 < wend; i0 += 32) {
                int i = i0 + c.lane;
                int a0 = 0, deg = 0, li = -1;
                float cost = 0.f;
                if (i < wend) {
                    int r = c.rank[i];
                    if (r >= 0) {
                        cost = t_cost[i];
                        li = log_count + r;
                        int s = t_state[i];
                        if (log_ok) {
                            int pv = t_prev[i];
                            if (pv <= -2) {  // epsilon predecessor lives in this frame: it must have survived too
                                int pr = c.rank[-2 - pv];
                                if (pr < 0) sh.error = 5;  // only possible with negative epsilon weights
                                pv = pr < 0 ? -1 : log_count + pr;
                            }
                            log_prev[li] = pv;
                            log_arc[li] = t_arc[i];
                            log_cost[li] = cost;
                            if (log_state) log_state[li] = s;
                        }
                        if (!final_pass) {
                            a0 = __ldg(a.g.e_begin + s);
                            deg = __ldg(a.g.eps_begin + s) - a0;
                        }
                    }
                }
                if (!final_pass) {
                    arcs_seen += (unsigned)deg;
                    warp_expand<true>(c, a0, deg, cost, li, cost_offset, adaptive_beam, INFINITY);
                }
            }
            if (log_ok && tid == 0) {
                frame_off[frame] = log_count;
                frame_off[frame + 1] = log_count + n_surv;
            }
            if (arcs_seen) atomicAdd(&sh.cnt_arc_e, arcs_seen);
            if (tid == 0) sh.cnt_tok += (unsigned)n_surv;
            log_count += log_ok ? n_surv : 0;
            __syncthreads();
            if (final_pass) break;
            const float next_cutoff = unord(sh.min_ord) + adaptive_beam;
            const int n_emit = min(sh.n_cand, a.cand_cap);
            __syncthreads();
            const int nc = closure(c, 0, n_emit, next_cutoff);
            finalize_tokens(c, n_emit, nc, next_cutoff, n_state, n_cost, n_arc, n_prev);
            n_cur = min(sh.n_next, a.tok_cap);
            if (tid == 0) sh.cnt_new += (unsigned)n_cur;
            parity ^= 1;
            frame++;
        }
        // ---- stream end: best token (cost + final, ties by state id) and on-device traceback ----
        if (ln.last) {
            const int lo = frame <= a.max_frames ? frame_off[frame] : 0, hi = log_count;
            const int *t_state = a.tok_state + tbase + (size_t)parity * a.tok_cap;
            // survivors of the final pass are all tokens, logged in list order: log index = lo + i
            unsigned long long bw[2] = {kValMax, kValMax};
            for (int i = tid; i < n_cur && lo + i < hi; i += kDecThreads) {
                int s = t_state[i];
                float cst = log_cost[lo + i];
                float fc = __ldg(a.g.final_cost + s);
                // 64-bit key: ordered(total) , state ; the winner's log index is recovered by a second scan
                if (fc != INFINITY) bw[0] = min(bw[0], ((unsigned long long)ford(cst + fc) << 32) | (unsigned)s);
                bw[1] = min(bw[1], ((unsigned long long)ford(cst) << 32) | (unsigned)s);
            }
            unsigned long long win[2];
            for (int p = 0; p < 2; p++) {
                unsigned long long v = bw[p];
                for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (c.lane == 0) sh.red_ull[c.warp] = v;
                __syncthreads();
                v = sh.red_ull[0];
                for (int w = 1; w < kDecWarps; w++) v = min(v, sh.red_ull[w]);
                win[p] = v;
                __syncthreads();
            }
            const int pass = win[0] != kValMax ? 0 : 1;
            const unsigned long long target = win[pass];
            if (tid == 0) {
                cs->reached_final = pass == 0;
                cs->best_cost = unord((unsigned)(target >> 32));
                cs->path_len = 0;
            }
            __syncthreads();
            if (target != kValMax) {
                const int want_state = (int)(unsigned)target;
                for (int i = tid; i < n_cur && lo + i < hi; i += kDecThreads) {
                    if (t_state[i] == want_state) {
                        // single winner thread walks the back pointers (stream end only)
                        int *path = a.path + (size_t)ch * a.path_cap;
                        int n = 0;
                        for (int li = lo + i; li >= 0 && n < a.path_cap; li = log_prev[li]) {
                            int arc = log_arc[li];
                            if (arc < 0) break;
                            path[n++] = arc;
                        }
                        cs->path_len = n;
                    }
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            cs->n_cur = n_cur;
            cs->parity = parity;
            cs->frame = frame;
            cs->log_count = log_count;
            cs->error = sh.error;
        }
    }
    __syncthreads();
    if (tid == 0 && a.counters) {
        atomicAdd(a.counters + 0, sh.cnt_tok);
        atomicAdd(a.counters + 1, sh.cnt_arc_e);
        atomicAdd(a.counters + 2, sh.cnt_arc_eps);
        atomicAdd(a.counters + 3, sh.cnt_new);
    }
}

extern "C" int vbk_decode_max_grid(int device) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms * 2;
}

extern "C" cudaError_t vbk_decode(const DecArgs *a, cudaStream_t s) {
    if (a->num_lanes <= 0) return cudaSuccess;
    int smem = (a->out_node.dim * 4 + 15) & ~15;
    static int configured = 0;
    if (smem > 40000 && configured < smem) {
        cudaError_t e = cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    int grid = a->num_lanes < a->grid ? a->num_lanes : a->grid;
    decode_kernel<<<grid, kDecThreads, smem, s>>>(*a);
    return cudaGetLastError();
}

}  // namespace vb
