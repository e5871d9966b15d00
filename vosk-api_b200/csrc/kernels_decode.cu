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
// no host round trips).  CTAs are small (256 threads, 4 per SM) so that all lanes of a step are resident at
// once and one lane's barrier/latency bubbles are filled by its neighbours.  Emitting-arc expansion is
// balanced per ARC, not per token: survivors with out-arcs are compacted with an exclusive prefix sum of
// their out-degrees, the concatenated arc list is cut into 32-arc windows dealt round-robin to the warps,
// and a window finds the owners of its arcs from one coalesced load of the prefix array plus a shared-memory
// marker scan — every lane then issues one 16-byte arc load.  The frame's log-likelihood row is staged in
// shared memory; per-state recombination is a 64-bit atomicMin on an open-addressing table private to the
// CTA whose active window is sized per frame so that small frames stay L2-resident.
#include <cfloat>
#include <climits>

#include "vb_kernels.h"

namespace vb {

namespace {
constexpr int kDecBlocksPerSM = 3;   // 256-thread variant; the 1024-thread variant (heavy lanes) runs 1 per SM
constexpr int kSmemSlots = 4096;  // level-1 (shared memory) table entries per CTA
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

template <int NT>
struct Shared {
    unsigned min_ord;  // running minimum of candidate costs (ordered)
    int n_cand, n_next, error;
    int warp_cnt[(NT / 32)], warp_exp[(NT / 32)], warp_deg[(NT / 32)];
    unsigned hist[256];
    unsigned sel_prefix, sel_mask;
    int sel_k;
    unsigned red_u[(NT / 32)];
    unsigned long long red_ull[(NT / 32)];
    int own[(NT / 32)][32];  // per-warp marker array of the arc-window owner scan
};

template <int NT>
struct Ctx {
    const DecArgs &a;
    Shared<NT> &sh;
    float *ll;  // shared log-likelihood row
    int *skey;                 // level-1 table in shared memory: keys / 64-bit best words
    unsigned long long *sval;
    int *hkey;                 // level-2 table in global memory (overflow of the bounded level-1 probe)
    unsigned long long *hval;
    int *htok;
    int4 *cand;                // {packed lo, packed hi, slot, src}
    int *rank;
    int *sv_pref, *sv_a0, *sv_src, *win_owner;
    float *sv_cost;
    int tid, warp, lane;
    bool use_l1;     // level-1 (shared) table enabled for this frame
    unsigned hmask;  // this frame's table window (power of two - 1): the table is empty between frames, so any
                     // power-of-two prefix of it is a valid table; small frames stay L2-resident
};

__device__ __forceinline__ int agg_inc(int *counter) {
    unsigned m = __activemask();
    int leader = __ffs(m) - 1;
    int base = 0;
    if ((int)(threadIdx.x & 31) == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    return base + __popc(m & lanemask_lt());
}

// Table slots: [0, kSmemSlots) = shared-memory level, kSmemSlots + g = global level.  A state lives in level 1 iff
// a free or matching slot existed within kProbe1 probes at its first insertion; slots are never freed inside a
// frame, so every later lookup of the same state takes the same decision.
constexpr int kProbe1 = 16;
template <int NT>
__device__ __forceinline__ unsigned long long tab_val(const Ctx<NT> &c, int slot) {
    return slot < kSmemSlots ? *(volatile unsigned long long *)(c.sval + slot) : __ldcg(c.hval + (slot - kSmemSlots));
}
template <int NT>
__device__ __forceinline__ int tab_key(const Ctx<NT> &c, int slot) {
    return slot < kSmemSlots ? *(volatile int *)(c.skey + slot) : __ldcg(c.hkey + (slot - kSmemSlots));
}
// after the winners are known the key field of a level-1 slot is reused for the token index
template <int NT>
__device__ __forceinline__ void tab_set_tok(const Ctx<NT> &c, int slot, int idx) {
    if (slot < kSmemSlots) c.skey[slot] = idx; else c.htok[slot - kSmemSlots] = idx;
}
template <int NT>
__device__ __forceinline__ int tab_tok(const Ctx<NT> &c, int slot) {
    return slot < kSmemSlots ? *(volatile int *)(c.skey + slot) : __ldcg(c.htok + (slot - kSmemSlots));
}

// insert (state, packed) ; records a candidate when it improved the state's best word
template <int NT>
__device__ __forceinline__ void relax(Ctx<NT> &c, int state, unsigned long long pk, int src) {
    const unsigned hash = ((unsigned)state * 2654435761u) >> 7;
    int slot = -1;
    unsigned h = hash & (kSmemSlots - 1);
#pragma unroll 1
    for (int p = 0; p < kProbe1 && c.use_l1; p++) {
        int prev = atomicCAS(c.skey + h, kEmpty, state);
        if (prev == kEmpty || prev == state) {
            slot = (int)h;
            break;
        }
        h = (h + 1) & (kSmemSlots - 1);
    }
    unsigned long long old;
    if (slot >= 0) {
        old = atomicMin(c.sval + slot, pk);
    } else {
        const unsigned mask = c.hmask;
        unsigned g = hash & mask;
        int probes = 0;
        for (;;) {
            int prev = atomicCAS(c.hkey + g, kEmpty, state);
            if (prev == kEmpty || prev == state) break;
            g = (g + 1) & mask;
            if (++probes > (int)mask) {
                c.sh.error = 1;
                return;
            }
        }
        old = atomicMin(c.hval + g, pk);
        slot = kSmemSlots + (int)g;
    }
    if (pk < old) {
        int idx = agg_inc(&c.sh.n_cand);
        if (idx < c.a.cand_cap) c.cand[idx] = make_int4((int)(unsigned)pk, (int)(unsigned)(pk >> 32), slot, src);
        else c.sh.error = 2;
    }
}

// minimum cost and the index of one token attaining it
template <int NT>
__device__ float block_min(Ctx<NT> &c, const float *cost, int n, int *arg) {
    unsigned long long m = kValMax;
    for (int i = c.tid; i < n; i += NT) m = min(m, ((unsigned long long)ford(cost[i]) << 32) | (unsigned)i);
    for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (c.lane == 0) c.sh.red_ull[c.warp] = m;
    __syncthreads();
    m = c.sh.red_ull[0];
    for (int w = 1; w < (NT / 32); w++) m = min(m, c.sh.red_ull[w]);
    __syncthreads();
    *arg = (int)(unsigned)m;
    return unord((unsigned)(m >> 32));
}

// exact k-th smallest (0-based) of cost[0..n) by 4-pass radix select on the ordered key
template <int NT>
__device__ float block_select(Ctx<NT> &c, const float *cost, int n, int k) {
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
        for (int i = c.tid; i < n; i += NT) {
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

// GetCutoff of LatticeFasterDecoder (see oracle/orc_decode.cc get_cutoff).  The order statistics only matter
// when they fall on the right side of best+beam, which a counting pass decides:
//   kth(max_active) <  beam_cutoff  <=>  #{cost <  beam_cutoff} >  max_active
//   kth(min_active) >  beam_cutoff  <=>  #{cost <= beam_cutoff} <= min_active
// so the exact radix select runs only on the frames where it changes the result.
template <int NT>
__device__ float get_cutoff(Ctx<NT> &c, const float *cost, int n, float *adaptive_beam, float *best_out, int *best_idx) {
    const DecArgs &a = c.a;
    float best = block_min(c, cost, n, best_idx);
    *best_out = best;
    const float beam_cutoff = best + a.beam;
    int lt = 0, le = 0;
    for (int i = c.tid; i < n; i += NT) {
        float v = cost[i];
        lt += v < beam_cutoff;
        le += v <= beam_cutoff;
    }
    unsigned long long tot = ((unsigned long long)lt << 32) | (unsigned)le;
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (c.lane == 0) c.sh.red_ull[c.warp] = tot;
    __syncthreads();
    tot = 0;
    for (int w = 0; w < (NT / 32); w++) tot += c.sh.red_ull[w];
    __syncthreads();
    const int n_lt = (int)(tot >> 32), n_le = (int)(unsigned)tot;
    if (n > a.max_active && n_lt > a.max_active) {
        float max_active_cutoff = block_select(c, cost, n, a.max_active);
        *adaptive_beam = max_active_cutoff - best + a.beam_delta;
        return max_active_cutoff;
    }
    if (n <= a.min_active) {  // fewer tokens than min_active: no pruning (Kaldi leaves min_active_cutoff at +inf)
        *adaptive_beam = INFINITY;
        return INFINITY;
    }
    if (a.min_active > 0 && n_le <= a.min_active) {
        float min_active_cutoff = block_select(c, cost, n, a.min_active);
        *adaptive_beam = min_active_cutoff - best + a.beam_delta;
        return min_active_cutoff;
    }
    *adaptive_beam = a.beam;
    return beam_cutoff;
}

// epsilon closure over candidates [lo, hi) until no candidate is added; returns total candidate count.
// Epsilon out-degrees are tiny (0-2), so one thread per candidate is balanced.
template <int NT>
__device__ int closure(Ctx<NT> &c, int lo, int hi, float cutoff, unsigned long long *arcs_seen) {
    const DecArgs &a = c.a;
    while (lo < hi) {
        for (int i = lo + c.tid; i < hi; i += NT) {
            const int4 cd = c.cand[i];
            const unsigned long long pk = ((unsigned long long)(unsigned)cd.y << 32) | (unsigned)cd.x;
            const int slot = cd.z;
            const float cost = unord((unsigned)cd.y);
            if (cost < cutoff && tab_val(c, slot) == pk) {
                const int s = tab_key(c, slot);
                const int a0 = __ldg(&a.g.state_arcs[s].y), a1 = __ldg(&a.g.state_arcs[s + 1].x);
                *arcs_seen += (unsigned)(a1 - a0);
                for (int arc = a0; arc < a1; arc++) {
                    const int4 av = __ldg(a.g.arcs + arc);
                    const float tot = cost + __int_as_float(av.x);
                    if (tot < cutoff) relax(c, av.y, pack(tot, arc), slot);
                }
            }
        }
        __syncthreads();
        lo = hi;
        hi = min(c.sh.n_cand, a.cand_cap);
        __syncthreads();
    }
    return hi;
}

// turn the winning candidates into the next frame's token list; clears both table levels
template <int NT>
__device__ void finalize_tokens(Ctx<NT> &c, int n_emit, int n_cand, float cutoff, int *t_state, float *t_cost, int *t_arc,
                                int *t_prev) {
    const DecArgs &a = c.a;
    for (int i = c.tid; i < n_cand; i += NT) {
        const int4 cd = c.cand[i];
        const unsigned long long pk = ((unsigned long long)(unsigned)cd.y << 32) | (unsigned)cd.x;
        const float cost = unord((unsigned)cd.y);
        if (cost < cutoff && tab_val(c, cd.z) == pk) {
            int idx = agg_inc(&c.sh.n_next);
            if (idx < a.tok_cap) {
                t_state[idx] = tab_key(c, cd.z);
                t_cost[idx] = cost;
                t_arc[idx] = cd.x;
                tab_set_tok(c, cd.z, idx);  // one winner per slot: nobody else reads this slot's key any more
                if (i < n_emit) t_prev[idx] = cd.w;
            } else {
                c.sh.error = 3;
                tab_set_tok(c, cd.z, 0);
            }
        }
    }
    __syncthreads();
    for (int i = n_emit + c.tid; i < n_cand; i += NT) {
        const int4 cd = c.cand[i];
        const unsigned long long pk = ((unsigned long long)(unsigned)cd.y << 32) | (unsigned)cd.x;
        const float cost = unord((unsigned)cd.y);
        if (cost < cutoff && tab_val(c, cd.z) == pk) {
            int idx = tab_tok(c, cd.z);
            if (idx < a.tok_cap) t_prev[idx] = -2 - tab_tok(c, cd.w);
        }
    }
    __syncthreads();
    if (c.sh.error == 1 || c.sh.error == 2) {
        // a candidate was dropped (table/candidate overflow): entries may exist that no candidate points at
        for (int i = c.tid; i < a.hash_size; i += NT) {
            c.hkey[i] = kEmpty;
            c.hval[i] = kValMax;
        }
    } else {
        for (int i = c.tid; i < n_cand; i += NT) {
            const int slot = c.cand[i].z;
            if (slot >= kSmemSlots) {
                c.hkey[slot - kSmemSlots] = kEmpty;
                c.hval[slot - kSmemSlots] = kValMax;
            }
        }
    }
    for (int i = c.tid; i < kSmemSlots; i += NT) {
        c.skey[i] = kEmpty;
        c.sval[i] = kValMax;
    }
    __syncthreads();
}
}  // namespace

template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? kDecBlocksPerSM : NT == 512 ? 2 : 1) decode_kernel(DecArgs a) {
    extern __shared__ __align__(16) float s_ll[];  // [npdf floats | level-1 values u64[kSmemSlots] | level-1 keys int[kSmemSlots]]
    __shared__ Shared<NT> sh;
    __shared__ int s_lane;
    const int tid = threadIdx.x;
    const size_t g = (size_t)a.scratch_base + blockIdx.x;
    const int nwin_cap = a.cand_cap / 32 + 2;
    unsigned long long *s_val = reinterpret_cast<unsigned long long *>(s_ll + ((a.out_node.dim + 3) & ~3));
    int *s_key = reinterpret_cast<int *>(s_val + kSmemSlots);
    for (int i = tid; i < kSmemSlots; i += NT) {
        s_key[i] = kEmpty;
        s_val[i] = kValMax;
    }
    Ctx<NT> c{a, sh, s_ll, s_key, s_val,
          a.hash_key + g * a.hash_size, a.hash_val + g * a.hash_size, a.hash_tok + g * a.hash_size,
          a.cand + g * a.cand_cap,
          a.rank + g * a.tok_cap,
          a.sv_pref + g * a.tok_cap, a.sv_a0 + g * a.tok_cap, a.sv_src + g * a.tok_cap, a.win_owner + g * nwin_cap,
          a.sv_cost + g * a.tok_cap,
          tid, tid >> 5, tid & 31, true, (unsigned)a.hash_size - 1};
    unsigned long long cnt_tok = 0, cnt_arc_e = 0, cnt_arc_eps = 0, cnt_new = 0;  // per-thread profiling counters
    const int npdf = a.out_node.dim;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_lane = a.lane_begin + atomicAdd(a.queue, 1);
        __syncthreads();
        const int l = s_lane;
        if (l >= a.lane_end) break;
        const LaneDesc ln = a.lanes[l];
        const int ch = ln.channel;
        const long long clk0 = clock64();
        int max_tok = 0;
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
            c.hmask = (unsigned)a.hash_size - 1;
            c.use_l1 = true;
            __syncthreads();
            if (tid == 0) relax(c, a.g.start, pack(0.f, -1), -1);
            __syncthreads();
            int nc = closure(c, 0, min(sh.n_cand, a.cand_cap), a.beam, &cnt_arc_eps);
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
            max_tok = max(max_tok, n_cur);
            float adaptive_beam = a.beam, best = 0.f, cur_cutoff = INFINITY;
            if (!final_pass) {
                // stage this frame's log-likelihood row
                const float *row = a.out_node.buf + ((size_t)ch * a.out_node.ring +
                                                     (((t_first + fi * a.out_node.step) - a.out_node.t_start) / a.out_node.step & (a.out_node.ring - 1))) * npdf;
                for (int i = tid * 4; i < npdf; i += NT * 4) *reinterpret_cast<float4 *>(s_ll + i) = *reinterpret_cast<const float4 *>(row + i);
                int best_idx = 0;
                cur_cutoff = get_cutoff(c, t_cost, n_cur, &adaptive_beam, &best, &best_idx);
                // seed the running minimum with the best token's own arcs (as LatticeFasterDecoder does), so the
                // loose cutoff used while expanding is already close to the final one and few arcs touch the table
                if (c.warp == 0) {
                    const int2 sa = __ldg(&a.g.state_arcs[t_state[best_idx]]);
                    unsigned m = 0xffffffffu;
                    for (int arc = sa.x + c.lane; arc < sa.y; arc += 32) {
                        const int4 av = __ldg(a.g.arcs + arc);
                        const float ac = -best - a.acoustic_scale * s_ll[av.z];
                        m = min(m, ford(best + ac + __int_as_float(av.x)));
                    }
                    for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
                    if (c.lane == 0) {
                        sh.min_ord = m;
                        sh.n_cand = 0;
                        sh.n_next = 0;
                    }
                }
            } else if (tid == 0) {
                sh.min_ord = 0xffffffffu;
                sh.n_cand = 0;
                sh.n_next = 0;
            }
            // ---- pass A: per-warp counts of survivors, survivors with out-arcs, and out-arcs ----
            const int span = ((n_cur + (NT / 32) - 1) / (NT / 32) + 31) & ~31;
            const int wbeg = min(c.warp * span, n_cur), wend = min(wbeg + span, n_cur);
            {
                int cnt = 0, cexp = 0, degsum = 0;
                for (int i0 = wbeg; i0 < wend; i0 += 32) {
                    const int i = i0 + c.lane;
                    const bool f = i < wend && t_cost[i] <= cur_cutoff;
                    int deg = 0;
                    if (f && !final_pass) {
                        const int2 sa = __ldg(&a.g.state_arcs[t_state[i]]);
                        deg = sa.y - sa.x;
                    }
                    cnt += __popc(__ballot_sync(0xffffffffu, f));
                    cexp += __popc(__ballot_sync(0xffffffffu, deg > 0));
                    degsum += deg;
                }
                for (int o = 16; o; o >>= 1) degsum += __shfl_xor_sync(0xffffffffu, degsum, o);
                if (c.lane == 0) {
                    sh.warp_cnt[c.warp] = cnt;
                    sh.warp_exp[c.warp] = cexp;
                    sh.warp_deg[c.warp] = degsum;
                }
            }
            __syncthreads();
            int rbase = 0, ebase = 0, abase = 0, n_surv = 0, n_exp = 0, n_arcs = 0;
            for (int w = 0; w < (NT / 32); w++) {
                if (w < c.warp) {
                    rbase += sh.warp_cnt[w];
                    ebase += sh.warp_exp[w];
                    abase += sh.warp_deg[w];
                }
                n_surv += sh.warp_cnt[w];
                n_exp += sh.warp_exp[w];
                n_arcs += sh.warp_deg[w];
            }
            {
                // table window: room for every emitting arc's target plus closure growth, at load factor <= 1/3
                unsigned want = 3u * (unsigned)n_arcs + 1024u, win = 4096;
                while (win < want && win < (unsigned)a.hash_size) win <<= 1;
                c.hmask = min(win, (unsigned)a.hash_size) - 1;
                c.use_l1 = n_arcs <= kSmemSlots + kSmemSlots / 2;  // heavy frames would only collide in level 1
            }
            if (n_arcs > a.cand_cap) {  // more emitting arcs than candidate slots: flag and truncate
                if (tid == 0) sh.error = 6;
                n_arcs = a.cand_cap;
            }
            const bool log_ok = frame <= a.max_frames && log_count + n_surv <= a.log_cap;
            if (!log_ok && tid == 0) sh.error = 4;
            // ---- pass B: stable ranks; compact the expandable survivors with their arc prefix; window owners ----
            for (int i0 = wbeg; i0 < wend; i0 += 32) {
                const int i = i0 + c.lane;
                const bool f = i < wend && t_cost[i] <= cur_cutoff;
                int deg = 0, a0 = 0;
                float cost = 0.f;
                if (f) {
                    cost = t_cost[i];
                    if (!final_pass) {
                        const int2 sa = __ldg(&a.g.state_arcs[t_state[i]]);
                        a0 = sa.x;
                        deg = sa.y - a0;
                    }
                }
                const unsigned bal = __ballot_sync(0xffffffffu, f), bexp = __ballot_sync(0xffffffffu, deg > 0);
                const int r = rbase + __popc(bal & lanemask_lt());
                if (i < wend) c.rank[i] = f ? r : -1;
                int incl = deg;
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (c.lane >= o) incl += v;
                }
                if (deg > 0) {
                    const int e = ebase + __popc(bexp & lanemask_lt());
                    const int p = abase + incl - deg;
                    c.sv_pref[e] = p;
                    c.sv_a0[e] = a0;
                    c.sv_cost[e] = cost;
                    c.sv_src[e] = log_count + r;
                    for (int w = (p + 31) >> 5; w <= (p + deg - 1) >> 5 && w < nwin_cap; w++) c.win_owner[w] = e;
                }
                rbase += __popc(bal);
                ebase += __popc(bexp);
                abase += __shfl_sync(0xffffffffu, incl, 31);
            }
            __syncthreads();
            // ---- pass C1: token log of the survivors (prev of an epsilon-created token = a survivor of this frame) ----
            if (log_ok) {
                for (int i = tid; i < n_cur; i += NT) {
                    const int r = c.rank[i];
                    if (r < 0) continue;
                    const int li = log_count + r;
                    int pv = t_prev[i];
                    if (pv <= -2) {
                        const int pr = c.rank[-2 - pv];
                        if (pr < 0) sh.error = 5;  // only possible with negative epsilon weights
                        pv = pr < 0 ? -1 : log_count + pr;
                    }
                    log_prev[li] = pv;
                    log_arc[li] = t_arc[i];
                    log_cost[li] = t_cost[i];
                    if (log_state) log_state[li] = t_state[i];
                }
                if (tid == 0) {
                    frame_off[frame] = log_count;
                    frame_off[frame + 1] = log_count + n_surv;
                }
            }
            cnt_tok += tid == 0 ? (unsigned)n_surv : 0u;
            log_count += log_ok ? n_surv : 0;
            if (final_pass) {
                __syncthreads();
                break;
            }
            // ---- pass C2: emitting arcs, one 32-arc window per warp iteration ----
            cnt_arc_e += tid == 0 ? (unsigned)n_arcs : 0u;
            {
                const float cost_offset = -best;
                const int nwin = (n_arcs + 31) >> 5;
                int *own = sh.own[c.warp];
                for (int w = c.warp; w < nwin; w += (NT / 32)) {
                    const int j = (w << 5) + c.lane;
                    const int q0 = c.win_owner[w];
                    const int my = q0 + c.lane;
                    int pref = INT_MAX, a0 = 0, src = 0;
                    float cost = 0.f;
                    if (my < n_exp) {
                        pref = c.sv_pref[my];
                        a0 = c.sv_a0[my];
                        cost = c.sv_cost[my];
                        src = c.sv_src[my];
                    }
                    own[c.lane] = 0;
                    __syncwarp();
                    const int rel = pref - (w << 5);
                    if (c.lane > 0 && rel >= 0 && rel < 32) own[rel] = c.lane;  // prefixes are strictly increasing
                    __syncwarp();
                    int o = own[c.lane];
                    for (int d = 1; d < 32; d <<= 1) {
                        int v = __shfl_up_sync(0xffffffffu, o, d);
                        if (c.lane >= d) o = max(o, v);
                    }
                    const int opref = __shfl_sync(0xffffffffu, pref, o);
                    const int oa0 = __shfl_sync(0xffffffffu, a0, o);
                    const float ocost = __shfl_sync(0xffffffffu, cost, o);
                    const int osrc = __shfl_sync(0xffffffffu, src, o);
                    const bool valid = j < n_arcs;
                    float tot = INFINITY;
                    int arc = 0, next = 0;
                    if (valid) {
                        arc = oa0 + (j - opref);
                        const int4 av = __ldg(a.g.arcs + arc);
                        next = av.y;
                        const float ac = cost_offset - a.acoustic_scale * s_ll[av.z];
                        tot = ocost + ac + __int_as_float(av.x);
                    }
                    unsigned m = ford(tot);
                    for (int d = 16; d; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
                    if (c.lane == 0) atomicMin(&sh.min_ord, m);
                    __syncwarp();
                    const float cut = unord(*(volatile unsigned *)&sh.min_ord) + adaptive_beam;  // loose (>= final) cutoff
                    if (valid && tot < cut) relax(c, next, pack(tot, arc), osrc);
                    __syncwarp();
                }
            }
            __syncthreads();
            const float next_cutoff = unord(sh.min_ord) + adaptive_beam;
            const int n_emit = min(sh.n_cand, a.cand_cap);
            __syncthreads();
            const int nc = closure(c, 0, n_emit, next_cutoff, &cnt_arc_eps);
            finalize_tokens(c, n_emit, nc, next_cutoff, n_state, n_cost, n_arc, n_prev);
            n_cur = min(sh.n_next, a.tok_cap);
            cnt_new += tid == 0 ? (unsigned)n_cur : 0u;
            parity ^= 1;
            frame++;
        }
        // ---- stream end: best token (cost + final, ties by state id) and on-device traceback ----
        if (ln.last) {
            const int lo = frame <= a.max_frames ? frame_off[frame] : 0, hi = log_count;
            const int *t_state = a.tok_state + tbase + (size_t)parity * a.tok_cap;
            // survivors of the final pass are all tokens, logged in list order: log index = lo + i
            unsigned long long bw[2] = {kValMax, kValMax};
            for (int i = tid; i < n_cur && lo + i < hi; i += NT) {
                int s = t_state[i];
                float cst = log_cost[lo + i];
                float fc = __ldg(a.g.final_cost + s);
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
                for (int w = 1; w < (NT / 32); w++) v = min(v, sh.red_ull[w]);
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
                for (int i = tid; i < n_cur && lo + i < hi; i += NT) {
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
            if (a.counters) {  // lane-level balance: sum and max of the cycles one lane took in this launch
                const unsigned long long cyc = (unsigned long long)(clock64() - clk0);
                atomicAdd(a.counters + 4, cyc);
                atomicMax(a.counters + 5, cyc);
                atomicMax(a.counters + 6, (unsigned long long)max_tok);
                atomicAdd(a.counters + 7, 1ull);
            }
            if (a.lane_load) a.lane_load[l] = max_tok;  // fed back to the batcher: heavy streams are grouped together
        }
    }
    if (a.counters) {
        for (int o = 16; o; o >>= 1) cnt_arc_eps += __shfl_xor_sync(0xffffffffu, cnt_arc_eps, o);
        if (tid == 0) {
            atomicAdd(a.counters + 0, cnt_tok);
            atomicAdd(a.counters + 1, cnt_arc_e);
            atomicAdd(a.counters + 3, cnt_new);
        }
        if ((tid & 31) == 0 && cnt_arc_eps) atomicAdd(a.counters + 2, cnt_arc_eps);
    }
}

extern "C" int vbk_decode_max_grid(int device) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    return sms * kDecBlocksPerSM;
}

// heavy = threads per CTA (256 / 512 / 1024): 1024-thread CTAs (one per SM) for batches whose lanes carry thousands of tokens per frame — the
// per-frame critical path of such a lane is what bounds the step, so it gets 4x the threads.
extern "C" cudaError_t vbk_decode(const DecArgs *a, int heavy, cudaStream_t s) {
    const int n = a->lane_end - a->lane_begin;
    if (n <= 0) return cudaSuccess;
    int smem = ((a->out_node.dim + 3) & ~3) * 4 + kSmemSlots * 12;
    static int sms[16] = {};
    // the opt-in shared-memory size is a per-device function attribute: set it once per (variant, device, size)
    static int done[3][16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int v = heavy >= 1024 ? 2 : heavy >= 512 ? 1 : 0;
    if (dev >= 16) return cudaErrorInvalidDevice;
    if (!sms[dev]) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    int grid = sms[dev] * (v == 2 ? 1 : v == 1 ? 2 : kDecBlocksPerSM);
    if (grid > n) grid = n;
    if (grid > a->grid - a->scratch_base) grid = a->grid - a->scratch_base;
    if (grid <= 0) return cudaErrorInvalidValue;
    if (done[v][dev] < smem) {
        cudaError_t e = v == 2   ? cudaFuncSetAttribute(decode_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                        : v == 1 ? cudaFuncSetAttribute(decode_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                 : cudaFuncSetAttribute(decode_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        done[v][dev] = smem;
    }
    if (v == 2) decode_kernel<1024><<<grid, 1024, smem, s>>>(*a);
    else if (v == 1) decode_kernel<512><<<grid, 512, smem, s>>>(*a);
    else decode_kernel<256><<<grid, 256, smem, s>>>(*a);
    return cudaGetLastError();
}

}  // namespace vb
