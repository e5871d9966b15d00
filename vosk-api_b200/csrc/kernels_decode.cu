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
//
// Expansion is staged: the gather pass only loads (two 32-arc windows in flight per warp) and parks every arc below
// the running cutoff as a candidate record; once the frame's minimum is known the insertion pass touches the table
// only with arcs below the FINAL cutoff (thread per record, independent iterations).
//
// Lattice generation (lattice=1; CudaDecoder's per-state extra_prev_tokens / LatticeFasterDecoder's ForwardLinks):
// every arc below the final cutoff whose cost is within lattice_beam of its destination state's best cost is logged as
// a link {src token, dst token, arc, acoustic cost}.  At stream end lattice_prune_kernel restates
// PruneForwardLinksFinal / PruneForwardLinks / PruneTokensForFrame: backward extra-cost propagation (epsilon links of
// a frame to their fixed point, then the emitting links into it), links and tokens above lattice_beam dropped, the
// survivors renumbered and compacted into the raw lattice the host reads.
#include <algorithm>
#include <cfloat>
#include <climits>

#include "vb_kernels.h"

namespace vb {

namespace {
#ifndef VB_DEC_LIGHT_PER_SM
#define VB_DEC_LIGHT_PER_SM 3
#endif
constexpr int kDecBlocksPerSM = VB_DEC_LIGHT_PER_SM;   // 256-thread variant; the 1024-thread variant (heavy lanes) runs 1 per SM
// level-1 (shared memory) table entries per CTA (Ctx::l1_slots, a power of two chosen per launch): as many as the shared memory of an
// SM allows at the variant's residency (12 bytes per entry beside the log-likelihood row) — one 1024-thread CTA, two 512-thread CTAs
// or three 256-thread CTAs per SM.  A lane's
// search runs on one SM and is bounded by that SM's path to L2 (~64 B / clock): every table access kept in shared memory is a 32-byte
// sector less on that path
constexpr int kMaxSlots = 16384;
constexpr unsigned long long kValMax = ~0ull;
constexpr int kEmpty = -1;
constexpr int kAltFlag = 0x40000000;  // candidate did not improve its state's best word; kept as a lattice link only
constexpr unsigned kInfBits = 0x7f800000u;

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
    int n_cand, n_next, n_links, n_work, error;
    int warp_cnt[(NT / 32)], warp_exp[(NT / 32)], warp_deg[(NT / 32)];
    unsigned hist[256];
    unsigned sel_prefix, sel_mask;
    int sel_k;
    unsigned red_u[(NT / 32)];
    unsigned long long red_ull[(NT / 32)];
    long long phase[8];  // cycles per phase of this CTA (tid 0), flushed to counters[16..23] of the heavy / [24..31] of the light variants
    long long lphase[8];  // the same for the lane in hand: the slowest lane's go to counters[32 + 8 * tier ..]
    int own[(NT / 32)][2][32];  // per-warp marker arrays of the arc-window owner scan (two windows in flight)
    // what GetCutoff needs to know about a token list, gathered while the list is written (two sets: the frame in hand reads one,
    // its finalize pass fills the other): smallest state among the cheapest tokens, #{cost < best + beam}, #{cost <= best + beam}
    int nx_state[2], nx_lt[2], nx_le[2];
    unsigned nx_best[2];  // ordered bits of the list's minimum cost
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
    int4 *cand;                // {arc (packed lo), ordered cost (packed hi), slot | kAltFlag (-1 = dead; state before insertion), src}
    int *cand_next;            // destination state of a candidate
    int *work;                 // epsilon-closure work list (candidate indices)
    int *rank;
    int *sv_pref, *sv_a0, *sv_src, *win_owner;
    float *sv_cost;
    int tid, warp, lane;
    int l1_slots;    // level-1 table size of this launch
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

// Table slots: [0, c.l1_slots) = shared-memory level, c.l1_slots + g = global level.  A state lives in level 1 iff
// a free or matching slot existed within kProbe1 probes at its first insertion; slots are never freed inside a
// frame, so every later lookup of the same state takes the same decision.
constexpr int kProbe1 = 16;
constexpr int kSmallRound = 32;   // closure rounds of at most this many entries are left to one warp
template <int NT>
__device__ __forceinline__ unsigned long long tab_val(const Ctx<NT> &c, int slot) {
    return slot < c.l1_slots ? *(volatile unsigned long long *)(c.sval + slot) : __ldcg(c.hval + (slot - c.l1_slots));
}
template <int NT>
__device__ __forceinline__ int tab_key(const Ctx<NT> &c, int slot) {
    return slot < c.l1_slots ? *(volatile int *)(c.skey + slot) : __ldcg(c.hkey + (slot - c.l1_slots));
}
// after the winners are known the key field of a level-1 slot is reused for the token index
template <int NT>
__device__ __forceinline__ void tab_set_tok(const Ctx<NT> &c, int slot, int idx) {
    if (slot < c.l1_slots) c.skey[slot] = idx; else c.htok[slot - c.l1_slots] = idx;
}
template <int NT>
__device__ __forceinline__ int tab_tok(const Ctx<NT> &c, int slot) {
    return slot < c.l1_slots ? *(volatile int *)(c.skey + slot) : __ldcg(c.htok + (slot - c.l1_slots));
}

// claims / finds the table slot of a state and folds pk into its best word; returns the slot (-1: table full)
template <int NT>
__device__ __forceinline__ int table_insert(Ctx<NT> &c, int state, unsigned long long pk, unsigned long long *old_out) {
    const unsigned hash = ((unsigned)state * 2654435761u) >> 7;
    if (c.use_l1) {
        unsigned h = hash & (c.l1_slots - 1);
#pragma unroll 1
        for (int p = 0; p < kProbe1; p++) {
            int prev = atomicCAS(c.skey + h, kEmpty, state);
            if (prev == kEmpty || prev == state) {
                *old_out = atomicMin(c.sval + h, pk);
                return (int)h;
            }
            h = (h + 1) & (c.l1_slots - 1);
        }
    }
    const unsigned mask = c.hmask;
    unsigned g = hash & mask;
    int probes = 0;
    for (;;) {
        int prev = atomicCAS(c.hkey + g, kEmpty, state);
        if (prev == kEmpty || prev == state) break;
        g = (g + 1) & mask;
        if (++probes > (int)mask) {
            c.sh.error = 1;
            return -1;
        }
    }
    *old_out = atomicMin(c.hval + g, pk);
    return c.l1_slots + (int)g;
}

// 0: pk became the state's best word (a token candidate); kAltFlag: not the best, but within lattice_beam of the best
// seen so far (the best only decreases, so nothing the final test keeps is lost here); -1: drop
template <int NT>
__device__ __forceinline__ int keep_flags(const Ctx<NT> &c, unsigned long long pk, unsigned long long old, float cost) {
    if (pk < old) return 0;
    if (c.a.lattice && cost - unord((unsigned)(old >> 32)) <= c.a.lattice_beam) return kAltFlag;
    return -1;
}

// epsilon-closure insertion: appends a candidate record (src = index of the generating candidate)
template <int NT>
__device__ __forceinline__ void relax(Ctx<NT> &c, int state, unsigned long long pk, int src, bool has_eps) {
    unsigned long long old;
    const int slot = table_insert(c, state, pk, &old);
    if (slot < 0) return;
    const int fl = keep_flags(c, pk, old, unord((unsigned)(pk >> 32)));
    if (fl < 0) return;
    int idx = agg_inc(&c.sh.n_cand);
    if (idx < c.a.cand_cap) {
        c.cand[idx] = make_int4((int)(unsigned)pk, (int)(unsigned)(pk >> 32), slot | fl, src);
        c.cand_next[idx] = state;
        if (fl == 0 && has_eps) c.work[agg_inc(&c.sh.n_work)] = idx;  // at most one entry per candidate: never beyond cand_cap
    } else {
        c.sh.error = 2;
    }
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
// when they fall on the right side of best+beam, which two counts decide:
//   kth(max_active) <  beam_cutoff  <=>  #{cost <  beam_cutoff} >  max_active
//   kth(min_active) >  beam_cutoff  <=>  #{cost <= beam_cutoff} <= min_active
// so the exact radix select runs only on the frames where it changes the result.  The minimum and the two counts are not
// computed here: the pass that wrote the token list knew the minimum beforehand (the cheapest emitting candidate always
// becomes a token, and epsilon arcs — weights >= 0 — add to a cost) and counted while it wrote (finalize_tokens), which
// takes two sweeps over the tokens and four block barriers out of every frame.
template <int NT>
__device__ float cutoff_from_stats(Ctx<NT> &c, const float *cost, int n, float best, int n_lt, int n_le, float *adaptive_beam) {
    const DecArgs &a = c.a;
    const float beam_cutoff = best + a.beam;
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

// Epsilon closure to the fixed point; returns the total candidate count.  The work list holds the candidates that became
// their state's best word AND whose state has epsilon arcs (a static property flagged in the arc records), so a round
// costs what it expands, not a scan of all candidates.  Only an entry that still is its state's best word is expanded; if
// it was superseded, the better candidate sits further down the list.  Epsilon out-degrees are tiny: thread per entry.
template <int NT>
__device__ int closure(Ctx<NT> &c, float cutoff, unsigned long long *arcs_seen) {
    const DecArgs &a = c.a;
    int lo = 0, hi = min(c.sh.n_work, a.cand_cap);
    __syncthreads();
    auto expand = [&](int w) {
        const int i = c.work[w];
        const int4 cd = c.cand[i];
        const unsigned long long pk = ((unsigned long long)(unsigned)cd.y << 32) | (unsigned)cd.x;
        const float cost = unord((unsigned)cd.y);
        if (cost < cutoff && tab_val(c, cd.z) == pk) {
            const int s = c.cand_next[i];
            const int a0 = __ldg(&a.g.state_arcs[s].y), a1 = __ldg(&a.g.state_arcs[s + 1].x);
            *arcs_seen += (unsigned)(a1 - a0);
            for (int arc = a0; arc < a1; arc++) {
                const int4 av = __ldg(a.g.arcs + arc);
                const float tot = cost + __int_as_float(av.x);
                if (tot < cutoff) relax(c, av.y, pack(tot, arc), i, (av.w & kNextHasEps) != 0);
            }
        }
    };
    while (lo < hi) {
        if (hi - lo <= kSmallRound) {
            // the tail of the closure — a few dozen entries per round, each a chain of dependent loads — is run to its fixed point by
            // one warp with warp barriers, instead of two block barriers per round with every other warp waiting at them
            if (c.warp == 0) {
                while (lo < hi) {
                    for (int w = lo + c.lane; w < hi; w += 32) expand(w);
                    __syncwarp();
                    lo = hi;
                    hi = min(*(volatile int *)&c.sh.n_work, a.cand_cap);
                    __syncwarp();
                }
            }
            break;
        }
        for (int w = lo + c.tid; w < hi; w += NT) expand(w);
        __syncthreads();
        lo = hi;
        hi = min(c.sh.n_work, a.cand_cap);
        __syncthreads();
    }
    __syncthreads();
    return min(c.sh.n_cand, a.cand_cap);
}

// turn the winning candidates into the next frame's token list, log the lattice links, clear both table levels.
// Links are appended at links[link_base ...] with token-LIST indices of the new frame in the destination (and, for
// epsilon links, source) field; the next frame's pass translates them to log indices once the survivors are ranked.
template <int NT>
__device__ void finalize_tokens(Ctx<NT> &c, int n_emit, int n_cand, float cutoff, float cost_offset, int4 *links, int link_base,
                                int *t_state, float *t_cost, int *t_arc, int *t_prev, float best_next, int stat_set) {
    const DecArgs &a = c.a;
    // best_next = the minimum cost of the list being written (known beforehand); the counts GetCutoff needs of it are taken here
    const float beam_cutoff_next = best_next + a.beam;
    int my_lt = 0, my_le = 0, my_state = INT_MAX;
    for (int i = c.tid; i < n_cand; i += NT) {
        const int4 cd = c.cand[i];
        if (cd.z < 0 || (cd.z & kAltFlag)) continue;
        const unsigned long long pk = ((unsigned long long)(unsigned)cd.y << 32) | (unsigned)cd.x;
        const float cost = unord((unsigned)cd.y);
        if (cost < cutoff && tab_val(c, cd.z) == pk) {
            int idx = agg_inc(&c.sh.n_next);
            if (idx < a.tok_cap) {
                const int st = c.cand_next[i];
                my_lt += cost < beam_cutoff_next;
                my_le += cost <= beam_cutoff_next;
                if (cost == best_next) my_state = min(my_state, st);
                t_state[idx] = st;
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
    {
        for (int o = 16; o; o >>= 1) {
            my_lt += __shfl_xor_sync(0xffffffffu, my_lt, o);
            my_le += __shfl_xor_sync(0xffffffffu, my_le, o);
            my_state = min(my_state, __shfl_xor_sync(0xffffffffu, my_state, o));
        }
        if (c.lane == 0) {
            if (my_le) {
                atomicAdd(&c.sh.nx_lt[stat_set], my_lt);
                atomicAdd(&c.sh.nx_le[stat_set], my_le);
            }
            if (my_state != INT_MAX) atomicMin(&c.sh.nx_state[stat_set], my_state);
        }
        if (c.tid == 0) c.sh.nx_best[stat_set] = ford(best_next);
    }
    __syncthreads();
    const int first = a.lattice ? 0 : n_emit;  // without lattice generation only the epsilon winners need this pass
    for (int i = first + c.tid; i < n_cand; i += NT) {
        const int4 cd = c.cand[i];
        if (cd.z < 0) continue;
        const int slot = cd.z & ~kAltFlag;
        const unsigned long long pk = ((unsigned long long)(unsigned)cd.y << 32) | (unsigned)cd.x;
        const float cost = unord((unsigned)cd.y);
        if (!(cost < cutoff)) continue;
        const unsigned long long best = tab_val(c, slot);
        const bool winner = !(cd.z & kAltFlag) && best == pk;
        const bool is_eps = i >= n_emit;
        int src_tok = 0;
        bool src_ok = true;
        if (is_eps) {  // the generating candidate must still be its state's best word (else this record is stale)
            const int4 sc = c.cand[cd.w];
            const unsigned long long spk = ((unsigned long long)(unsigned)sc.y << 32) | (unsigned)sc.x;
            src_ok = sc.z >= 0 && !(sc.z & kAltFlag) && tab_val(c, sc.z) == spk;
            // the source STATE's token (a superseded twin of equal cost still names the right predecessor token)
            if (sc.z >= 0) src_tok = tab_tok(c, sc.z & ~kAltFlag);
        }
        const int dst_tok = tab_tok(c, slot);
        if (winner && is_eps && dst_tok < a.tok_cap) t_prev[dst_tok] = -2 - src_tok;
        if (a.lattice && src_ok && cd.x >= 0 && cost - unord((unsigned)(best >> 32)) <= a.lattice_beam) {
            float ac = 0.f;
            if (!is_eps) ac = cost_offset - __fmul_rn(a.acoustic_scale, c.ll[__ldg(a.g.arcs + cd.x).z]);
            const int k = link_base + agg_inc(&c.sh.n_links);
            if (k < a.link_cap) links[k] = make_int4(is_eps ? -2 - src_tok : cd.w, dst_tok | (is_eps ? kEpsLinkFlag : 0), cd.x, __float_as_int(ac));
            else c.sh.error = 7;
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
            const int z = c.cand[i].z;
            const int slot = z & ~kAltFlag;
            if (z < 0) continue;
            if (slot >= c.l1_slots) {
                c.hkey[slot - c.l1_slots] = kEmpty;
                c.hval[slot - c.l1_slots] = kValMax;
            } else {  // level 1 is cleared slot by slot too: a full sweep per frame costs more than the frame's few entries
                c.skey[slot] = kEmpty;
                c.sval[slot] = kValMax;
            }
        }
    }
    if (c.sh.error == 1 || c.sh.error == 2)
        for (int i = c.tid; i < c.l1_slots; i += NT) {
            c.skey[i] = kEmpty;
            c.sval[i] = kValMax;
        }
    // (no barrier here: every caller runs into one — the top of the frame loop / of the lane loop — before the tables or the
    // shared counters are touched again)
}
}  // namespace

template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? kDecBlocksPerSM : NT == 512 ? 2 : 1) decode_kernel(DecArgs a) {
    extern __shared__ __align__(16) float s_ll[];  // [npdf floats | level-1 values u64[l1_slots] | level-1 keys int[l1_slots]]
    const int l1_slots = a.l1_slots;
    __shared__ Shared<NT> sh;
    __shared__ int s_lane;
    const int tid = threadIdx.x;
    const size_t g = (size_t)a.scratch_base + blockIdx.x;
    const int nwin_cap = a.cand_cap / 32 + 2;
    unsigned long long *s_val = reinterpret_cast<unsigned long long *>(s_ll + ((a.out_node.dim + 3) & ~3));
    int *s_key = reinterpret_cast<int *>(s_val + l1_slots);
    for (int i = tid; i < l1_slots; i += NT) {
        s_key[i] = kEmpty;
        s_val[i] = kValMax;
    }
    Ctx<NT> c{a, sh, s_ll, s_key, s_val,
          a.hash_key + g * a.hash_size, a.hash_val + g * a.hash_size, a.hash_tok + g * a.hash_size,
          a.cand + g * a.cand_cap, a.cand_next + g * a.cand_cap, a.eps_work + g * a.cand_cap,
          a.rank + g * a.tok_cap,
          a.sv_pref + g * a.tok_cap, a.sv_a0 + g * a.tok_cap, a.sv_src + g * a.tok_cap, a.win_owner + g * nwin_cap,
          a.sv_cost + g * a.tok_cap,
          tid, tid >> 5, tid & 31, l1_slots, true, (unsigned)a.hash_size - 1};
    long long tph = 0;
    if (tid == 0)
        for (int k = 0; k < 8; k++) sh.phase[k] = 0;
#define VB_PHASE(k)                                   \
    if (tid == 0 && a.counters) {                     \
        const long long t_ = clock64();               \
        sh.phase[k] += t_ - tph;                      \
        sh.lphase[k] += t_ - tph;                     \
        tph = t_;                                     \
    }
    unsigned long long cnt_tok = 0, cnt_arc_e = 0, cnt_arc_eps = 0, cnt_new = 0, cnt_stage = 0, cnt_links = 0;  // per-thread profiling counters
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
        if (tid == 0)
            for (int k = 0; k < 8; k++) sh.lphase[k] = 0;
        int max_tok = 0;
        DecChannelState *cs = a.cs + ch;
        const size_t tbase = (size_t)ch * 2 * a.tok_cap;
        int *log_prev = a.log_prev + (size_t)ch * a.log_cap;
        int *log_arc = a.log_arc + (size_t)ch * a.log_cap;
        float *log_cost = a.log_cost + (size_t)ch * a.log_cap;
        int *log_state = a.log_state ? a.log_state + (size_t)ch * a.log_cap : nullptr;
        int *frame_off = a.log_frame_off + (size_t)ch * (a.max_frames + 2);
        int4 *links = a.lattice ? a.links + (size_t)ch * a.link_cap : nullptr;
        int *link_off = a.lattice ? a.link_off + (size_t)ch * (a.max_frames + 3) : nullptr;
        __syncthreads();
        if (tid == 0) sh.error = ln.dec_first ? 0 : cs->error;
        int n_cur, parity, frame, log_count, link_count = 0, seg_begin = 0;
        if (ln.dec_first) {
            // InitDecoding: start token + epsilon closure with cutoff = beam
            if (tid == 0) {
                sh.n_cand = 0;
                sh.n_next = 0;
                sh.n_links = 0;
                sh.n_work = 0;
                sh.nx_state[0] = INT_MAX;
                sh.nx_lt[0] = 0;
                sh.nx_le[0] = 0;
                if (link_off) link_off[0] = 0;
            }
            c.hmask = (unsigned)a.hash_size - 1;
            c.use_l1 = true;
            __syncthreads();
            if (tid == 0) relax(c, a.g.start, pack(0.f, -1), -1, true);
            __syncthreads();
            int nc = closure(c, a.beam, &cnt_arc_eps);
            // (the start token costs 0 and epsilon weights are not negative: the list's minimum is 0)
            finalize_tokens(c, 1, nc, INFINITY, 0.f, links, 0, a.tok_state + tbase, a.tok_cost + tbase, a.tok_arc + tbase, a.tok_prev + tbase, 0.f, 0);
            __syncthreads();
            n_cur = min(sh.n_next, a.tok_cap);
            link_count = a.lattice ? min(sh.n_links, a.link_cap) : 0;
            parity = 0;
            frame = 0;
            log_count = 0;
        } else {
            n_cur = cs->n_cur;
            parity = cs->parity;
            frame = cs->frame;
            log_count = cs->log_count;
            link_count = cs->link_count;
            if (tid == 0) {  // the statistics of the token list the previous chunk left
                sh.nx_state[0] = cs->nx_state;
                sh.nx_lt[0] = cs->nx_lt;
                sh.nx_le[0] = cs->nx_le;
                sh.nx_best[0] = cs->nx_best;
            }
            if (link_off) seg_begin = link_off[min(frame, a.max_frames + 1)];
        }
        const int nf = a.out_table[l].n_rows;
        const int t_first = a.out_table[l].t_begin;
        const int total_frames = nf + (ln.dec_last ? 1 : 0);  // the extra pass logs the final frame's tokens
        int sset = 0;  // which set of list statistics describes the current token list
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
            __syncthreads();  // every thread has read the previous frame's shared counters (n_next, n_links) before they are reset
            if (tid == 0) tph = clock64();
            float adaptive_beam = a.beam, best = 0.f, cur_cutoff = INFINITY;
            if (!final_pass) {
                // stage this frame's log-likelihood row
                const float *row = a.out_node.buf + ((size_t)ch * a.out_node.ring +
                                                     (((t_first + fi * a.out_node.step) - a.out_node.t_start) / a.out_node.step & (a.out_node.ring - 1))) * npdf;
                for (int i = tid * 4; i < npdf; i += NT * 4) *reinterpret_cast<float4 *>(s_ll + i) = *reinterpret_cast<const float4 *>(row + i);
                best = unord(sh.nx_best[sset]);
                const int best_state = sh.nx_state[sset], n_lt = sh.nx_lt[sset], n_le = sh.nx_le[sset];
                cur_cutoff = cutoff_from_stats(c, t_cost, n_cur, best, n_lt, n_le, &adaptive_beam);
                // seed the running minimum with a best token's own arcs (as LatticeFasterDecoder does), so the
                // loose cutoff used while expanding is already close to the final one and few arcs touch the table
                if (c.warp == 0) {
                    // (no such token only after a capacity overflow dropped it: then the running minimum starts unseeded)
                    const int2 sa = best_state != INT_MAX ? __ldg(&a.g.state_arcs[best_state]) : make_int2(0, 0);
                    unsigned m = 0xffffffffu;
                    for (int arc = sa.x + c.lane; arc < sa.y; arc += 32) {
                        const int4 av = __ldg(a.g.arcs + arc);
                        // (read from the ring row itself: the shared copy is being written by the other warps, and no barrier is spent on it)
                        const float ac = -best - __fmul_rn(a.acoustic_scale, row[av.z]);
                        m = min(m, ford(best + ac + __int_as_float(av.x)));
                    }
                    for (int o = 16; o; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
                    if (c.lane == 0) {
                        sh.min_ord = m;
                        sh.n_cand = 0;
                        sh.n_next = 0;
                        sh.n_links = 0;
                        sh.n_work = 0;
                        sh.nx_state[sset ^ 1] = INT_MAX;  // (the other set: read last at the top of the previous frame)
                        sh.nx_lt[sset ^ 1] = 0;
                        sh.nx_le[sset ^ 1] = 0;
                    }
                }
            } else if (tid == 0) {
                sh.min_ord = 0xffffffffu;
                sh.n_cand = 0;
                sh.n_next = 0;
                sh.n_links = 0;
                sh.n_work = 0;
            }
            VB_PHASE(0)
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
            int rbase, ebase, abase, n_surv, n_exp, n_arcs;
            {
                // exclusive prefix over the warps' counts (at most 32 of them): one value per lane, shuffle scan, this warp's entry
                constexpr int NW = NT / 32;
                int vc = c.lane < NW ? sh.warp_cnt[c.lane] : 0, ve = c.lane < NW ? sh.warp_exp[c.lane] : 0, vd = c.lane < NW ? sh.warp_deg[c.lane] : 0;
                int ic = vc, ie = ve, id = vd;
                for (int o = 1; o < 32; o <<= 1) {
                    const int uc = __shfl_up_sync(0xffffffffu, ic, o), ue = __shfl_up_sync(0xffffffffu, ie, o), ud = __shfl_up_sync(0xffffffffu, id, o);
                    if (c.lane >= o) {
                        ic += uc;
                        ie += ue;
                        id += ud;
                    }
                }
                rbase = __shfl_sync(0xffffffffu, ic - vc, c.warp);
                ebase = __shfl_sync(0xffffffffu, ie - ve, c.warp);
                abase = __shfl_sync(0xffffffffu, id - vd, c.warp);
                n_surv = __shfl_sync(0xffffffffu, ic, 31);
                n_exp = __shfl_sync(0xffffffffu, ie, 31);
                n_arcs = __shfl_sync(0xffffffffu, id, 31);
            }
            {
                // table window: room for every emitting arc's target plus closure growth, at load factor <= 1/3
                unsigned want = 3u * (unsigned)n_arcs + 1024u, win = 4096;
                while (win < want && win < (unsigned)a.hash_size) win <<= 1;
                c.hmask = min(win, (unsigned)a.hash_size) - 1;
                c.use_l1 = n_arcs <= c.l1_slots + c.l1_slots / 2;  // heavy frames would only collide in level 1
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
            VB_PHASE(1)
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
            if (links) {
                // links created by the previous frame point at this frame's tokens by LIST index: now that the survivors
                // are ranked, rewrite them to log indices (a link into a token that was not logged is dropped)
                for (int k = seg_begin + tid; k < link_count; k += NT) {
                    int4 l = links[k];
                    const int r = log_ok ? c.rank[l.y & ~kEpsLinkFlag] : -1;
                    int y = r < 0 ? -1 : ((log_count + r) | (l.y & kEpsLinkFlag));
                    if (l.x <= -2) {
                        const int rs = log_ok ? c.rank[-2 - l.x] : -1;
                        if (rs < 0) y = -1;
                        l.x = log_count + max(rs, 0);
                    }
                    l.y = y;
                    links[k] = l;
                }
                if (tid == 0 && frame <= a.max_frames) link_off[frame + 1] = link_count;
                seg_begin = link_count;
            }
            cnt_tok += tid == 0 ? (unsigned)n_surv : 0u;
            log_count += log_ok ? n_surv : 0;
            if (final_pass) {
                __syncthreads();
                break;
            }
            VB_PHASE(2)
            // ---- pass C2a: gather.  Two 32-arc windows per warp iteration, all loads of a stage issued before their
            // first use; arcs below the running cutoff are parked as candidate records {arc, cost, next state, src} ----
            cnt_arc_e += tid == 0 ? (unsigned)n_arcs : 0u;
            const float cost_offset = -best;
            if (links && tid == 0 && frame <= a.max_frames) a.frame_offset[(size_t)ch * (a.max_frames + 2) + frame] = cost_offset;
            {
                constexpr int NW = NT / 32;
                const int nwin = (n_arcs + 31) >> 5;
                int *ownA = sh.own[c.warp][0], *ownB = sh.own[c.warp][1];
                for (int w0 = c.warp; w0 < nwin; w0 += 2 * NW) {
                    const int w1 = w0 + NW;
                    const bool h1 = w1 < nwin;  // warp-uniform
                    const int qA = c.win_owner[w0], qB = h1 ? c.win_owner[w1] : 0;
                    const int myA = qA + c.lane, myB = qB + c.lane;
                    int prefA = INT_MAX, a0A = 0, srcA = 0, prefB = INT_MAX, a0B = 0, srcB = 0;
                    float costA = 0.f, costB = 0.f;
                    if (myA < n_exp) {
                        prefA = c.sv_pref[myA];
                        a0A = c.sv_a0[myA];
                        costA = c.sv_cost[myA];
                        srcA = c.sv_src[myA];
                    }
                    if (h1 && myB < n_exp) {
                        prefB = c.sv_pref[myB];
                        a0B = c.sv_a0[myB];
                        costB = c.sv_cost[myB];
                        srcB = c.sv_src[myB];
                    }
                    ownA[c.lane] = 0;
                    ownB[c.lane] = 0;
                    __syncwarp();
                    const int relA = prefA - (w0 << 5), relB = prefB - (w1 << 5);
                    if (c.lane > 0 && relA >= 0 && relA < 32) ownA[relA] = c.lane;  // prefixes are strictly increasing
                    if (c.lane > 0 && relB >= 0 && relB < 32) ownB[relB] = c.lane;
                    __syncwarp();
                    int oA = ownA[c.lane], oB = ownB[c.lane];
                    for (int d = 1; d < 32; d <<= 1) {
                        const int vA = __shfl_up_sync(0xffffffffu, oA, d), vB = __shfl_up_sync(0xffffffffu, oB, d);
                        if (c.lane >= d) {
                            oA = max(oA, vA);
                            oB = max(oB, vB);
                        }
                    }
                    const int jA = (w0 << 5) + c.lane, jB = (w1 << 5) + c.lane;
                    const bool validA = jA < n_arcs, validB = h1 && jB < n_arcs;
                    const int arcA = __shfl_sync(0xffffffffu, a0A, oA) + (jA - __shfl_sync(0xffffffffu, prefA, oA));
                    const int arcB = __shfl_sync(0xffffffffu, a0B, oB) + (jB - __shfl_sync(0xffffffffu, prefB, oB));
                    const float ocostA = __shfl_sync(0xffffffffu, costA, oA), ocostB = __shfl_sync(0xffffffffu, costB, oB);
                    const int osrcA = __shfl_sync(0xffffffffu, srcA, oA), osrcB = __shfl_sync(0xffffffffu, srcB, oB);
                    int4 avA = make_int4(0, 0, 0, 0), avB = make_int4(0, 0, 0, 0);
                    if (validA) avA = __ldg(a.g.arcs + arcA);
                    if (validB) avB = __ldg(a.g.arcs + arcB);
                    float totA = INFINITY, totB = INFINITY;
                    if (validA) totA = ocostA + (cost_offset - __fmul_rn(a.acoustic_scale, s_ll[avA.z])) + __int_as_float(avA.x);
                    if (validB) totB = ocostB + (cost_offset - __fmul_rn(a.acoustic_scale, s_ll[avB.z])) + __int_as_float(avB.x);
                    unsigned m = min(ford(totA), ford(totB));
                    for (int d = 16; d; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
                    if (c.lane == 0) atomicMin(&sh.min_ord, m);
                    __syncwarp();
                    const float cut = unord(*(volatile unsigned *)&sh.min_ord) + adaptive_beam;  // loose (>= final) cutoff
                    if (validA && totA < cut) {
                        const int idx = agg_inc(&sh.n_cand);
                        if (idx < a.cand_cap) c.cand[idx] = make_int4(arcA, (int)ford(totA), avA.y | (avA.w & kNextHasEps), osrcA);
                    }
                    if (validB && totB < cut) {
                        const int idx = agg_inc(&sh.n_cand);
                        if (idx < a.cand_cap) c.cand[idx] = make_int4(arcB, (int)ford(totB), avB.y | (avB.w & kNextHasEps), osrcB);
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            const float next_cutoff = unord(sh.min_ord) + adaptive_beam;
            if (tid == 0 && sh.n_cand > a.cand_cap) sh.error = 2;
            const int n_emit = min(sh.n_cand, a.cand_cap);
            cnt_stage += tid == 0 ? (unsigned)n_emit : 0u;
            VB_PHASE(3)
            // ---- pass C2b: insertion of the records below the FINAL cutoff (thread per record) ----
            for (int i = tid; i < n_emit; i += NT) {
                const int4 cd = c.cand[i];
                const float cost = unord((unsigned)cd.y);
                int z = -1;
                if (cost < next_cutoff) {
                    const unsigned long long pk = ((unsigned long long)(unsigned)cd.y << 32) | (unsigned)cd.x;
                    const int state = cd.z & ~kArcFlagMask;
                    unsigned long long old;
                    const int slot = table_insert(c, state, pk, &old);
                    if (slot >= 0) {
                        const int fl = keep_flags(c, pk, old, cost);
                        if (fl >= 0) {
                            z = slot | fl;
                            c.cand_next[i] = state;
                            if (fl == 0 && (cd.z & kNextHasEps)) c.work[agg_inc(&sh.n_work)] = i;
                        }
                    }
                }
                reinterpret_cast<int *>(c.cand + i)[2] = z;
            }
            __syncthreads();
            VB_PHASE(4)
            const int nc = closure(c, next_cutoff, &cnt_arc_eps);
            VB_PHASE(5)
            finalize_tokens(c, n_emit, nc, next_cutoff, cost_offset, links, link_count, n_state, n_cost, n_arc, n_prev, unord(sh.min_ord), sset ^ 1);
            sset ^= 1;
            VB_PHASE(6)
            if (a.lattice) {
                cnt_links += tid == 0 ? (unsigned)sh.n_links : 0u;
                link_count = min(link_count + sh.n_links, a.link_cap);
            }
            n_cur = min(sh.n_next, a.tok_cap);
            cnt_new += tid == 0 ? (unsigned)n_cur : 0u;
            parity ^= 1;
            frame++;
        }
        // ---- stream end: best token (cost + final, ties by state id) and on-device traceback ----
        if (ln.dec_last) {
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
            cs->link_count = link_count;
            cs->nx_state = sh.nx_state[sset];
            cs->nx_lt = sh.nx_lt[sset];
            cs->nx_le = sh.nx_le[sset];
            cs->nx_best = sh.nx_best[sset];
            cs->error = sh.error;
            if (a.counters) {  // lane-level balance: sum and max of the cycles one lane took in this launch
                const unsigned long long cyc = (unsigned long long)(clock64() - clk0);
                atomicAdd(a.counters + 4, cyc);
                atomicMax(a.counters + 5, cyc);
                atomicMax(a.counters + 6, (unsigned long long)max_tok);
                atomicAdd(a.counters + 7, 1ull);
                constexpr int tier = NT >= 1024 ? 0 : NT >= 512 ? 1 : 2;
                atomicAdd(a.counters + 59 + tier, 1ull);
                if (cyc >= atomicMax(a.counters + 11 + tier, cyc)) {  // the slowest lane of this tier so far, give or take a race: a profile, not a result
                    for (int k = 0; k < 8; k++) a.counters[32 + 8 * tier + k] = (unsigned long long)sh.lphase[k];
                    a.counters[56 + tier] = (unsigned long long)max_tok;
                }
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
            atomicAdd(a.counters + 8, cnt_stage);
            atomicAdd(a.counters + 9, cnt_links);
            for (int k = 0; k < 8; k++) atomicAdd(a.counters + (NT >= 1024 ? 16 : 24) + k, (unsigned long long)sh.phase[k]);
        }
        if ((tid & 31) == 0 && cnt_arc_eps) atomicAdd(a.counters + 2, cnt_arc_eps);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Lattice pruning + compaction at stream end (one CTA per finished lane).
// Restates LatticeFasterDecoder::PruneForwardLinksFinal / PruneForwardLinks(delta = 0) / PruneTokensForFrame:
//   extra(tok of the last frame) = (cost + final) - best_final            (final = 0 if no final state was reached)
//   link_extra = extra(dst) + ((cost(src) + acoustic + graph) - cost(dst)), clamped at 0
//   a link survives iff link_extra <= lattice_beam; extra(tok) = min over its surviving links (inf: token dropped)
// Frames are visited last to first; within a frame the epsilon links are iterated to their fixed point (monotone
// atomicMin on the ordered-float bits of non-negative costs) before the emitting links push into the previous frame.
//
// B200 mapping: one CTA per finished lane sweeps the frames last to first — a chain of ~400 dependent steps, so what counts
// is the latency of one step.  The extra costs of the two frames in play live in shared memory (global only for a frame
// wider than the buffer); each thread holds its share of the NEXT frame's links in registers, requested one frame ahead, so
// that the link log streams in behind the work on the current frame; only the links whose destination is alive (1-2 % of
// the log) go on to touch costs and graph weights.  Survivors are written out during the sweep (frames descending; the block
// turns the list around at the end), the epsilon fixed point costs one barrier per round (__syncthreads_or), and the token
// renumbering is a warp-per-frame pass over per-frame counts instead of a barrier per 256 tokens.
// ------------------------------------------------------------------------------------------------------------
constexpr int kPruneThreads = 512;
constexpr int kPruneRegLinks = 4;     // links of a frame held in registers per thread (the rest of a wider frame is read in place)
constexpr int kPruneSmemToks = 8192;  // tokens of a frame whose extra costs fit the shared-memory buffer

struct PruneFrame {
    const DecArgs *a;
    const float *cost;
    unsigned *ex_cur, *ex_prev;  // extra costs (float bits; monotone under atomicMin as they are non-negative) of frame f / f-1
    int t0, p0;                  // first log index of frame f / f-1
    float lb, off;
};

// one link of frame f against the current extras.  Returns 0: dead, 1: alive (le <= lattice_beam) and stores le.
__device__ __forceinline__ int prune_link(const PruneFrame &fr, const int4 lk, float *le_out) {
    const int dst = lk.y & ~kEpsLinkFlag;
    const float ed = __uint_as_float(*(volatile unsigned *)(fr.ex_cur + (dst - fr.t0)));
    if (ed == INFINITY) return 0;
    const float w = __int_as_float(__ldg(fr.a->g.arcs + lk.z).x);
    float le = ed + (((fr.cost[lk.x] + __int_as_float(lk.w)) + w) - fr.cost[dst]);
    if (le < 0.f) le = 0.f;
    *le_out = le;
    return le <= fr.lb;
}

__global__ void __launch_bounds__(kPruneThreads, 2) lattice_prune_kernel(DecArgs a) {
    const int l = blockIdx.x;
    const LaneDesc ln = a.lanes[l];
    if (!ln.dec_last) return;
    constexpr int NT = kPruneThreads, NW = NT / 32, K = kPruneRegLinks;
    extern __shared__ __align__(16) unsigned s_dyn[];  // [2][kPruneSmemToks] extras | [max_frames + 3] per-frame counts
    auto s_ex = [&](int b) { return s_dyn + (b ? kPruneSmemToks : 0); };
    int *s_cnt = reinterpret_cast<int *>(s_dyn + 2 * kPruneSmemToks);
    __shared__ int s_nlinks, s_nfinal, s_start;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ch = ln.channel;
    const DecChannelState cs = a.cs[ch];
    const float *cost = a.log_cost + (size_t)ch * a.log_cap;
    const int *log_arc = a.log_arc + (size_t)ch * a.log_cap;
    unsigned *extra = a.lat_extra ? a.lat_extra + (size_t)ch * a.log_cap : reinterpret_cast<unsigned *>(a.log_prev + (size_t)ch * a.log_cap);
    int *remap = reinterpret_cast<int *>(extra);  // the new state numbers overwrite the extra costs, frame by frame
    const int *frame_off = a.log_frame_off + (size_t)ch * (a.max_frames + 2);
    const int4 *links = a.links + (size_t)ch * a.link_cap;
    const int *link_off = a.link_off + (size_t)ch * (a.max_frames + 3);
    const int F = min(cs.frame, a.max_frames);
    const int n_tok = cs.log_count;
    const float lb = a.lattice_beam;
    LatHeader *hdr = a.lat_hdr + l;
    int4 *out_links = a.lat_links + (size_t)l * a.lat_link_cap;
    int2 *out_final = a.lat_final + (size_t)l * a.lat_final_cap;
    int *out_frame = a.lat_tok_frame + (size_t)l * a.lat_tok_cap;
    int *out_state = a.lat_tok_state ? a.lat_tok_state + (size_t)l * a.lat_tok_cap : nullptr;
    const int *log_state = a.log_state ? a.log_state + (size_t)ch * a.log_cap : nullptr;
    int error = cs.error;
    if (cs.n_cur == 0 || n_tok == 0 || cs.frame > a.max_frames) {  // search died / log overflow: no lattice
        if (tid == 0) *hdr = LatHeader{0, 0, 0, -1, error ? error : 11, F, 0, 0};
        return;
    }
    if (tid == 0) {
        s_nlinks = 0;
        s_nfinal = 0;
        s_start = -1;
    }
    // last frame: all its tokens were logged in list order, so the token list still gives their states
    const int lo_last = frame_off[F], hi_last = frame_off[F + 1];
    const int *t_state = a.tok_state + (size_t)ch * 2 * a.tok_cap + (size_t)cs.parity * a.tok_cap;
    PruneFrame fr;
    fr.a = &a;
    fr.cost = cost;
    fr.lb = lb;
    int buf = 0;
    fr.ex_cur = hi_last - lo_last <= kPruneSmemToks ? s_ex(0) : extra + lo_last;
    for (int i = lo_last + tid; i < hi_last; i += NT) {
        const float fc = cs.reached_final ? __ldg(a.g.final_cost + t_state[i - lo_last]) : 0.f;
        const float e = (cost[i] + fc) - cs.best_cost;
        fr.ex_cur[i - lo_last] = e <= lb ? __float_as_uint(fmaxf(e, 0.f)) : kInfBits;
    }
    const float *frame_offset = a.frame_offset + (size_t)ch * (a.max_frames + 2);
    // this thread's share of frame F's links
    int4 cur[K];
    {
        const int k0 = link_off[F], k1 = min(link_off[F + 1], a.link_cap);
#pragma unroll
        for (int j = 0; j < K; j++) {
            const int k = k0 + j * NT + tid;
            cur[j] = k < k1 ? links[k] : make_int4(0, -1, 0, 0);
        }
    }
    __syncthreads();
    for (int f = F; f >= 0; f--) {
        const int k0 = link_off[f], k1 = min(link_off[f + 1], a.link_cap);
        fr.t0 = frame_off[f];
        const int n_f = frame_off[f + 1] - fr.t0;
        fr.p0 = f > 0 ? frame_off[f - 1] : 0;
        const int n_p = f > 0 ? fr.t0 - fr.p0 : 0;
        fr.ex_prev = n_p <= kPruneSmemToks ? s_ex(buf ^ 1) : extra + fr.p0;
        fr.off = f > 0 ? frame_offset[f - 1] : 0.f;
        // request the next frame's links (in flight while this frame is worked on)
        int4 nxt[K];
        {
            const int q0 = f > 0 ? link_off[f - 1] : 0, q1 = f > 0 ? k0 : 0;
#pragma unroll
            for (int j = 0; j < K; j++) {
                const int k = q0 + j * NT + tid;
                nxt[j] = k < q1 ? links[k] : make_int4(0, -1, 0, 0);
            }
        }
        for (int i = tid; i < n_p; i += NT) fr.ex_prev[i] = kInfBits;  // (ordered before the emitting pushes by the barriers of the epsilon rounds)
        // epsilon links inside frame f: iterate to the fixed point
        for (;;) {
            int changed = 0;
            auto relax_eps = [&](const int4 lk) {
                if (lk.y < 0 || !(lk.y & kEpsLinkFlag)) return;
                float le;
                if (prune_link(fr, lk, &le)) {
                    const unsigned b = __float_as_uint(le);
                    if (b < atomicMin(fr.ex_cur + (lk.x - fr.t0), b)) changed = 1;
                }
            };
#pragma unroll
            for (int j = 0; j < K; j++) relax_eps(cur[j]);
            for (int k = k0 + K * NT + tid; k < k1; k += NT) relax_eps(links[k]);
            if (!__syncthreads_or(changed)) break;
        }
        // every link into frame f is now decided: survivors go out (old token indices for now; GetRawLattice takes the frame's
        // cost offset back out of the acoustic cost of an emitting arc), emitting survivors push into frame f-1
        auto settle = [&](const int4 lk) {
            if (lk.y < 0) return;
            float le;
            if (!prune_link(fr, lk, &le)) return;
            const bool eps = (lk.y & kEpsLinkFlag) != 0;
            if (!eps) atomicMin(fr.ex_prev + (lk.x - fr.p0), __float_as_uint(le));
            const int o = agg_inc(&s_nlinks);
            const float ac = eps ? 0.f : __int_as_float(lk.w) - fr.off;
            if (o < a.lat_link_cap) out_links[o] = make_int4(lk.x, lk.y & ~kEpsLinkFlag, lk.z, __float_as_int(ac));
        };
#pragma unroll
        for (int j = 0; j < K; j++) settle(cur[j]);
        for (int k = k0 + K * NT + tid; k < k1; k += NT) settle(links[k]);
        __syncthreads();
        // frame f is final: its extras go to the global array the renumbering reads (same thread -> same index as the
        // initialisation of this buffer two frames on, so no barrier is needed in between)
        if (fr.ex_cur == s_ex(buf))
            for (int i = tid; i < n_f; i += NT) extra[fr.t0 + i] = fr.ex_cur[i];
        fr.ex_cur = fr.ex_prev;
        buf ^= 1;
#pragma unroll
        for (int j = 0; j < K; j++) cur[j] = nxt[j];
    }
    __syncthreads();
    // surviving tokens, renumbered in log order (frame by frame): per-frame counts (a warp per frame), their prefix, then the
    // new numbers.  remap[] shares storage with the extra costs: a warp reads a frame's flags before it overwrites them.
    for (int f = warp; f <= F; f += NW) {
        const int lo = frame_off[f], hi = frame_off[f + 1];
        int c = 0;
        for (int i0 = lo; i0 < hi; i0 += 32) {
            const int i = i0 + lane;
            c += __popc(__ballot_sync(0xffffffffu, i < hi && __ldcg(extra + i) != kInfBits));
        }
        if (lane == 0) s_cnt[f] = c;
    }
    __syncthreads();
    if (warp == 0) {  // exclusive prefix of the per-frame counts; s_cnt[F + 1] = total
        int carry = 0;
        for (int f0 = 0; f0 <= F + 1; f0 += 32) {
            const int f = f0 + lane;
            const int v = f <= F ? s_cnt[f] : 0;
            int incl = v;
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (f <= F + 1) s_cnt[f] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
    __syncthreads();
    const int n_keep = s_cnt[F + 1];
    for (int f = warp; f <= F; f += NW) {
        const int lo = frame_off[f], hi = frame_off[f + 1];
        int base = s_cnt[f];
        for (int i0 = lo; i0 < hi; i0 += 32) {
            const int i = i0 + lane;
            const bool keep = i < hi && __ldcg(extra + i) != kInfBits;
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (i < hi) {
                int ni = -1;
                if (keep) {
                    ni = base + __popc(bal & lanemask_lt());
                    if (ni < a.lat_tok_cap) {
                        out_frame[ni] = f;
                        if (out_state) out_state[ni] = log_state ? log_state[i] : -1;
                        if (log_arc[i] < 0 && f == 0) s_start = ni;
                        if (f == F) {  // a token kept only through epsilon links into final tokens is not final itself
                            const float fc = cs.reached_final ? __ldg(a.g.final_cost + t_state[i - lo_last]) : 0.f;
                            if (fc != INFINITY) {
                                const int o = atomicAdd(&s_nfinal, 1);
                                if (o < a.lat_final_cap) out_final[o] = make_int2(ni, __float_as_int(fc));
                            }
                        }
                    }
                }
                remap[i] = ni;
            }
            base += __popc(bal);
        }
    }
    __syncthreads();
    if (n_keep > a.lat_tok_cap) error = error ? error : 8;
    if (s_nlinks > a.lat_link_cap) error = error ? error : 9;
    if (s_nfinal > a.lat_final_cap) error = error ? error : 10;
    // new state numbers into the links; the list is turned around so that the frames ascend again
    const int n_out = min(s_nlinks, a.lat_link_cap);
    for (int k = tid; k < (n_out + 1) / 2; k += NT) {
        const int k2 = n_out - 1 - k;
        int4 x = out_links[k], y = out_links[k2];
        x.x = remap[x.x];
        x.y = remap[x.y];
        y.x = remap[y.x];
        y.y = remap[y.y];
        out_links[k] = y;
        if (k2 != k) out_links[k2] = x;
    }
    if (tid == 0) {
        hdr->n_tok = min(n_keep, a.lat_tok_cap);
        hdr->n_links = n_out;
        hdr->n_final = min(s_nfinal, a.lat_final_cap);
        hdr->start = s_start;
        hdr->error = error;
        hdr->frames = F;
        if (a.counters) atomicAdd(a.counters + 10, (unsigned long long)n_out);
    }
}

extern "C" cudaError_t vbk_lattice_prune(const DecArgs *a, cudaStream_t s) {
    if (!a->lattice || a->num_lanes <= 0) return cudaSuccess;
    const int smem = (2 * kPruneSmemToks + a->max_frames + 3) * 4;
    static int done[16] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 16) return cudaErrorInvalidDevice;
    if (done[dev] < smem) {
        cudaError_t e = cudaFuncSetAttribute(lattice_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        done[dev] = smem;
    }
    lattice_prune_kernel<<<a->num_lanes, kPruneThreads, smem, s>>>(*a);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Partial result: argmin over the current tokens (no final costs, as the reference's partial results do
// [REF src/recognizer.cc:790-793]), then a walk over the back pointers collecting the output labels.  One warp per lane:
// the walk is a dependent chain, so it runs off the search kernel's critical path in its own small launch.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) partial_kernel(DecArgs a) {
    const int l = a.lane_begin + blockIdx.x, lane = threadIdx.x;
    const LaneDesc ln = a.lanes[l];
    if (ln.dec_last) return;
    const int ch = ln.channel;
    const DecChannelState cs = a.cs[ch];
    int *out = a.partial_words ? a.partial_words + (size_t)l * kPartialCap : nullptr;
    const size_t tbase = (size_t)ch * 2 * a.tok_cap + (size_t)cs.parity * a.tok_cap;
    const float *t_cost = a.tok_cost + tbase;
    const int *t_arc = a.tok_arc + tbase, *t_prev = a.tok_prev + tbase;
    const int *t_state = a.tok_state + tbase;
    unsigned long long m = kValMax;
    unsigned mf = 0xffffffffu;  // best cost + final cost (ordered)
    for (int i = lane; i < cs.n_cur; i += 32) {
        const float c = t_cost[i];
        m = min(m, ((unsigned long long)ford(c) << 32) | (unsigned)i);
        if (a.endp_silence) {
            const float fc = __ldg(a.g.final_cost + t_state[i]);
            if (fc != INFINITY) mf = min(mf, ford(c + fc));
        }
    }
    for (int o = 16; o; o >>= 1) {
        m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
        mf = min(mf, __shfl_xor_sync(0xffffffffu, mf, o));
    }
    if (lane != 0) return;
    int n = 0, n_sil = 0;
    bool counting = true;  // trailing silence: emitting arcs of silence phones, back from the newest frame to the first other phone
    if (m != kValMax) {
        const int *log_prev = a.log_prev + (size_t)ch * a.log_cap, *log_arc = a.log_arc + (size_t)ch * a.log_cap;
        int i = (int)(unsigned)m, li = -1;
        for (int guard = 0; guard < a.tok_cap; guard++) {  // epsilon predecessors live in the same token list
            const int arc = t_arc[i];
            if (arc < 0) break;
            const int4 av = __ldg(a.g.arcs + arc);
            const int ol = av.w & ~kArcFlagMask;
            if (ol != 0 && out) {
                if (n < kPartialCap) out[n] = ol;
                n++;
            }
            if (counting && av.z >= 0) {
                if (av.w & kArcSilence) n_sil++;
                else counting = false;
            }
            const int pv = t_prev[i];
            if (pv <= -2) {
                i = -2 - pv;
            } else {
                li = pv;
                break;
            }
        }
        while (li >= 0 && (out || counting)) {
            const int arc = log_arc[li];
            if (arc < 0) break;
            const int4 av = __ldg(a.g.arcs + arc);
            const int ol = av.w & ~kArcFlagMask;
            if (ol != 0 && out) {
                if (n < kPartialCap) out[n] = ol;
                n++;
            }
            if (counting && av.z >= 0) {
                if (av.w & kArcSilence) n_sil++;
                else counting = false;
            }
            li = log_prev[li];
        }
    }
    if (out) a.partial_count[l] = n;
    if (a.endp_silence) {
        a.endp_silence[l] = n_sil;
        // FinalRelativeCost of LatticeFasterDecoder: best (cost + final) - best cost; infinity without a final token
        a.endp_relcost[l] = (m == kValMax || mf == 0xffffffffu) ? INFINITY : unord(mf) - unord((unsigned)(m >> 32));
    }
}

extern "C" cudaError_t vbk_partial(const DecArgs *a, cudaStream_t s) {
    if ((!a->partial_words && !a->endp_silence) || a->lane_end <= a->lane_begin) return cudaSuccess;
    partial_kernel<<<a->lane_end - a->lane_begin, 32, 0, s>>>(*a);
    return cudaGetLastError();
}

extern "C" int vbk_decode_blocks_per_sm(int threads) { return threads >= 1024 ? 1 : threads >= 512 ? 2 : kDecBlocksPerSM; }

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
    // level-1 table: the largest power of two that keeps the variant's residency (1 / 2 / 3 CTAs per SM) within the SM's shared memory
    static int slots_for[3][16] = {};
    static int slots_dim[3][16] = {};
    if (!slots_for[v][dev] || slots_dim[v][dev] != a->out_node.dim) {
        cudaFuncAttributes fa{};
        cudaError_t e = v == 2 ? cudaFuncGetAttributes(&fa, decode_kernel<1024>) : v == 1 ? cudaFuncGetAttributes(&fa, decode_kernel<512>) : cudaFuncGetAttributes(&fa, decode_kernel<256>);
        if (e != cudaSuccess) return e;
        int sm_bytes = 0, cta_max = 0;
        cudaDeviceGetAttribute(&sm_bytes, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
        cudaDeviceGetAttribute(&cta_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const int per_sm = v == 2 ? 1 : v == 1 ? 2 : kDecBlocksPerSM;
        const int ll_bytes = ((a->out_node.dim + 3) & ~3) * 4;
        long long avail = (long long)sm_bytes / per_sm - 1024 - (long long)fa.sharedSizeBytes - ll_bytes;
        avail = std::min<long long>(avail, (long long)cta_max - (long long)fa.sharedSizeBytes - ll_bytes);
        int slots = 1024;
        while (slots * 2 <= kMaxSlots && (long long)slots * 2 * 12 <= avail) slots *= 2;
        if ((long long)slots * 12 > avail) return cudaErrorInvalidConfiguration;
        slots_for[v][dev] = slots;
        slots_dim[v][dev] = a->out_node.dim;
    }
    DecArgs args = *a;
    args.l1_slots = slots_for[v][dev];
    const int smem = ((a->out_node.dim + 3) & ~3) * 4 + args.l1_slots * 12;
    if (done[v][dev] < smem) {
        cudaError_t e = v == 2   ? cudaFuncSetAttribute(decode_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                        : v == 1 ? cudaFuncSetAttribute(decode_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                 : cudaFuncSetAttribute(decode_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        done[v][dev] = smem;
    }
    if (v == 2) decode_kernel<1024><<<grid, 1024, smem, s>>>(args);
    else if (v == 1) decode_kernel<512><<<grid, 512, smem, s>>>(args);
    else decode_kernel<256><<<grid, 256, smem, s>>>(args);
    return cudaGetLastError();
}

}  // namespace vb
