// vb_engine.cu — per-GPU batch engine (host side).  See vb_engine.h for the reference objects it replaces.
#include "vb_engine.h"

#include "vb_result.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

namespace vb {

namespace {
int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}
}  // namespace

template <typename T>
static T *dev_alloc(std::vector<void *> &allocs, size_t n, int fill = -2) {
    void *p = nullptr;
    VB_CUDA_CHECK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    if (fill != -2) VB_CUDA_CHECK(cudaMemset(p, fill, std::max<size_t>(n, 1) * sizeof(T)));
    allocs.push_back(p);
    return reinterpret_cast<T *>(p);
}
template <typename T>
static T *dev_upload(std::vector<void *> &allocs, const T *src, size_t n) {
    T *p = dev_alloc<T>(allocs, n);
    VB_CUDA_CHECK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}
template <typename T>
static T *dev_upload(std::vector<void *> &allocs, const std::vector<T> &v) {
    return dev_upload(allocs, v.data(), v.size());
}

Engine::Engine(const Model &model, const Config &cfg) : model_(model), cfg_(cfg) {
    if (cfg_.frames_per_chunk < 3) throw std::runtime_error("frames-per-chunk must be >= 3");
    if (cfg_.max_lanes > 1024) cfg_.max_lanes = 1024;
    if (cfg_.max_lanes > cfg_.num_channels) cfg_.max_lanes = cfg_.num_channels;
    if (cfg_.pipeline_slots < 1) cfg_.pipeline_slots = 1;
    if (cfg_.debug_capture) cfg_.pipeline_slots = 1;  // the test taps read per-channel state right after each step
    if (cfg_.pipeline_slots > 8) cfg_.pipeline_slots = 8;
    if (cfg_.hash_size & (cfg_.hash_size - 1)) throw std::runtime_error("hash-size must be a power of two");
    if (model.graph.has_negative_eps)
        throw std::runtime_error("HCLG has negative-weight epsilon arcs: not supported by the token log (DESIGN.md)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= cfg_.device)
        throw std::runtime_error(std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range"));
    VB_CUDA_CHECK(cudaSetDevice(cfg_.device));
    VB_CUDA_CHECK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    {
        // stream priorities decide who gets an SM when a search CTA retires: the short front-end CTAs of the next step or the
        // queued search CTAs of this one (fe-priority: 1 = front end first, -1 = search first, 0 = none)
        int lo = 0, hi = 0;
        VB_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const int pfe = cfg_.fe_priority > 0 ? hi : cfg_.fe_priority < 0 ? lo : 0, pdec = cfg_.fe_priority > 0 ? lo : cfg_.fe_priority < 0 ? hi : 0;
        VB_CUDA_CHECK(cudaStreamCreateWithPriority(&fe_stream_, cudaStreamNonBlocking, pfe));
        for (cudaStream_t &q : fe_stream2_) VB_CUDA_CHECK(cudaStreamCreateWithPriority(&q, cudaStreamNonBlocking, pfe));
        VB_CUDA_CHECK(cudaStreamCreateWithPriority(&dec_stream_, cudaStreamNonBlocking, pdec));
        VB_CUDA_CHECK(cudaStreamCreateWithPriority(&dec_stream2_, cudaStreamNonBlocking, pdec));
        VB_CUDA_CHECK(cudaStreamCreateWithPriority(&dec_stream3_, cudaStreamNonBlocking, pdec));
    }
    VB_CUDA_CHECK(cudaStreamCreateWithFlags(&post_stream_, cudaStreamNonBlocking));
    upload_model();
    alloc_state();
    thread_ = std::thread([this] { worker(); });
    if (cfg_.lattice == 1) {
        // the host chain (determinization, alignment, MBR) of finished segments runs on all host threads this engine can claim:
        // hardware threads / (engines of this model x ranks sharing the host under torchrun)
        int share = std::max(1, cfg_.num_engines);
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) share *= std::max(1, atoi(e));
        int n = cfg_.post_threads > 0 ? cfg_.post_threads : std::max(1, (int)std::thread::hardware_concurrency() / share);
        for (int i = 0; i < n; i++) post_threads_.emplace_back([this] { post_worker(); });
    }
}

void Engine::post_worker() {
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> lk(post_mu_);
            post_cv_.wait(lk, [&] { return post_stop_ || !post_queue_.empty(); });
            if (post_queue_.empty()) return;
            job = std::move(post_queue_.front());
            post_queue_.pop_front();
        }
        const auto t0 = std::chrono::steady_clock::now();
        try {
            job();
        } catch (const std::exception &ex) {
            log_msg(-1, "lattice post-processing failed: %s", ex.what());
        }
        {
            std::lock_guard<std::mutex> lk(stats_mu_);
            stats_.post_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            stats_.post_jobs++;
        }
        {
            std::lock_guard<std::mutex> lk(mu_);
            post_outstanding_--;
        }
        cv_done_.notify_all();
    }
}

Engine::~Engine() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_work_.notify_all();
    if (thread_.joinable()) thread_.join();
    {
        std::lock_guard<std::mutex> lk(post_mu_);
        post_stop_ = true;
    }
    post_cv_.notify_all();
    for (auto &t : post_threads_) t.join();
    cudaSetDevice(cfg_.device);
    cudaStreamSynchronize(stream_);
    for (cudaStream_t q : {fe_stream_, fe_stream2_[0], fe_stream2_[1], fe_stream2_[2], dec_stream_, dec_stream2_, dec_stream3_, post_stream_})
        if (q) cudaStreamSynchronize(q);
    for (void *p : allocs_) cudaFree(p);
    for (Slot &sl : slots_) {
        if (sl.stream) cudaStreamSynchronize(sl.stream);
        if (sl.h_staging) cudaFreeHost(sl.h_staging);
        if (sl.h_raw) cudaFreeHost(sl.h_raw);
        if (sl.h_segs) cudaFreeHost(sl.h_segs);
        if (sl.d_raw) cudaFree(sl.d_raw);
        if (sl.d_segs) cudaFree(sl.d_segs);
        if (sl.h_lanes) cudaFreeHost(sl.h_lanes);
        if (sl.h_cs) cudaFreeHost(sl.h_cs);
        if (sl.h_path) cudaFreeHost(sl.h_path);
        if (sl.h_load) cudaFreeHost(sl.h_load);
        if (sl.h_partial) cudaFreeHost(sl.h_partial);
        if (sl.h_endp) cudaFreeHost(sl.h_endp);
        if (sl.h_lat_hdr) cudaFreeHost(sl.h_lat_hdr);
        if (sl.h_lat_pool) cudaFreeHost(sl.h_lat_pool);
        for (auto &ev : sl.ev) if (ev) cudaEventDestroy(ev);
        if (sl.done) cudaEventDestroy(sl.done);
        if (sl.fork) cudaEventDestroy(sl.fork);
        for (auto &j : sl.join) if (j) cudaEventDestroy(j);
        for (auto &j : sl.tier_done) if (j) cudaEventDestroy(j);
        if (sl.fe_done) cudaEventDestroy(sl.fe_done);
        if (sl.dec_done) cudaEventDestroy(sl.dec_done);
        if (sl.stream) cudaStreamDestroy(sl.stream);
    }
    if (h_capture_) cudaFreeHost(h_capture_);
    for (cudaStream_t q : {fe_stream_, fe_stream2_[0], fe_stream2_[1], fe_stream2_[2], dec_stream_, dec_stream2_, dec_stream3_, post_stream_})
        if (q) cudaStreamDestroy(q);
    cudaStreamDestroy(stream_);
}

void Engine::upload_model() {
    const Model &m = model_;
    // ---- MFCC tables (double on the host, fp32 on the device) ----
    {
        std::vector<float> window(kFrameLen), tw(512), dct_t(1600), lifter(40), mel_w(kNumMel * kMelMaxLen, 0.f);
        std::vector<int> mel_start(kNumMel), mel_len(kNumMel);
        for (int i = 0; i < kFrameLen; i++) window[i] = (float)std::pow(0.5 - 0.5 * std::cos(2.0 * M_PI * i / (kFrameLen - 1)), 0.85);
        for (int k = 0; k < 256; k++) {
            tw[2 * k] = (float)std::cos(-2.0 * M_PI * k / kFftSize);
            tw[2 * k + 1] = (float)std::sin(-2.0 * M_PI * k / kFftSize);
        }
        auto mel = [](double f) { return 1127.0 * std::log(1.0 + f / 700.0); };
        const double f_lo = cfg_.mfcc_low_freq, f_hi = cfg_.mfcc_high_freq > 0 ? cfg_.mfcc_high_freq : 8000.0 + cfg_.mfcc_high_freq;
        const double ml = mel(f_lo), mh = mel(f_hi), delta = (mh - ml) / (kNumMel + 1), bw = 16000.0 / kFftSize;
        for (int j = 0; j < kNumMel; j++) {
            double l = ml + j * delta, c = ml + (j + 1) * delta, r = ml + (j + 2) * delta;
            int first = -1, last = -1;
            std::vector<float> w(256, 0.f);
            for (int i = 0; i < 256; i++) {
                double mm = mel(bw * i);
                if (mm > l && mm < r) {
                    w[i] = (float)(mm <= c ? (mm - l) / (c - l) : (r - mm) / (r - c));
                    if (first < 0) first = i;
                    last = i;
                }
            }
            mel_start[j] = first < 0 ? 0 : first;
            mel_len[j] = first < 0 ? 0 : last - first + 1;
            if (mel_len[j] > kMelMaxLen) throw std::runtime_error("mel filter too wide");
            for (int i = 0; i < mel_len[j]; i++) mel_w[j * kMelMaxLen + i] = w[first + i];
        }
        for (int k = 0; k < 40; k++)
            for (int j = 0; j < 40; j++)
                dct_t[j * 40 + k] = (float)(k == 0 ? std::sqrt(1.0 / 40) : std::sqrt(2.0 / 40) * std::cos(M_PI / 40 * (j + 0.5) * k));
        for (int i = 0; i < 40; i++) lifter[i] = (float)(1.0 + 0.5 * 22.0 * std::sin(M_PI * i / 22.0));
        feat_tab_.window = dev_upload(allocs_, window);
        feat_tab_.twiddle = dev_upload(allocs_, tw);
        feat_tab_.mel_start = dev_upload(allocs_, mel_start);
        feat_tab_.mel_len = dev_upload(allocs_, mel_len);
        feat_tab_.mel_w = dev_upload(allocs_, mel_w);
        feat_tab_.dct_t = dev_upload(allocs_, dct_t);
        feat_tab_.lifter = dev_upload(allocs_, lifter);
    }
    // ---- i-vector extractor: derived quantities in double (Kaldi IvectorExtractor::ComputeDerivedVars) ----
    {
        const int F = m.feat_dim, D = m.ivec_dim, G = m.num_gauss, S = 7 * F;
        const Tensor &lda = m.iv_lda.at("lda");
        if (lda.shape[0] != F || lda.shape[1] != S + 1) throw std::runtime_error("final.mat: unexpected shape");
        std::vector<float> lda_t((size_t)(S + 1) * F);
        for (int a = 0; a < F; a++)
            for (int k = 0; k <= S; k++) lda_t[(size_t)k * F + a] = lda.f32()[(size_t)a * (S + 1) + k];
        const float *mi = m.iv_dubm.at("means_invvars").f32(), *iv = m.iv_dubm.at("inv_vars").f32();
        std::vector<float2> ubm_t((size_t)F * G);
        for (int g = 0; g < G; g++)
            for (int a = 0; a < F; a++) ubm_t[(size_t)a * G + g] = make_float2(mi[(size_t)g * F + a], -0.5f * iv[(size_t)g * F + a]);
        const float *M = m.iv_ie.at("M").f32(), *Si = m.iv_ie.at("sigma_inv").f32();
        const size_t NT = (size_t)D * (D + 1) / 2;
        std::vector<float> sim((size_t)G * F * D), U(G * NT);  // U: packed lower triangle (Kaldi keeps it as an SpMatrix)
        std::vector<double> tmp((size_t)F * D);
        for (int g = 0; g < G; g++) {
            const float *Mg = M + (size_t)g * F * D, *Sg = Si + (size_t)g * F * F;
            for (int a = 0; a < F; a++)
                for (int d = 0; d < D; d++) {
                    double s = 0;
                    for (int b = 0; b < F; b++) s += (double)Sg[a * F + b] * Mg[b * D + d];
                    tmp[(size_t)a * D + d] = s;
                    sim[((size_t)g * F + a) * D + d] = (float)s;
                }
            for (int d = 0; d < D; d++)
                for (int e2 = 0; e2 <= d; e2++) {
                    double s = 0;
                    for (int a = 0; a < F; a++) s += (double)Mg[a * D + d] * tmp[(size_t)a * D + e2];
                    U[g * NT + (size_t)d * (d + 1) / 2 + e2] = (float)s;
                }
        }
        const double *cm = m.iv_cmvn.at("stats").f64();
        std::vector<double> gsum(cm, cm + F);
        iv_model_ = IvecModel{F, D, G, S,
                              dev_upload(allocs_, lda_t), dev_upload(allocs_, m.iv_dubm.at("gconsts").f32(), (size_t)G),
                              dev_upload(allocs_, ubm_t), dev_upload(allocs_, sim), dev_upload(allocs_, U),
                              dev_upload(allocs_, gsum), cm[F], m.prior_offset,
                              cfg_.num_gselect, cfg_.min_post, cfg_.posterior_scale, cfg_.max_count, cfg_.cmn_window, cfg_.global_frames};
    }
    // ---- acoustic model nodes: row grid (every frame / every 3rd), first needed time, look-ahead ----
    const int ctx = m.context, C = cfg_.num_channels;
    const int nn = (int)m.node_dim.size();
    max_in_rows_ = cfg_.frames_per_chunk + 3 + 2 * ctx;
    std::vector<int> step(nn, kSubsample), t_first(nn, 1 << 30), cum_right(nn, 0);
    t_first[nn - 1] = 0;
    for (int o = (int)m.ops.size() - 1; o >= 0; o--) {
        const AmOp &op = m.ops[o];
        const int out = o + 1;
        bool mult3 = true;
        int mn = 0;
        for (int off : op.offs) {
            if (off % kSubsample) mult3 = false;
            mn = std::min(mn, off);
        }
        if (step[out] == 1 || !mult3) step[op.in_node] = 1;
        t_first[op.in_node] = std::min(t_first[op.in_node], t_first[out] + mn);
        if (op.byp_node >= 0) {
            if (step[out] == 1) step[op.byp_node] = 1;
            t_first[op.byp_node] = std::min(t_first[op.byp_node], t_first[out]);
        }
    }
    if (t_first[0] != -ctx || step[0] != 1) throw std::runtime_error("internal: context computation mismatch");
    for (size_t o = 0; o < m.ops.size(); o++) {
        int mx = 0;
        for (int off : m.ops[o].offs) mx = std::max(mx, off);
        cum_right[o + 1] = cum_right[m.ops[o].in_node] + mx;
    }
    nodes_.resize(nn);
    for (int n = 0; n < nn; n++) {
        int rows = step[n] == 1 ? max_in_rows_ + 8 : max_in_rows_ / kSubsample + 8;
        if (n == 0) rows += cfg_.cmn_window;
        nodes_[n] = NodeDesc{m.node_dim[n], step[n], next_pow2(rows), t_first[n], cum_right[n], nullptr};
        nodes_[n].buf = dev_alloc<float>(allocs_, (size_t)C * nodes_[n].ring * nodes_[n].dim, 0);
    }
    d_nodes_ = dev_upload(allocs_, nodes_);
    ops_.resize(m.ops.size());
    maps_.resize(m.ops.size());
    for (size_t o = 0; o < m.ops.size(); o++) {
        const AmOp &op = m.ops[o];
        OpDesc d{};
        d.in_node = op.in_node;
        d.out_node = (int)o + 1;
        d.byp_node = op.byp_node;
        d.n_off = (int)op.offs.size();
        for (int i = 0; i < d.n_off; i++) d.offs[i] = op.offs[i];
        d.uses_ivec = op.uses_ivec;
        d.K = op.K;
        d.N = op.N;
        d.relu = op.relu_bn;
        d.has_bn = op.relu_bn;
        float *w = dev_upload(allocs_, op.W->f32(), (size_t)op.N * op.K);
        float *hi = dev_alloc<float>(allocs_, (size_t)op.N * op.K), *lo = dev_alloc<float>(allocs_, (size_t)op.N * op.K);
        if (cfg_.use_tensor_cores) VB_CUDA_CHECK(vbk_split_weights(w, op.N, op.K, cfg_.use_tensor_cores, hi, lo, stream_));
        d.W = w;
        d.W_hi = hi;
        d.W_lo = lo;
        d.bias = op.b ? dev_upload(allocs_, op.b->f32(), (size_t)op.N) : nullptr;
        d.bn_scale = op.bn_s ? dev_upload(allocs_, op.bn_s->f32(), (size_t)op.N) : nullptr;
        d.bn_offset = op.bn_o ? dev_upload(allocs_, op.bn_o->f32(), (size_t)op.N) : nullptr;
        d.bypass_scale = m.bypass_scale;
        ops_[o] = d;
        if (cfg_.use_tensor_cores) {
            VB_CUDA_CHECK(vbk_make_weight_map(hi, op.N, op.K, cfg_.use_tensor_cores, maps_[o].hi));
            VB_CUDA_CHECK(vbk_make_weight_map(lo, op.N, op.K, cfg_.use_tensor_cores, maps_[o].lo));
        }
    }
    // ---- decoding graph: one 16-byte record per arc ----
    std::vector<char> silence;  // endpointing silence phones ("1:2:3") -> flag per phone id
    {
        int v = 0;
        bool have = false;
        for (const char *c = cfg_.endpoint_silence_phones;; c++) {
            if (*c >= '0' && *c <= '9') {
                v = v * 10 + (*c - '0');
                have = true;
            } else {
                if (have) {
                    if ((size_t)v >= silence.size()) silence.resize((size_t)v + 1, 0);
                    silence[(size_t)v] = 1;
                    endpointing_ = true;
                }
                v = 0;
                have = false;
                if (!*c) break;
            }
        }
    }
    {
        const Graph &g = m.graph;
        std::vector<int4> arcs((size_t)g.num_arcs);
        for (int a = 0; a < g.num_arcs; a++) {
            int wbits;
            memcpy(&wbits, &g.arc_w[a], 4);
            const int nx = g.arc_next[a];
            const bool next_has_eps = g.eps_begin[nx] < g.e_begin[nx + 1];
            if (g.arc_olabel[a] >= kArcSilence || nx >= kArcSilence) throw std::runtime_error("graph too large for the packed arc record");
            const int il = g.arc_ilabel[a];
            const bool sil = il > 0 && il < (int)m.tid2phone.size() && m.tid2phone[il] >= 0 && m.tid2phone[il] < (int)silence.size() && silence[m.tid2phone[il]];
            arcs[a] = make_int4(wbits, nx, g.arc_pdf[a], g.arc_olabel[a] | (next_has_eps ? kNextHasEps : 0) | (sil ? kArcSilence : 0));
        }
        std::vector<int2> sa((size_t)g.num_states + 1);
        for (int s2 = 0; s2 < g.num_states; s2++) sa[s2] = make_int2(g.e_begin[s2], g.eps_begin[s2]);
        sa[g.num_states] = make_int2(g.num_arcs, g.num_arcs);
        graph_ = GraphDev{g.num_states, g.num_arcs, g.start, dev_upload(allocs_, g.final_cost), dev_upload(allocs_, g.e_begin),
                          dev_upload(allocs_, g.eps_begin), dev_upload(allocs_, arcs), dev_upload(allocs_, sa)};
    }
    VB_CUDA_CHECK(cudaStreamSynchronize(stream_));
}

void Engine::alloc_state() {
    const int C = cfg_.num_channels, F = model_.feat_dim, D = model_.ivec_dim;
    const int spc = samples_per_chunk();
    const int nn = (int)nodes_.size();
    slot_lanes_ = std::max(1, cfg_.max_lanes);
    const int L = slot_lanes_;
    iv_state_.cmvn_sum = dev_alloc<double>(allocs_, (size_t)C * F, 0);
    iv_state_.norm_ring = dev_alloc<float>(allocs_, (size_t)C * kNormRing * F, 0);
    iv_state_.lin = dev_alloc<double>(allocs_, (size_t)C * D, 0);
    iv_state_.quad = dev_alloc<double>(allocs_, (size_t)C * D * (D + 1) / 2, 0);
    iv_state_.num_frames = dev_alloc<double>(allocs_, (size_t)C, 0);
    iv_state_.ivec = dev_alloc<float>(allocs_, (size_t)C * D, 0);
    iv_frames_cap_ = vbk_ivector_frames_cap(spc);
    d_iv_sel_g_ = dev_alloc<int>(allocs_, (size_t)L * iv_frames_cap_ * 8, 0);
    d_iv_sel_w_ = dev_alloc<float>(allocs_, (size_t)L * iv_frames_cap_ * 8, 0);
    d_iv_fu_ = dev_alloc<float>(allocs_, (size_t)L * iv_frames_cap_ * F, 0);
    d_carry_ = dev_alloc<int16_t>(allocs_, (size_t)C * kCarryMax, 0);
    d_node_end_ = dev_alloc<int>(allocs_, (size_t)C * kMaxNodes, 0);
    // decoder: per-channel state shared by all slots
    max_frames_ = cfg_.max_seconds * 100 / kSubsample + 2;
    log_cap_ = max_frames_ * std::min(cfg_.max_active + 1, cfg_.log_tokens_per_frame) + cfg_.tok_cap;
    path_cap_ = 4 * max_frames_ + 64;
    DecArgs &d = dec_;
    d.g = graph_;
    d.beam = cfg_.beam;
    d.beam_delta = cfg_.beam_delta;
    d.acoustic_scale = cfg_.acoustic_scale;
    d.max_active = cfg_.max_active;
    d.min_active = cfg_.min_active;
    d.tok_cap = cfg_.tok_cap;
    d.cand_cap = cfg_.cand_cap;
    d.hash_size = cfg_.hash_size;
    d.log_cap = log_cap_;
    d.max_frames = max_frames_;
    d.path_cap = path_cap_;
    d.out_node = nodes_.back();
    d.cs = dev_alloc<DecChannelState>(allocs_, (size_t)C, 0);
    d.tok_state = dev_alloc<int>(allocs_, (size_t)C * 2 * cfg_.tok_cap);
    d.tok_cost = dev_alloc<float>(allocs_, (size_t)C * 2 * cfg_.tok_cap);
    d.tok_arc = dev_alloc<int>(allocs_, (size_t)C * 2 * cfg_.tok_cap);
    d.tok_prev = dev_alloc<int>(allocs_, (size_t)C * 2 * cfg_.tok_cap);
    d.log_prev = dev_alloc<int>(allocs_, (size_t)C * log_cap_);
    d.log_arc = dev_alloc<int>(allocs_, (size_t)C * log_cap_);
    d.log_cost = dev_alloc<float>(allocs_, (size_t)C * log_cap_);
    d.log_state = cfg_.debug_capture ? dev_alloc<int>(allocs_, (size_t)C * log_cap_) : nullptr;
    d.log_frame_off = dev_alloc<int>(allocs_, (size_t)C * (max_frames_ + 2), 0);
    d.path = dev_alloc<int>(allocs_, (size_t)C * path_cap_, 0);
    d.counters = dev_alloc<unsigned long long>(allocs_, 64, 0);
    d.lattice = cfg_.lattice;
    d.lattice_beam = cfg_.lattice_beam;
    if (cfg_.lattice) {
        link_cap_ = max_frames_ * cfg_.log_links_per_frame + cfg_.cand_cap;
        d.link_cap = link_cap_;
        d.links = dev_alloc<int4>(allocs_, (size_t)C * link_cap_);
        d.link_off = dev_alloc<int>(allocs_, (size_t)C * (max_frames_ + 3), 0);
        d.frame_offset = dev_alloc<float>(allocs_, (size_t)C * (max_frames_ + 2), 0);
        d.lat_extra = cfg_.debug_capture ? dev_alloc<unsigned>(allocs_, (size_t)C * log_cap_) : nullptr;
        d.lat_link_cap = cfg_.lat_link_cap;
        d.lat_tok_cap = cfg_.lat_tok_cap;
        d.lat_final_cap = cfg_.tok_cap;
    }
    {   // fixed scratch partition per tier (tiers of different steps overlap in time): as many CTAs as the tier can have resident
        int sms = 0;
        VB_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg_.device));
        auto cap = [&](int threads) { return std::min(sms * vbk_decode_blocks_per_sm(threads), L); };  // never more CTAs than lanes
        tier_scratch_[0] = 0;
        tier_scratch_[1] = cap(cfg_.heavy_threads);
        tier_scratch_[2] = tier_scratch_[1] + cap(cfg_.mid_threads);
        d.grid = tier_scratch_[2] + cap(cfg_.light_threads);
    }
    {   // search scratch, one copy: the searches of all steps run in order on dec_stream_
        const size_t G = (size_t)d.grid;
        d.hash_key = dev_alloc<int>(allocs_, G * cfg_.hash_size, 0xff);
        d.hash_val = dev_alloc<unsigned long long>(allocs_, G * cfg_.hash_size, 0xff);
        d.hash_tok = dev_alloc<int>(allocs_, G * cfg_.hash_size, 0);
        d.cand = dev_alloc<int4>(allocs_, G * cfg_.cand_cap);
        d.cand_next = dev_alloc<int>(allocs_, G * cfg_.cand_cap);
        d.eps_work = dev_alloc<int>(allocs_, G * cfg_.cand_cap);
        d.rank = dev_alloc<int>(allocs_, G * cfg_.tok_cap);
        d.sv_pref = dev_alloc<int>(allocs_, G * cfg_.tok_cap);
        d.sv_a0 = dev_alloc<int>(allocs_, G * cfg_.tok_cap);
        d.sv_src = dev_alloc<int>(allocs_, G * cfg_.tok_cap);
        d.sv_cost = dev_alloc<float>(allocs_, G * cfg_.tok_cap);
        d.win_owner = dev_alloc<int>(allocs_, G * (cfg_.cand_cap / 32 + 2), 0);
    }
    // a slot is reused only after its step completed, so the front end runs at most (slots - 1) steps ahead of the
    // search: the log-likelihood ring must hold that many chunks beside the one being searched
    {
        const int rows_per_chunk = cfg_.frames_per_chunk / kSubsample + 2;
        const int depth = std::max(1, nodes_.back().ring / rows_per_chunk);
        if (cfg_.pipeline_slots > depth) cfg_.pipeline_slots = depth;
    }
    // pipeline slots: staging, per-step tables, result buffers
    slots_.resize(cfg_.pipeline_slots);
    active_slots_ = cfg_.pipeline_slots;
    for (Slot &sl : slots_) {
        VB_CUDA_CHECK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        for (auto &ev : sl.ev) VB_CUDA_CHECK(cudaEventCreate(&ev));
        VB_CUDA_CHECK(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
        VB_CUDA_CHECK(cudaEventCreateWithFlags(&sl.fe_done, cudaEventDisableTiming));
        VB_CUDA_CHECK(cudaEventCreateWithFlags(&sl.dec_done, cudaEventDisableTiming));
        VB_CUDA_CHECK(cudaEventCreateWithFlags(&sl.fork, cudaEventDisableTiming));
        for (auto &j : sl.join) VB_CUDA_CHECK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
        for (auto &j : sl.tier_done) VB_CUDA_CHECK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
        sl.d_queue = dev_alloc<int>(allocs_, 4, 0);
        sl.d_staging = dev_alloc<int16_t>(allocs_, (size_t)L * spc, 0);
        VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_staging, (size_t)L * spc * sizeof(int16_t)));
        sl.d_lanes = dev_alloc<LaneDesc>(allocs_, (size_t)L, 0);
        VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_lanes, (size_t)L * sizeof(LaneDesc)));
        sl.d_table = dev_alloc<NodeLane>(allocs_, (size_t)nn * L, 0);
        sl.d_rowoff = dev_alloc<int>(allocs_, (size_t)nn * (L + 1), 0);
        rows_cap_ = L * (max_in_rows_ + 8);
        sl.d_rows = dev_alloc<int2>(allocs_, (size_t)nn * rows_cap_, 0);
        for (int k = 0; k + 1 < std::min(cfg_.fe_split, 4); k++) {
            sl.d_rowoff2[k] = dev_alloc<int>(allocs_, (size_t)nn * (L + 1), 0);
            sl.d_rows2[k] = dev_alloc<int2>(allocs_, (size_t)nn * rows_cap_, 0);
        }
        VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_cs, (size_t)L * sizeof(DecChannelState)));
        VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_path, (size_t)L * path_cap_ * sizeof(int)));
        sl.d_load = dev_alloc<int>(allocs_, (size_t)L, 0);
        VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_load, (size_t)L * sizeof(int)));
        DecArgs &sd = sl.dec;
        sd = dec_;
        sd.out_table = sl.d_table + (size_t)(nn - 1) * L;
        sd.lane_load = sl.d_load;
        if (endpointing_) {
            sd.endp_silence = dev_alloc<int>(allocs_, (size_t)L, 0);
            sd.endp_relcost = dev_alloc<float>(allocs_, (size_t)L, 0);
            VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_endp, (size_t)L * 2 * sizeof(int)));
        }
        if (cfg_.partials) {
            sd.partial_words = dev_alloc<int>(allocs_, (size_t)L * kPartialCap, 0);
            sd.partial_count = dev_alloc<int>(allocs_, (size_t)L, 0);
            VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_partial, (size_t)L * (kPartialCap + 1) * sizeof(int)));
        }
        if (cfg_.lattice) {
            sd.lat_hdr = dev_alloc<LatHeader>(allocs_, (size_t)L, 0);
            sd.lat_links = dev_alloc<int4>(allocs_, (size_t)L * cfg_.lat_link_cap);
            sd.lat_final = dev_alloc<int2>(allocs_, (size_t)L * cfg_.tok_cap);
            sd.lat_tok_frame = dev_alloc<int>(allocs_, (size_t)L * cfg_.lat_tok_cap);
            sd.lat_tok_state = cfg_.debug_capture ? dev_alloc<int>(allocs_, (size_t)L * cfg_.lat_tok_cap) : nullptr;
            VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_lat_hdr, (size_t)L * sizeof(LatHeader)));
            VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_lat_pool, std::max(kLatPoolBytes, (size_t)cfg_.lat_link_cap * sizeof(int4) + (size_t)cfg_.tok_cap * sizeof(int2) + (size_t)2 * cfg_.lat_tok_cap * sizeof(int) + 64)));
        }
    }
    if (cfg_.debug_capture) {
        capture_floats_ = (size_t)(max_in_rows_ + 8) * std::max(model_.node_dim.back(), F);
        d_capture_ = dev_alloc<float>(allocs_, capture_floats_);
        VB_CUDA_CHECK(cudaMallocHost((void **)&h_capture_, capture_floats_ * sizeof(float)));
    }
    free_channels_.resize(C);
    for (int i = 0; i < C; i++) free_channels_[i] = C - 1 - i;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    log_msg(0, "engine on device %d: %d channels, %d slots x %d lanes, chunk %d frames, HBM used %.1f GB of %.1f GB", cfg_.device, C,
            cfg_.pipeline_slots, L, cfg_.frames_per_chunk, (total_b - free_b) / 1e9, total_b / 1e9);
}

std::shared_ptr<Stream> Engine::open_stream() {
    auto s = std::make_shared<Stream>();
    std::lock_guard<std::mutex> lk(mu_);
    s->id = next_id_++;
    s->load = cfg_.mid_tokens + 1;  // no history yet: the middle tier (a first chunk with speech in it is no light lane)
    return s;
}

// Phase table of Kaldi's LinearResample for one input rate, with the reference's filter settings
// [REF src/batch_recognizer.cc:27-29]; called on the worker thread, tables are few and built once.
int Engine::resample_table(int rate) {
    auto it = resample_ids_.find(rate);
    if (it != resample_ids_.end()) return it->second;
    constexpr int kMaxTables = 16;
    if ((int)resample_tables_.size() >= kMaxTables) throw std::runtime_error("too many distinct input sample rates");
    LinearResampler rs((float)rate, 16000.0f, std::min((float)rate / 2, 8000.0f), 6);
    const int U = rs.out_unit();
    int max_taps = 0;
    for (const auto &w : rs.weights()) max_taps = std::max(max_taps, (int)w.size());
    std::vector<int> ntaps(U);
    std::vector<float> w((size_t)U * max_taps, 0.f);
    for (int i = 0; i < U; i++) {
        ntaps[i] = (int)rs.weights()[i].size();
        std::copy(rs.weights()[i].begin(), rs.weights()[i].end(), w.begin() + (size_t)i * max_taps);
    }
    ResampleTable t{rs.in_unit(), U, max_taps, dev_upload(allocs_, rs.first_index()), dev_upload(allocs_, ntaps), dev_upload(allocs_, w)};
    if (!d_resample_tables_) d_resample_tables_ = dev_alloc<ResampleTable>(allocs_, kMaxTables);
    resample_tables_.push_back(t);
    VB_CUDA_CHECK(cudaMemcpy(d_resample_tables_, resample_tables_.data(), resample_tables_.size() * sizeof(ResampleTable), cudaMemcpyHostToDevice));
    const int id = (int)resample_tables_.size() - 1;
    resample_ids_[rate] = id;
    return id;
}

void Engine::push(const std::shared_ptr<Stream> &s, const int16_t *samples, int n, bool last) {
    Stream::Chunk ch;
    ch.samples.assign(samples, samples + n);
    ch.last = last;
    push_chunk(s, std::move(ch));
}

void Engine::push_chunk(const std::shared_ptr<Stream> &s, Stream::Chunk &&ch_in) {
    Stream::Chunk ch = std::move(ch_in);
    const bool last = ch.last;
    ch.t_push = std::chrono::steady_clock::now();
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (s->finished) return;  // chunks after the last one are ignored
        if (last) s->finished = true;
        s->pending.push_back(std::move(ch));
        s->pending_chunks.fetch_add(1);
        outstanding_++;
        if (!s->queued && !s->in_flight) {
            s->queued = true;
            ready_.push_back(s);
        }
    }
    cv_work_.notify_one();
}

void Engine::wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [this] { return outstanding_ == 0 && post_outstanding_ == 0; });
}

StepStats Engine::stats() {
    unsigned long long c[64] = {};
    cudaSetDevice(cfg_.device);
    cudaMemcpy(c, dec_.counters, sizeof c, cudaMemcpyDeviceToHost);
    std::lock_guard<std::mutex> lk(stats_mu_);
    StepStats s = stats_;
    s.tok = c[0];
    s.arc_e = c[1];
    s.arc_eps = c[2];
    s.tok_new = c[3];
    s.lane_cycles_sum = c[4];
    s.lane_cycles_max = c[5];
    s.max_tokens = c[6];
    s.lane_launches = c[7];
    s.arcs_staged = c[8];
    s.links = c[9];
    s.lat_arcs = c[10];
    for (int k = 0; k < 16; k++) s.phase[k] = c[16 + k];
    for (int k = 0; k < 24; k++) s.phase_slowest[k] = c[32 + k];
    for (int t = 0; t < 3; t++) {
        s.tier_slowest_cycles[t] = c[11 + t];
        s.tier_slowest_tokens[t] = c[56 + t];
        s.tier_lane_launches[t] = c[59 + t];
    }
    s.lattice_fallbacks = lattice_fallbacks_.load();
    return s;
}
void Engine::latency(double *out, bool reset) {
    std::vector<float> v;
    {
        std::lock_guard<std::mutex> lk(stats_mu_);
        v = latencies_ms_;
        if (reset) latencies_ms_.clear();
    }
    for (int k = 0; k < 5; k++) out[k] = 0;
    if (v.empty()) return;
    std::sort(v.begin(), v.end());
    auto pct = [&](double p) { return (double)v[std::min(v.size() - 1, (size_t)(p * v.size()))]; };
    double sum = 0;
    for (float x : v) sum += x;
    out[0] = pct(0.50);
    out[1] = pct(0.90);
    out[2] = pct(0.99);
    out[3] = sum / v.size();
    out[4] = (double)v.size();
}

void Engine::reset_stats() {
    cudaSetDevice(cfg_.device);
    cudaMemset(dec_.counters, 0, 64 * sizeof(unsigned long long));
    std::lock_guard<std::mutex> lk(stats_mu_);
    stats_ = StepStats{};
    lattice_fallbacks_ = 0;
}

// The batcher: keeps up to pipeline_slots steps in flight.  Steps complete in launch order (both pipes are in order),
// so the slots form a ring: `head` is the oldest step in flight, `tail` the next free slot.  A stream goes back to the
// ready queue as soon as its chunk has been launched — per-stream order is guaranteed by the pipes, not by waiting.
void Engine::worker() {
    cudaSetDevice(cfg_.device);
    int head = 0, tail = 0, n_busy = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_work_.wait(lk, [&] { return stop_ || !ready_.empty() || n_busy > 0; });
            if (stop_ && n_busy == 0) return;
        }
        const int nslots = std::min((int)slots_.size(), active_slots_.load());
        bool have_ready;
        {
            std::lock_guard<std::mutex> lk(mu_);
            have_ready = !ready_.empty();
        }
        // complete the oldest step when it is done, or when nothing else can be started
        if (n_busy > 0) {
            Slot &old = slots_[head];
            const bool must = n_busy >= nslots || !have_ready;
            if (must || cudaEventQuery(old.done) == cudaSuccess) {
                try {
                    if (old.failed) throw std::runtime_error("the step could not be launched");
                    complete_step(old);
                } catch (const std::exception &ex) {
                    log_msg(-1, "engine step failed: %s", ex.what());
                    for (auto &ln : old.lanes)
                        if (ln.seg_end && ln.s->on_result) {  // an (empty) result in the segment's place keeps the stream's results in order
                            BestPath bp;
                            bp.error = 100;
                            bp.seq = ln.seg_index;
                            bp.offset = ln.seg_offset;
                            ln.s->on_result(bp);
                        }
                }
                old.failed = false;
                old.busy = false;
                n_busy--;
                head = (head + 1) % (int)slots_.size();
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    for (auto &ln : old.lanes) {
                        ln.s->pending_chunks.fetch_sub(1);
                        if (ln.chunk.last) {
                            free_channels_.push_back(ln.s->channel);
                            ln.s->channel = -1;
                        }
                        if (ln.holds) {
                            // the stream was held back — until the endpoint decision of this chunk (silence endpointing), or because
                            // the chunk closed a segment in mid-stream and its traceback / lattice still lived in the channel's
                            // search state.  A detected endpoint first closes the segment with an empty chunk, then the stream
                            // goes on (a new segment) with what is queued
                            if (endpointing_ && !ln.s->resident && ln.endpoint && !ln.chunk.last) {
                                Stream::Chunk fin;
                                fin.last = false;
                                fin.close_segment = true;
                                fin.t_push = std::chrono::steady_clock::now();
                                ln.s->pending.push_front(std::move(fin));
                                ln.s->pending_chunks.fetch_add(1);
                                outstanding_++;
                            }
                            ln.s->in_flight = false;
                            if (!ln.s->pending.empty() && !ln.s->queued) {
                                ln.s->queued = true;
                                ready_.push_back(ln.s);
                            }
                        }
                    }
                    outstanding_ -= (long long)old.lanes.size();
                }
                old.lanes.clear();
                cv_done_.notify_all();
                continue;
            }
        }
        if (!have_ready || n_busy >= nslots) {
            std::this_thread::yield();
            continue;
        }
        Slot &sl = slots_[tail];
        bool starved = false;
        {
            std::unique_lock<std::mutex> lk(mu_);
            std::deque<std::shared_ptr<Stream>> deferred, again;
            while (!ready_.empty() && (int)sl.lanes.size() < slot_lanes_) {
                std::shared_ptr<Stream> s = ready_.front();
                ready_.pop_front();
                if (s->channel < 0) {
                    if (free_channels_.empty()) {  // all channels busy: the stream waits for one to finish
                        deferred.push_back(s);
                        continue;
                    }
                    s->channel = free_channels_.back();
                    free_channels_.pop_back();
                }
                Lane ln;
                ln.s = s;
                ln.chunk = std::move(s->pending.front());
                s->pending.pop_front();
                if (endpointing_ && !s->resident) {
                    s->in_flight = true;  // the next chunk is queued when this one has completed (endpoint decision)
                    s->queued = false;
                    ln.holds = true;
                } else if (!s->pending.empty()) again.push_back(s);  // next chunk: a later step (one chunk per stream per step)
                else s->queued = false;
                sl.lanes.push_back(std::move(ln));
            }
            starved = sl.lanes.empty() && !deferred.empty() && n_busy == 0;
            for (auto it = deferred.rbegin(); it != deferred.rend(); ++it) ready_.push_front(*it);
            for (auto &s : again) ready_.push_back(s);
            if (starved) cv_work_.wait_for(lk, std::chrono::milliseconds(1));
        }
        if (sl.lanes.empty()) {
            if (!starved) std::this_thread::yield();
            continue;
        }
        try {
            launch_step(sl, resident_audio_, resident_stride_);
        } catch (const std::exception &ex) {
            log_msg(-1, "engine launch failed: %s", ex.what());
            sl.failed = true;
            cudaEventRecord(sl.done, dec_stream_);  // let the completion path release the lanes
        }
        {
            // A chunk that closes a segment in mid-stream (rule 5, or an injected close) leaves its result — path, token log,
            // link log, the scratch lattice pruning borrows — in the channel until the step has completed; the stream's next
            // chunk starts a new search in the same channel, so it must not be launched before that: hold the stream back.
            std::lock_guard<std::mutex> lk(mu_);
            for (auto &ln : sl.lanes) {
                if (!ln.seg_end || ln.chunk.last || ln.s->in_flight) continue;
                ln.s->in_flight = true;
                ln.holds = true;
                if (ln.s->queued) {
                    auto it = std::find(ready_.begin(), ready_.end(), ln.s);
                    if (it != ready_.end()) ready_.erase(it);
                    ln.s->queued = false;
                }
            }
        }
        sl.busy = true;
        n_busy++;
        tail = (tail + 1) % (int)slots_.size();
    }
}

void Engine::launch_step(Slot &sl, const int16_t *d_resident, int resident_stride) {
    const auto host_t0 = std::chrono::steady_clock::now();
    std::vector<Lane> &lanes = sl.lanes;
    const int L = (int)lanes.size();
    const int ctx = model_.context, spc = samples_per_chunk();
    const int nn = (int)nodes_.size();
    const int SL = slot_lanes_;
    cudaStream_t st = fe_stream_;
    // heaviest lanes first: the search kernels pull lanes from a queue in this order (longest first), and the lanes
    // above heavy_tokens form a prefix that gets the 1024-thread CTAs
    std::stable_sort(lanes.begin(), lanes.end(), [](const Lane &x, const Lane &y) { return x.s->load > y.s->load; });
    sl.audio = 0;
    long long in_rows = 0;
    size_t raw_used = 0;
    int n_segs = 0;
    for (int i = 0; i < L; i++) {
        Stream &s = *lanes[i].s;
        const Stream::Chunk &ck = lanes[i].chunk;
        LaneDesc &d = sl.h_lanes[i];
        const int n = s.resident ? ck.n_resident : ck.rate ? ck.n_out : (int)ck.samples.size();
        d.channel = s.channel;
        d.n_samples = n;
        d.carry = s.carry;
        d.first = !s.started;
        d.last = ck.last;
        d.frames_before = s.frames;
        const int64_t total = s.samples + n;
        d.frames_after = num_frames_for(total);
        d.iv_end_before = s.iv_end;
        d.iv_end_after = std::max(s.iv_end, ck.last ? d.frames_after : d.frames_after - 3);
        d.in_end_before = d.first ? -ctx : s.in_end;
        d.in_end_after = d.frames_after > 0 ? (ck.last ? d.frames_after + ctx : d.frames_after) : d.in_end_before;
        d.dec_frames_before = s.dec_frames;
        d.src_row = s.resident ? s.resident_row : i;
        d.src_off = s.resident ? (int)s.samples : 0;
        if (ck.rate && !s.resident) {
            // resampled on the device: stage the raw samples and the segment list of this lane
            if (!sl.h_raw) {
                const size_t cap = (size_t)SL * max_resample_raw(), scap = (size_t)SL * kMaxResampleSegs;
                VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_raw, cap * sizeof(int16_t)));
                VB_CUDA_CHECK(cudaMallocHost((void **)&sl.h_segs, scap * sizeof(ResampleSeg)));
                VB_CUDA_CHECK(cudaMalloc((void **)&sl.d_raw, cap * sizeof(int16_t)));
                VB_CUDA_CHECK(cudaMalloc((void **)&sl.d_segs, scap * sizeof(ResampleSeg)));
            }
            const int table = resample_table(ck.rate);
            if (!ck.raw.empty()) memcpy(sl.h_raw + raw_used, ck.raw.data(), ck.raw.size() * sizeof(int16_t));
            for (const Stream::Chunk::Seg &g : ck.segs)
                sl.h_segs[n_segs++] = ResampleSeg{i, (int)raw_used + g.raw_off, g.in_base, g.n_in, g.out_first, g.out_pos, g.n_out, table};
            raw_used += ck.raw.size();
        } else if (!s.resident && n) {
            memcpy(sl.h_staging + (size_t)i * spc, ck.samples.data(), (size_t)n * sizeof(int16_t));
        }
        in_rows += d.in_end_after - d.in_end_before + 2;
        sl.audio += n / 16000.0;
        // advance the host mirror of the stream state (a stream has at most one chunk in flight)
        s.started = true;
        s.samples = total;
        s.frames = d.frames_after;
        s.iv_end = d.iv_end_after;
        s.in_end = d.in_end_after;
        s.carry = d.frames_after > 0 ? (int)(total - (int64_t)kFrameShift * d.frames_after) : (int)total;
        s.dec_frames = d.in_end_after > ctx ? (d.in_end_after - ctx + kSubsample - 1) / kSubsample : 0;
        lanes[i].dec_frames_after = s.dec_frames;
        // segmentation: a segment ends with the stream, or at the first chunk boundary where the decoded length reaches rule 5
        const int rule5 = cfg_.endpoint_rule5_seconds > 0 ? (int)std::ceil(cfg_.endpoint_rule5_seconds / 0.03 - 1e-6) : 0;
        const bool seg_end = ck.last || ck.close_segment || (rule5 > 0 && s.dec_frames - s.seg_start >= rule5);
        lanes[i].seg_start = s.seg_start;
        d.dec_first = !s.seg_open;
        d.dec_last = seg_end;
        lanes[i].seg_end = seg_end;
        lanes[i].seg_offset = (float)(s.seg_start * 0.03);
        lanes[i].seg_index = s.seg_index;
        s.seg_open = !seg_end;
        if (seg_end) {
            s.seg_start = s.dec_frames;
            s.seg_index++;
        }
    }
    sl.launches = 0;
    sl.h2d = 0;
    sl.d2h = 0;
    sl.gemms = 0;
    sl.resample_segs = n_segs;
    sl.timed = timing_;
    sl.pruned = false;
    if (sl.timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[0], st));
    VB_CUDA_CHECK(cudaMemcpyAsync(sl.d_lanes, sl.h_lanes, (size_t)L * sizeof(LaneDesc), cudaMemcpyHostToDevice, st));
    sl.h2d += (size_t)L * sizeof(LaneDesc);
    if (!d_resident) {
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.d_staging, sl.h_staging, (size_t)L * spc * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        sl.h2d += (size_t)L * spc * sizeof(int16_t);
    }
    if (n_segs) {
        // K0: the non-16 kHz lanes' chunks are produced on the device (after the staging copy, which they overwrite row-wise)
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.d_raw, sl.h_raw, raw_used * sizeof(int16_t), cudaMemcpyHostToDevice, st));
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.d_segs, sl.h_segs, (size_t)n_segs * sizeof(ResampleSeg), cudaMemcpyHostToDevice, st));
        sl.h2d += raw_used * sizeof(int16_t) + (size_t)n_segs * sizeof(ResampleSeg);
        ResampleArgs ra{sl.d_segs, n_segs, sl.d_raw, d_resample_tables_, sl.d_staging, spc};
        VB_CUDA_CHECK(vbk_resample(&ra, st));
        sl.launches++;
    }
    // The front end of lanes [first, first + count) as one chain of launches on stream q: features, i-vector, row plan, the
    // TDNN-F GEMMs.  Its launches cover a fraction of a wave each (79 row tiles for 512 lanes), so a full-width step runs two such
    // chains side by side, half of the lanes each (fe-split): the lanes of the halves are different channels, nothing is shared
    // but the read-only model, and the per-lane row table of the output node (what the search reads) is one array for both.
    auto front_end = [&](int first, int count, cudaStream_t q, int *d_rowoff, int2 *d_rows, long long rows_in, bool timed) {
        FeatArgs fa{sl.d_lanes + first, count, d_resident ? d_resident : sl.d_staging, d_resident ? (long long)resident_stride : (long long)spc, spc,
                    d_carry_, nodes_[0], ctx, feat_tab_};
        VB_CUDA_CHECK(vbk_mfcc(&fa, q));
        sl.launches++;
        if (timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[1], q));
        IvecArgs ia{sl.d_lanes + first, count, nodes_[0], ctx, iv_model_, iv_state_, d_iv_sel_g_ + (size_t)first * iv_frames_cap_ * 8,
                    d_iv_sel_w_ + (size_t)first * iv_frames_cap_ * 8, d_iv_fu_ + (size_t)first * iv_frames_cap_ * model_.feat_dim, iv_frames_cap_};
        VB_CUDA_CHECK(vbk_ivector(&ia, q));
        sl.launches += 3;
        if (timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[2], q));
        NnetPlanArgs pa{sl.d_lanes + first, count, nn, d_nodes_, d_node_end_, sl.d_table + first, d_rowoff, SL, d_rows, rows_cap_};
        VB_CUDA_CHECK(vbk_nnet_plan(&pa, q));
        sl.launches++;
        for (size_t o = 0; o < ops_.size(); o++) {
            const OpDesc &op = ops_[o];
            GemmArgs ga{};
            ga.op = op;
            ga.in = nodes_[op.in_node];
            ga.out = nodes_[op.out_node];
            ga.byp = nodes_[op.byp_node >= 0 ? op.byp_node : 0];
            ga.lanes = sl.d_lanes + first;
            ga.num_lanes = count;
            ga.table = sl.d_table + (size_t)op.out_node * SL + first;
            ga.rowoff = d_rowoff + (size_t)op.out_node * (SL + 1);
            ga.rows = d_rows + (size_t)op.out_node * rows_cap_;
            ga.ivec = iv_state_.ivec;
            ga.ivec_dim = model_.ivec_dim;
            ga.max_rows = (int)(ga.out.step == 1 ? rows_in : rows_in / kSubsample + 2 * count);
            ga.tc_mode = cfg_.use_tensor_cores;
            ga.map_hi = maps_[o].hi;
            ga.map_lo = maps_[o].lo;
            VB_CUDA_CHECK(cfg_.use_tensor_cores ? vbk_gemm_tc(&ga, q) : vbk_gemm_fp32(&ga, q));
            sl.launches++;
            sl.gemms++;
        }
    };
    // (stage timing keeps the single chain: its events bracket the stages of one stream)
    // (while the host lattice pool has a backlog the host is what bounds the throughput: the extra launches of several chains
    // would only take cycles from it)
    bool host_bound = false;
    if (!post_threads_.empty()) {
        // a pool that is behind at (nearly) every launch is the bottleneck; one that drains between the bursts of finished streams is not
        size_t backlog;
        {
            std::lock_guard<std::mutex> lk(post_mu_);
            backlog = post_queue_.size();
        }
        post_backlog_ema_ = 0.9f * post_backlog_ema_ + 0.1f * (backlog > post_threads_.size() ? 1.f : 0.f);
        host_bound = post_backlog_ema_ > 0.85f;
    }
    const int chains = sl.timed || L < 128 || host_bound ? 1 : std::max(1, std::min(cfg_.fe_split, 4));
    if (chains == 1) {
        front_end(0, L, st, sl.d_rowoff, sl.d_rows, in_rows, sl.timed);
    } else {
        // the other chains start once the lane descriptors / samples are on the device, which is after the previous step's chains
        // (a channel may change shares between steps: the lanes are sorted by load)
        VB_CUDA_CHECK(cudaEventRecord(sl.fork, st));
        int first = L / chains;  // chain 0 (this stream) takes [0, L / chains)
        for (int k = 1; k < chains; k++) {
            const int end = k + 1 == chains ? L : (int)((long long)L * (k + 1) / chains);
            long long rows_k = 0;
            for (int i = first; i < end; i++) rows_k += sl.h_lanes[i].in_end_after - sl.h_lanes[i].in_end_before + 2;
            VB_CUDA_CHECK(cudaStreamWaitEvent(fe_stream2_[k - 1], sl.fork, 0));
            front_end(first, end - first, fe_stream2_[k - 1], sl.d_rowoff2[k - 1], sl.d_rows2[k - 1], rows_k, false);
            VB_CUDA_CHECK(cudaEventRecord(sl.join[k - 1], fe_stream2_[k - 1]));
            first = end;
        }
        long long rows_0 = 0;
        for (int i = 0; i < L / chains; i++) rows_0 += sl.h_lanes[i].in_end_after - sl.h_lanes[i].in_end_before + 2;
        front_end(0, L / chains, st, sl.d_rowoff, sl.d_rows, rows_0, false);
        for (int k = 1; k < chains; k++) VB_CUDA_CHECK(cudaStreamWaitEvent(st, sl.join[k - 1], 0));
    }
    if (sl.timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[3], st));
    VB_CUDA_CHECK(cudaMemsetAsync(sl.d_queue, 0, 4 * sizeof(int), st));
    VB_CUDA_CHECK(cudaEventRecord(sl.fe_done, st));
    if (sl.timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[5], st));
    // ---- search: three in-order tier pipes ----
    // Lanes are split by the token load of their previous step (they are sorted by it) into 1024-, 512- and 256-thread CTA
    // tiers.  Each tier is an in-order pipe of its own across steps (own stream, own scratch range, own lane queue), so the
    // light lanes of step s+1 start as soon as their front end and their own previous search are done, while the few heavy
    // lanes of step s — the critical path of a full-width step — are still running.  A lane that changes tier waits for
    // the pipe it comes from.
    sl.dec.lanes = sl.d_lanes;
    sl.dec.num_lanes = L;
    sl.dec.queue = sl.d_queue;
    int n_heavy = 0, n_mid = 0;
    while (n_heavy < L && lanes[n_heavy].s->load > cfg_.heavy_tokens) n_heavy++;
    n_mid = n_heavy;
    while (n_mid < L && lanes[n_mid].s->load > cfg_.mid_tokens) n_mid++;
    const int tier_begin[3] = {0, n_heavy, n_mid}, tier_end[3] = {n_heavy, n_mid, L};
    const int tier_threads[3] = {cfg_.heavy_threads, cfg_.mid_threads, cfg_.light_threads};
    cudaStream_t tier_stream[3] = {dec_stream_, dec_stream2_, dec_stream3_};
    int n_last = 0;
    for (int i = 0; i < L; i++) n_last += lanes[i].seg_end ? 1 : 0;
    bool tier_used[3] = {false, false, false};
    // all waits are issued before any of this step's tier events is (re-)recorded: with few slots the event a lane has to
    // wait for can be the very event object this step records next
    for (int t = 0; t < 3; t++) {
        if (tier_end[t] <= tier_begin[t]) continue;
        tier_used[t] = true;
        cudaStream_t ts = tier_stream[t];
        VB_CUDA_CHECK(cudaStreamWaitEvent(ts, sl.fe_done, 0));
        bool from[3] = {false, false, false};
        for (int i = tier_begin[t]; i < tier_end[t]; i++) {
            Stream &s = *lanes[i].s;
            if (s.last_tier >= 0 && s.last_tier != t) from[s.last_tier] = true;
            s.last_tier = t;
        }
        for (int u = 0; u < 3; u++)
            if (from[u] && last_tier_done_[u]) VB_CUDA_CHECK(cudaStreamWaitEvent(ts, last_tier_done_[u], 0));
    }
    for (int t = 0; t < 3; t++) {
        if (!tier_used[t]) continue;
        cudaStream_t ts = tier_stream[t];
        DecArgs da = sl.dec;
        da.queue = sl.d_queue + t;
        da.lane_begin = tier_begin[t];
        da.lane_end = tier_end[t];
        da.scratch_base = tier_scratch_[t];
        VB_CUDA_CHECK(vbk_decode(&da, tier_threads[t], ts));
        sl.launches++;
        if (cfg_.partials || endpointing_) {  // the partial walk reads the lane's tokens: it must precede the lane's next search, so it rides the tier pipe
            VB_CUDA_CHECK(vbk_partial(&da, ts));
            sl.launches++;
        }
        VB_CUDA_CHECK(cudaEventRecord(sl.tier_done[t], ts));
        last_tier_done_[t] = sl.tier_done[t];
    }
    // ---- finish: results of the step, after all its tiers ----
    st = post_stream_;
    for (int t = 0; t < 3; t++)
        if (tier_used[t]) VB_CUDA_CHECK(cudaStreamWaitEvent(st, sl.tier_done[t], 0));
    if (cfg_.partials && n_last < L) {
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_partial, sl.dec.partial_words, (size_t)L * kPartialCap * sizeof(int), cudaMemcpyDeviceToHost, st));
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_partial + (size_t)L * kPartialCap, sl.dec.partial_count, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost, st));
        sl.d2h += (size_t)L * (kPartialCap + 1) * sizeof(int);
    }
    if (endpointing_ && n_last < L) {
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_endp, sl.dec.endp_silence, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost, st));
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_endp + L, sl.dec.endp_relcost, (size_t)L * sizeof(float), cudaMemcpyDeviceToHost, st));
        sl.d2h += (size_t)L * 8;
    }
    if (sl.timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[4], st));
    // results of finished lanes
    int k_last = 0;
    for (int i = 0; i < L; i++)
        if (lanes[i].seg_end) {
            const int ch = lanes[i].s->channel;
            VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_cs + k_last, dec_.cs + ch, sizeof(DecChannelState), cudaMemcpyDeviceToHost, st));
            VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_path + (size_t)k_last * path_cap_, dec_.path + (size_t)ch * path_cap_,
                                          (size_t)path_cap_ * sizeof(int), cudaMemcpyDeviceToHost, st));
            sl.d2h += sizeof(DecChannelState) + (size_t)path_cap_ * sizeof(int);
            k_last++;
        }
    VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_load, sl.d_load, (size_t)L * sizeof(int), cudaMemcpyDeviceToHost, st));
    sl.d2h += (size_t)L * sizeof(int);
    if (cfg_.lattice && n_last > 0) {
        // lattice pruning only touches finished channels (nothing else does until they are reused after completion)
        sl.dec.lane_begin = 0;
        sl.dec.lane_end = L;
        if (sl.timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[6], st));
        VB_CUDA_CHECK(vbk_lattice_prune(&sl.dec, st));
        if (sl.timed) VB_CUDA_CHECK(cudaEventRecord(sl.ev[7], st));
        sl.pruned = true;
        sl.launches++;
        VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_lat_hdr, sl.dec.lat_hdr, (size_t)L * sizeof(LatHeader), cudaMemcpyDeviceToHost, st));
        sl.d2h += (size_t)L * sizeof(LatHeader);
    }
    VB_CUDA_CHECK(cudaEventRecord(sl.done, st));
    {
        std::lock_guard<std::mutex> lk(stats_mu_);
        stats_.host_launch_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    }
}

void Engine::complete_step(Slot &sl) {
    std::vector<Lane> &lanes = sl.lanes;
    const int L = (int)lanes.size();
    cudaStream_t st = sl.stream;
    if (!post_threads_.empty() && cfg_.batcher_sleep) {
        // with the host lattice chain running, the host is what bounds throughput: the batcher polls with short sleeps instead of
        // spinning inside the driver, so that its core goes to the lattice pool (costs up to ~0.1 ms per step)
        for (;;) {
            const cudaError_t q = cudaEventQuery(sl.done);
            if (q == cudaSuccess) break;
            if (q != cudaErrorNotReady) VB_CUDA_CHECK(q);
            std::this_thread::sleep_for(std::chrono::microseconds(50));
        }
    }
    VB_CUDA_CHECK(cudaEventSynchronize(sl.done));
    const auto host_t0 = std::chrono::steady_clock::now();
    // load feedback for the tiering of the stream's next chunks: the largest token count of this chunk, or what is left of an
    // earlier peak (token counts swing between chunks, and a lane that lands in too small a tier is the step's critical path)
    for (int i = 0; i < L; i++) lanes[i].s->load = std::max(sl.h_load[i], (int)((long long)lanes[i].s->load * cfg_.load_decay_percent / 100));
    {
        const auto now = std::chrono::steady_clock::now();
        std::lock_guard<std::mutex> lk(stats_mu_);
        for (int i = 0; i < L; i++) {
            if (lanes[i].s->resident) continue;
            if (latencies_ms_.size() < (1u << 22)) latencies_ms_.push_back(std::chrono::duration<float, std::milli>(now - lanes[i].chunk.t_push).count());
        }
    }
    if (endpointing_)
        for (int i = 0; i < L; i++) {
            // kaldi::EndpointDetected on the state after this chunk: rules 1-4 (rule 5 is applied when a chunk is launched)
            lanes[i].endpoint = false;
            const int frames = lanes[i].dec_frames_after - lanes[i].seg_start;
            if (lanes[i].seg_end || frames <= 0 || lanes[i].s->resident) continue;
            const float frame_shift = 0.03f;
            const float utterance_length = frames * frame_shift, trailing_silence = sl.h_endp[i] * frame_shift;
            float relative_cost;
            memcpy(&relative_cost, &sl.h_endp[L + i], 4);
            const bool contains_nonsilence = utterance_length > trailing_silence;
            for (int r = 0; r < 4 && !lanes[i].endpoint; r++)
                lanes[i].endpoint = (contains_nonsilence || !cfg_.ep_must_contain_nonsilence[r]) && trailing_silence >= cfg_.ep_min_trailing_silence[r] &&
                                    relative_cost <= cfg_.ep_max_relative_cost[r] && utterance_length >= cfg_.ep_min_utterance_length[r];
        }
    if (cfg_.partials)
        for (int i = 0; i < L; i++) {
            if (lanes[i].seg_end) continue;
            const int total = sl.h_partial[(size_t)L * kPartialCap + i], n = std::min(total, kPartialCap);
            const int *w = sl.h_partial + (size_t)i * kPartialCap;
            Stream &s = *lanes[i].s;
            std::lock_guard<std::mutex> lk(s.partial_mu);
            s.partial_words.assign(std::reverse_iterator<const int *>(w + n), std::reverse_iterator<const int *>(w));
            s.partial_frames = lanes[i].dec_frames_after;
        }
    // debug capture (tests): copy this step's new rows of the tapped stages
    for (int i = 0; i < L && cfg_.debug_capture; i++) {
        Stream &s = *lanes[i].s;
        if (!s.capture) continue;
        const LaneDesc &d = sl.h_lanes[i];
        Capture &cp = *s.capture;
        auto grab = [&](const NodeDesc &node, int t0, int rows, std::vector<float> &dst) {
            if (rows <= 0) return;
            VB_CUDA_CHECK(vbk_copy_rows(node, d.channel, t0, rows, d_capture_, st));
            VB_CUDA_CHECK(cudaMemcpyAsync(h_capture_, d_capture_, (size_t)rows * node.dim * sizeof(float), cudaMemcpyDeviceToHost, st));
            VB_CUDA_CHECK(cudaStreamSynchronize(st));
            dst.insert(dst.end(), h_capture_, h_capture_ + (size_t)rows * node.dim);
        };
        grab(nodes_[0], d.frames_before, d.frames_after - d.frames_before, cp.mfcc);
        grab(nodes_.back(), d.dec_frames_before * kSubsample, lanes[i].dec_frames_after - d.dec_frames_before, cp.loglikes);
        std::vector<float> iv(model_.ivec_dim);
        VB_CUDA_CHECK(cudaMemcpy(iv.data(), iv_state_.ivec + (size_t)d.channel * model_.ivec_dim, iv.size() * sizeof(float), cudaMemcpyDeviceToHost));
        cp.ivectors.insert(cp.ivectors.end(), iv.begin(), iv.end());
        if (d.last) {
            DecChannelState cs;
            VB_CUDA_CHECK(cudaMemcpy(&cs, dec_.cs + d.channel, sizeof cs, cudaMemcpyDeviceToHost));
            const int nfr = std::min(cs.frame, max_frames_);
            cp.error = cs.error;
            cp.frame_off.resize(nfr + 2);
            VB_CUDA_CHECK(cudaMemcpy(cp.frame_off.data(), dec_.log_frame_off + (size_t)d.channel * (max_frames_ + 2), (nfr + 2) * sizeof(int), cudaMemcpyDeviceToHost));
            const size_t nt = (size_t)cs.log_count, base = (size_t)d.channel * log_cap_;
            cp.tok_state.resize(nt);
            cp.tok_arc.resize(nt);
            cp.tok_prev.resize(nt);
            cp.tok_cost.resize(nt);
            if (nt) {
                VB_CUDA_CHECK(cudaMemcpy(cp.tok_state.data(), dec_.log_state + base, nt * 4, cudaMemcpyDeviceToHost));
                VB_CUDA_CHECK(cudaMemcpy(cp.tok_arc.data(), dec_.log_arc + base, nt * 4, cudaMemcpyDeviceToHost));
                VB_CUDA_CHECK(cudaMemcpy(cp.tok_prev.data(), dec_.log_prev + base, nt * 4, cudaMemcpyDeviceToHost));
                VB_CUDA_CHECK(cudaMemcpy(cp.tok_cost.data(), dec_.log_cost + base, nt * 4, cudaMemcpyDeviceToHost));
            }
        }
    }
    {
        float ms[5] = {0, 0, 0, 0, 0};
        if (sl.timed) {
            for (int k = 0; k < 3; k++) cudaEventElapsedTime(&ms[k], sl.ev[k], sl.ev[k + 1]);
            cudaEventElapsedTime(&ms[3], sl.ev[5], sl.ev[4]);
            if (sl.pruned) cudaEventElapsedTime(&ms[4], sl.ev[6], sl.ev[7]);
        }
        std::lock_guard<std::mutex> lk(stats_mu_);
        stats_.t_feat += ms[0];
        stats_.t_ivec += ms[1];
        stats_.t_nnet += ms[2];
        stats_.t_dec += ms[3];
        stats_.t_prune += ms[4];
        stats_.t_total += ms[0] + ms[1] + ms[2] + ms[3];
        stats_.audio_seconds += sl.audio;
        stats_.steps++;
        stats_.lanes += L;
        stats_.launches += sl.launches;
        stats_.gemm_launches += sl.gemms;
        stats_.resample_segments += sl.resample_segs;
        stats_.dec_launches++;
    }
    std::vector<std::shared_ptr<PackedLattice>> lats;
    const auto fetch_t0 = std::chrono::steady_clock::now();
    if (cfg_.lattice) fetch_lattices(sl, &lats);
    const double fetch_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - fetch_t0).count();
    int n_last = 0;
    for (int i = 0; i < L; i++)
        if (lanes[i].seg_end) finish_lane(sl, lanes[i], n_last++, cfg_.lattice ? lats[i] : nullptr);
    {
        std::lock_guard<std::mutex> lk(stats_mu_);
        stats_.host_fetch_ms += fetch_ms;
        stats_.h2d_bytes += sl.h2d;
        stats_.d2h_bytes += sl.d2h;
        stats_.host_complete_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
    }
}

std::shared_ptr<RawLattice> PackedLattice::unpack() const {
    auto lat = std::make_shared<RawLattice>();
    lat->n_states = n_tok;
    lat->start = start;
    lat->frames = frames;
    lat->error = error;
    const size_t nl = links.size(), nf = finals.size();
    lat->src.resize(nl);
    lat->dst.resize(nl);
    lat->arc.resize(nl);
    lat->acoustic.resize(nl);
    for (size_t k = 0; k < nl; k++) {
        const int4 l = links[k];
        lat->src[k] = l.x;
        lat->dst[k] = l.y;
        lat->arc[k] = l.z;
        memcpy(&lat->acoustic[k], &l.w, 4);
    }
    lat->final_state.resize(nf);
    lat->final_cost.resize(nf);
    for (size_t k = 0; k < nf; k++) {
        lat->final_state[k] = finals[k].x;
        memcpy(&lat->final_cost[k], &finals[k].y, 4);
    }
    lat->state_frame = tok_frame;
    lat->state_graph = tok_state;
    return lat;
}

// Copies the pruned lattices of the step's finished lanes to the host (sizes come from the headers that arrived with the
// step): as many lanes per synchronize as fit the pinned bounce buffer, then plain copies into per-lane vectors.
void Engine::fetch_lattices(Slot &sl, std::vector<std::shared_ptr<PackedLattice>> *out) {
    const int L = (int)sl.lanes.size();
    out->assign(L, nullptr);
    cudaStream_t st = sl.stream;
    const bool with_state = sl.dec.lat_tok_state != nullptr;
    struct Pending {
        int lane;
        size_t links, finals, tok, tok_state;
    };
    std::vector<Pending> batch;
    size_t used = 0;
    const size_t cap = std::max(kLatPoolBytes, (size_t)cfg_.lat_link_cap * sizeof(int4) + (size_t)cfg_.tok_cap * sizeof(int2) + (size_t)2 * cfg_.lat_tok_cap * sizeof(int) + 64);
    auto flush = [&]() {
        if (batch.empty()) return;
        VB_CUDA_CHECK(cudaStreamSynchronize(st));
        for (const Pending &p : batch) {
            PackedLattice &pl = *(*out)[p.lane];
            if (!pl.links.empty()) memcpy(pl.links.data(), sl.h_lat_pool + p.links, pl.links.size() * sizeof(int4));
            if (!pl.finals.empty()) memcpy(pl.finals.data(), sl.h_lat_pool + p.finals, pl.finals.size() * sizeof(int2));
            if (!pl.tok_frame.empty()) memcpy(pl.tok_frame.data(), sl.h_lat_pool + p.tok, pl.tok_frame.size() * sizeof(int));
            if (!pl.tok_state.empty()) memcpy(pl.tok_state.data(), sl.h_lat_pool + p.tok_state, pl.tok_state.size() * sizeof(int));
        }
        batch.clear();
        used = 0;
    };
    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    for (int i = 0; i < L; i++) {
        if (!sl.lanes[i].seg_end) continue;
        const LatHeader h = sl.h_lat_hdr[i];
        auto pl = std::make_shared<PackedLattice>();
        pl->n_tok = h.n_tok;
        pl->start = h.start;
        pl->frames = h.frames;
        pl->error = h.error;
        (*out)[i] = pl;
        if (h.n_tok <= 0) continue;
        const int nl = std::max(0, std::min(h.n_links, cfg_.lat_link_cap)), nf = std::max(0, std::min(h.n_final, cfg_.tok_cap)),
                  nt = std::max(0, std::min(h.n_tok, cfg_.lat_tok_cap));
        pl->links.resize(nl);
        pl->finals.resize(nf);
        pl->tok_frame.resize(nt);
        if (with_state) pl->tok_state.resize(nt);
        const size_t need = align16((size_t)nl * sizeof(int4)) + align16((size_t)nf * sizeof(int2)) + (with_state ? 2 : 1) * align16((size_t)nt * sizeof(int));
        if (used + need > cap) flush();
        Pending p{i, used, 0, 0, 0};
        p.finals = p.links + align16((size_t)nl * sizeof(int4));
        p.tok = p.finals + align16((size_t)nf * sizeof(int2));
        p.tok_state = p.tok + align16((size_t)nt * sizeof(int));
        if (nl) VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_lat_pool + p.links, sl.dec.lat_links + (size_t)i * cfg_.lat_link_cap, (size_t)nl * sizeof(int4), cudaMemcpyDeviceToHost, st));
        if (nf) VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_lat_pool + p.finals, sl.dec.lat_final + (size_t)i * cfg_.tok_cap, (size_t)nf * sizeof(int2), cudaMemcpyDeviceToHost, st));
        if (nt) VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_lat_pool + p.tok, sl.dec.lat_tok_frame + (size_t)i * cfg_.lat_tok_cap, (size_t)nt * sizeof(int), cudaMemcpyDeviceToHost, st));
        if (nt && with_state)
            VB_CUDA_CHECK(cudaMemcpyAsync(sl.h_lat_pool + p.tok_state, sl.dec.lat_tok_state + (size_t)i * cfg_.lat_tok_cap, (size_t)nt * sizeof(int), cudaMemcpyDeviceToHost, st));
        used += need;
        sl.d2h += (size_t)nl * sizeof(int4) + (size_t)nf * sizeof(int2) + (size_t)(with_state ? 2 : 1) * nt * sizeof(int);
        batch.push_back(p);
    }
    flush();
}

void Engine::finish_lane(Slot &sl, Lane &ln, int k, std::shared_ptr<PackedLattice> lat) {
    const DecChannelState &cs = sl.h_cs[k];
    BestPath bp;
    if (lat) {
        bp.packed = std::move(lat);
        if (bp.packed->error && !cs.error)
            log_msg(-1, "stream %llu: lattice capacity error %d", (unsigned long long)ln.s->id, bp.packed->error);
        if (ln.s->capture) ln.s->capture->lattice = bp.packed->unpack();
    }
    bp.cost = cs.best_cost;
    bp.reached_final = cs.reached_final != 0;
    bp.error = cs.error;
    bp.frames = cs.frame;
    bp.offset = ln.seg_offset;
    bp.seq = ln.seg_index;
    const int n = std::min(cs.path_len, path_cap_);
    bp.arcs.resize(n);
    const int *p = sl.h_path + (size_t)k * path_cap_;
    for (int i = 0; i < n; i++) bp.arcs[i] = p[n - 1 - i];
    if (cs.error) log_msg(-1, "stream %llu: decoder capacity error %d (result may be truncated)", (unsigned long long)ln.s->id, cs.error);
    if (cs.error || (bp.packed && bp.packed->error)) {
        std::lock_guard<std::mutex> lk(stats_mu_);
        stats_.truncated++;
    }
    if (!ln.s->on_result) return;
    if (post_threads_.empty()) {
        ln.s->on_result(bp);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        post_outstanding_++;
    }
    {
        std::lock_guard<std::mutex> lk(post_mu_);
        post_queue_.emplace_back([cb = ln.s->on_result, bp = std::move(bp)] { cb(bp); });
    }
    post_cv_.notify_one();
}

double Engine::run_resident(const int16_t *d_audio, int num_streams, int stride, const int *lengths, std::vector<BestPath> *out, int passes) {
    // Device-resident run: the streams' chunk descriptors (lengths only) go through the normal batcher, which reads
    // the samples straight from the resident matrix; no host<->device sample traffic.  With passes > 1 the same streams
    // are decoded `passes` times over as new streams (id = pass * num_streams + row), all queued at once: a stream of pass
    // p+1 starts as soon as a channel is free, and the lattice chain of pass p's results runs beside the search of pass
    // p+1, as in continuous serving; the call returns when every result of every pass has been delivered.
    if (num_streams > cfg_.num_channels) throw std::runtime_error("run_resident: more streams than channels");
    if (passes < 1) passes = 1;
    wait();
    const int spc = samples_per_chunk();
    const int total = num_streams * passes;
    std::vector<std::shared_ptr<Stream>> ss(total);
    std::vector<BestPath> res(num_streams);
    for (int i = 0; i < total; i++) {
        const int row = i % num_streams, pass = i / num_streams;
        const int len = lengths ? lengths[row] : stride;
        if (len < 0 || len > stride) throw std::runtime_error("run_resident: bad stream length");
        ss[i] = std::make_shared<Stream>();
        ss[i]->id = (uint64_t)i;
        ss[i]->resident = true;
        ss[i]->resident_row = row;
        ss[i]->load = cfg_.mid_tokens + 1;
        BestPath *slot = pass == passes - 1 ? &res[row] : nullptr;
        ss[i]->on_result = [slot, row, pass, this](const BestPath &bp) {
            if (slot) *slot = bp;
            if (resident_hook) resident_hook(pass, row, bp);
        };
    }
    cudaEvent_t e0, e1;
    VB_CUDA_CHECK(cudaEventCreate(&e0));
    VB_CUDA_CHECK(cudaEventCreate(&e1));
    {
        std::lock_guard<std::mutex> lk(mu_);
        resident_audio_ = d_audio;
        resident_stride_ = stride;
    }
    VB_CUDA_CHECK(cudaEventRecord(e0, stream_));
    for (int i = 0; i < total; i++) {
        const int len = lengths ? lengths[i % num_streams] : stride;
        const int nfull = len / spc;
        std::lock_guard<std::mutex> lk(mu_);
        for (int k = 0; k <= nfull; k++) {
            Stream::Chunk ck;
            ck.samples.resize(0);
            ck.n_resident = k == nfull ? len - nfull * spc : spc;
            ck.last = k == nfull;
            ss[i]->pending.push_back(std::move(ck));
            ss[i]->pending_chunks.fetch_add(1);
            outstanding_++;
        }
        ss[i]->finished = true;
        ss[i]->queued = true;
        ready_.push_back(ss[i]);
    }
    cv_work_.notify_all();
    wait();
    VB_CUDA_CHECK(cudaEventRecord(e1, stream_));
    VB_CUDA_CHECK(cudaEventSynchronize(e1));
    {
        std::lock_guard<std::mutex> lk(mu_);
        resident_audio_ = nullptr;
        resident_stride_ = 0;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (out) *out = std::move(res);
    return ms;
}

}  // namespace vb
