// vosk_api.cc — extern "C" surface: the reference's batch ABI [REF src/vosk_api.cc:176-282] plus the
// additive entry points of include/vosk_b200.h.  Handles are the C++ objects cast through opaque structs,
// as in the reference [REF src/vosk_api.cc:201,210,224,233]; nothing throws across the boundary.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>

#include "../../include/vosk_b200.h"
#include "batch_model.h"
#include "batch_recognizer.h"
#include "vb_lattice.h"

#include <thread>

namespace vb {
int g_log_level = 0;
void log_msg(int level, const char *fmt, ...) {
    // vosk_set_log_level: 0 info+errors, <0 errors only, >0 verbose [REF src/vosk_api.h:287-294]
    if (level > 0 && g_log_level < level) return;
    if (level == 0 && g_log_level < 0) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    fprintf(stderr, "%s (VoskAPI:b200) %s\n", level < 0 ? "ERROR" : "LOG", buf);
}
}  // namespace vb

static thread_local std::string g_last_error;

extern "C" {

void vosk_set_log_level(int log_level) { vb::g_log_level = log_level; }

void vosk_gpu_init() {
    // reference: CuDevice::SelectGpuId("yes") + AllowMultithreading [REF src/vosk_api.cc:181-189]
    int n = 0;
    if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) cudaFree(nullptr);
}

void vosk_gpu_thread_init() {}

const char *vosk_b200_last_error(void) { return g_last_error.c_str(); }

VoskBatchModel *vosk_batch_model_new_ex(const char *model_dir, const char *options) {
    try {
        g_last_error.clear();
        return (VoskBatchModel *)new BatchModel(model_dir ? model_dir : "model", options ? options : "");
    } catch (const std::exception &e) {
        g_last_error = e.what();
        vb::log_msg(-1, "cannot create batch model: %s", e.what());
        return nullptr;
    } catch (...) {
        g_last_error = "unknown error";
        return nullptr;
    }
}

VoskBatchModel *vosk_batch_model_new() {
    const char *p = getenv("VOSK_BATCH_MODEL_PATH");
    return vosk_batch_model_new_ex(p ? p : "model", "");
}

void vosk_batch_model_free(VoskBatchModel *model) {
    try {
        delete (BatchModel *)model;
    } catch (...) {
    }
}

void vosk_batch_model_wait(VoskBatchModel *model) {
    if (!model) return;
    try {
        ((BatchModel *)model)->WaitForCompletion();
    } catch (...) {
    }
}

VoskBatchRecognizer *vosk_batch_recognizer_new(VoskBatchModel *model, float sample_rate) {
    if (!model || !(sample_rate > 0)) return nullptr;
    try {
        return (VoskBatchRecognizer *)new BatchRecognizer((BatchModel *)model, sample_rate);
    } catch (const std::exception &e) {
        vb::log_msg(-1, "cannot create batch recognizer: %s", e.what());
        return nullptr;
    } catch (...) {
        return nullptr;
    }
}

void vosk_batch_recognizer_free(VoskBatchRecognizer *recognizer) {
    try {
        delete (BatchRecognizer *)recognizer;
    } catch (...) {
    }
}

void vosk_batch_recognizer_accept_waveform(VoskBatchRecognizer *recognizer, const char *data, int length) {
    if (!recognizer || !data || length <= 0) return;
    try {
        ((BatchRecognizer *)recognizer)->AcceptWaveform(data, length);
    } catch (const std::exception &e) {
        vb::log_msg(-1, "accept_waveform: %s", e.what());  // the ABI has no error channel [REF src/vosk_api.h:329]
    } catch (...) {
    }
}

void vosk_batch_recognizer_set_nlsml(VoskBatchRecognizer *recognizer, int nlsml) {
    if (recognizer) ((BatchRecognizer *)recognizer)->SetNLSML(nlsml != 0);
}

void vosk_batch_recognizer_finish_stream(VoskBatchRecognizer *recognizer) {
    if (!recognizer) return;
    try {
        ((BatchRecognizer *)recognizer)->FinishStream();
    } catch (const std::exception &e) {
        vb::log_msg(-1, "finish_stream: %s", e.what());
    } catch (...) {
    }
}

const char *vosk_batch_recognizer_front_result(VoskBatchRecognizer *recognizer) {
    if (!recognizer) return "";
    try {
        return ((BatchRecognizer *)recognizer)->FrontResult();
    } catch (...) {
        return "";
    }
}

void vosk_batch_recognizer_pop(VoskBatchRecognizer *recognizer) {
    if (recognizer) ((BatchRecognizer *)recognizer)->Pop();
}

int vosk_batch_recognizer_get_pending_chunks(VoskBatchRecognizer *recognizer) {
    return recognizer ? ((BatchRecognizer *)recognizer)->GetNumPendingChunks() : 0;
}

// ---------------------------------------------------------------------------------- additive surface
int vosk_batch_model_samples_per_chunk(VoskBatchModel *model) { return model ? ((BatchModel *)model)->samples_per_chunk() : 0; }

int vosk_batch_model_stats(VoskBatchModel *model, double *out, int n) {
    if (!model || !out) return 0;
    BatchModel *bm = (BatchModel *)model;
    double v[81] = {0};
    for (size_t i = 0; i < bm->num_engines(); i++) {
        vb::StepStats s = bm->engine(i).stats();
        v[0] += s.audio_seconds; v[1] += s.steps; v[2] += s.lanes; v[3] += s.launches;
        v[4] += s.tok; v[5] += s.arc_e; v[6] += s.arc_eps; v[7] += s.tok_new;
        v[8] += s.t_feat; v[9] += s.t_ivec; v[10] += s.t_nnet; v[11] += s.t_dec; v[12] += s.gemm_launches;
        v[13] += s.lane_cycles_sum; v[14] = std::max(v[14], (double)s.lane_cycles_max); v[15] = std::max(v[15], (double)s.max_tokens); v[16] += s.lane_launches; v[17] += s.host_launch_ms;
        v[18] += s.arcs_staged; v[19] += s.links; v[20] += s.lat_arcs;
        for (int k = 0; k < 16; k++) v[21 + k] += s.phase[k];
        v[37] += s.resample_segments;
        v[38] += s.truncated; v[39] += s.lattice_fallbacks; v[40] += s.post_ms; v[41] += s.post_jobs; v[42] += bm->engine(i).post_thread_count();
        v[43] += s.t_prune; v[44] += s.host_complete_ms; v[45] += s.host_fetch_ms; v[46] += (double)s.h2d_bytes; v[47] += (double)s.d2h_bytes;
        for (int k = 0; k < 24; k++) v[48 + k] = std::max(v[48 + k], (double)s.phase_slowest[k]);
        for (int t = 0; t < 3; t++) {
            v[72 + t] = std::max(v[72 + t], (double)s.tier_slowest_cycles[t]);
            v[75 + t] = std::max(v[75 + t], (double)s.tier_slowest_tokens[t]);
            v[78 + t] += (double)s.tier_lane_launches[t];
        }
    }
    int k = n < 81 ? n : 81;
    memcpy(out, v, k * sizeof(double));
    return k;
}
void vosk_batch_model_reset_stats(VoskBatchModel *model) {
    if (!model) return;
    BatchModel *bm = (BatchModel *)model;
    for (size_t i = 0; i < bm->num_engines(); i++) bm->engine(i).reset_stats();
}
void vosk_batch_model_set_timing(VoskBatchModel *model, int on) {
    if (!model) return;
    BatchModel *bm = (BatchModel *)model;
    for (size_t i = 0; i < bm->num_engines(); i++) bm->engine(i).set_timing(on != 0);
}

void vosk_batch_model_set_slots(VoskBatchModel *model, int n) {
    if (!model) return;
    BatchModel *bm = (BatchModel *)model;
    for (size_t i = 0; i < bm->num_engines(); i++) bm->engine(i).set_active_slots(n);
}

double vosk_batch_model_run_resident_passes(VoskBatchModel *model, const int16_t *audio, int num_streams, int samples_per_stream,
                                            const int *lengths, int passes, int *mismatches) {
    if (!model || !audio || num_streams <= 0 || samples_per_stream <= 0 || passes < 1) return -1.0;
    BatchModel *bm = (BatchModel *)model;
    int16_t *d_audio = nullptr;
    try {
        vb::Engine &eng = bm->engine(0);
        cudaSetDevice(eng.config().device);
        const size_t bytes = (size_t)num_streams * samples_per_stream * sizeof(int16_t);
        if (cudaMalloc((void **)&d_audio, bytes) != cudaSuccess) throw std::runtime_error("cudaMalloc(audio) failed");
        if (cudaMemcpy(d_audio, audio, bytes, cudaMemcpyHostToDevice) != cudaSuccess) throw std::runtime_error("audio upload failed");
        std::vector<vb::BestPath> res;
        bm->resident_results.assign(num_streams, std::string());
        const vb::Model *m = &bm->model();
        const float lattice_beam = eng.config().lattice_beam;
        const bool host_chain = eng.config().lattice == 1;
        std::vector<std::string> *texts = &bm->resident_results;
        // [pass][stream]: segment index -> text
        auto parts = std::make_shared<std::vector<std::map<int, std::string>>>((size_t)num_streams * passes);
        auto parts_mu = std::make_shared<std::mutex>();
        // result text is produced where the engine delivers results (the lattice pool when lattice=1), inside the timed region
        vb::Engine *engp = &eng;
        eng.resident_hook = [m, lattice_beam, parts, parts_mu, host_chain, engp, num_streams](int pass, int i, const vb::BestPath &bp) {
            std::vector<vb::WordSpan> words;
            bool done = false;
            if (host_chain) {
                const vb::RawLattice *lat = bp.raw_lattice();
                bool ran = false;
                if (lat && lat->error == 0 && lat->n_states > 0) words = vb::lattice_to_words(*lat, *m, lattice_beam, 0.9, nullptr, &ran);
                done = ran || bp.arcs.empty();
                if (!done) engp->count_fallback();
            }
            if (!done) words = vb::align_words(*m, bp.arcs);
            std::string t = vb::result_json(*m, words, bp.offset);
            std::lock_guard<std::mutex> lk(*parts_mu);
            (*parts)[(size_t)pass * num_streams + i][bp.seq] = std::move(t);
        };
        double ms = eng.run_resident(d_audio, num_streams, samples_per_stream, lengths, &res, passes);
        eng.resident_hook = nullptr;
        int bad = 0;
        for (int p = passes - 1; p >= 0; p--)
            for (int i = 0; i < num_streams; i++) {  // segments of one stream (rule-5 endpoints) are concatenated in order
                std::string all;
                for (auto &kv : (*parts)[(size_t)p * num_streams + i]) all += kv.second;
                if (p == passes - 1) (*texts)[i] = std::move(all);
                else if (all != (*texts)[i]) bad++;
            }
        if (mismatches) *mismatches = bad;
        cudaFree(d_audio);
        return ms;
    } catch (const std::exception &e) {
        if (d_audio) cudaFree(d_audio);
        vb::log_msg(-1, "run_resident: %s", e.what());
        return -1.0;
    }
}
double vosk_batch_model_run_resident(VoskBatchModel *model, const int16_t *audio, int num_streams, int samples_per_stream,
                                     const int *lengths) {
    return vosk_batch_model_run_resident_passes(model, audio, num_streams, samples_per_stream, lengths, 1, nullptr);
}
const char *vosk_batch_model_resident_result(VoskBatchModel *model, int stream) {
    if (!model) return "";
    BatchModel *bm = (BatchModel *)model;
    if (stream < 0 || stream >= (int)bm->resident_results.size()) return "";
    return bm->resident_results[stream].c_str();
}

const char *vosk_batch_recognizer_partial_result(VoskBatchRecognizer *recognizer) {
    if (!recognizer) return "";
    try {
        return ((BatchRecognizer *)recognizer)->PartialResult();
    } catch (...) {
        return "";
    }
}
int vosk_batch_recognizer_partial_frames(VoskBatchRecognizer *recognizer) {
    return recognizer ? ((BatchRecognizer *)recognizer)->PartialFrames() : 0;
}
int vosk_batch_model_latency(VoskBatchModel *model, double *out5, int reset) {
    if (!model || !out5) return 0;
    BatchModel *bm = (BatchModel *)model;
    bm->engine(0).latency(out5, reset != 0);  // per engine; engine 0 reported (streams are sharded evenly)
    return 5;
}

void vosk_batch_recognizer_debug_capture(VoskBatchRecognizer *recognizer) {
    if (recognizer) ((BatchRecognizer *)recognizer)->EnableCapture();
}

int64_t vosk_batch_recognizer_debug_get(VoskBatchRecognizer *recognizer, const char *what, void *out, int64_t cap) {
    if (!recognizer || !what) return -1;
    vb::Capture *c = ((BatchRecognizer *)recognizer)->capture();
    if (!c) return -1;
    const void *src = nullptr;
    int64_t bytes = 0;
    std::string w = what;
    auto set = [&](const void *p, size_t n) { src = p; bytes = (int64_t)n; };
    if (w == "mfcc") set(c->mfcc.data(), c->mfcc.size() * 4);
    else if (w == "ivectors") set(c->ivectors.data(), c->ivectors.size() * 4);
    else if (w == "loglikes") set(c->loglikes.data(), c->loglikes.size() * 4);
    else if (w == "frame_off") set(c->frame_off.data(), c->frame_off.size() * 4);
    else if (w == "tok_state") set(c->tok_state.data(), c->tok_state.size() * 4);
    else if (w == "tok_arc") set(c->tok_arc.data(), c->tok_arc.size() * 4);
    else if (w == "tok_prev") set(c->tok_prev.data(), c->tok_prev.size() * 4);
    else if (w == "tok_cost") set(c->tok_cost.data(), c->tok_cost.size() * 4);
    else if (w == "error") set(&c->error, 4);
    else if (w == "lat_hdr" || w == "lat_links" || w == "lat_final" || w == "lat_tok_frame" || w == "lat_tok_state") {
        if (!c->lattice) return 0;
        const vb::RawLattice &L = *c->lattice;
        if (c->lat_hdr.empty()) {
            c->lat_hdr = {L.n_states, (int)L.src.size(), (int)L.final_state.size(), L.start, L.error, L.frames};
            for (size_t k = 0; k < L.src.size(); k++) {
                int ab;
                memcpy(&ab, &L.acoustic[k], 4);
                c->lat_links.insert(c->lat_links.end(), {L.src[k], L.dst[k], L.arc[k], ab});
            }
            for (size_t k = 0; k < L.final_state.size(); k++) {
                int fb;
                memcpy(&fb, &L.final_cost[k], 4);
                c->lat_final.insert(c->lat_final.end(), {L.final_state[k], fb});
            }
        }
        if (w == "lat_hdr") set(c->lat_hdr.data(), c->lat_hdr.size() * 4);
        else if (w == "lat_links") set(c->lat_links.data(), c->lat_links.size() * 4);
        else if (w == "lat_final") set(c->lat_final.data(), c->lat_final.size() * 4);
        else if (w == "lat_tok_frame") set(L.state_frame.data(), L.state_frame.size() * 4);
        else set(L.state_graph.data(), L.state_graph.size() * 4);
    }
    else return -1;
    if (out && cap > 0 && bytes > 0) memcpy(out, src, (size_t)(bytes < cap ? bytes : cap));
    return bytes;
}

// ---------------------------------------------------------------------------------- host-only hooks (no GPU)
int vosk_b200_device_for_stream(unsigned long long stream_id, int num_devices) {
    return num_devices > 0 ? (int)(stream_id % (unsigned long long)num_devices) : 0;
}

int vosk_b200_format_result(const char *const *words, const int *begin, const int *end, const float *conf, int n, float offset, char *out, int cap) {
    std::vector<std::string> ws;
    std::vector<vb::WordSpan> sp;
    for (int i = 0; i < n; i++) {
        ws.push_back(words[i]);
        sp.push_back(vb::WordSpan{i, (float)begin[i], (float)end[i], conf[i]});
    }
    std::string s = vb::result_json_words(ws, sp, offset);
    if (out && cap > 0) {
        size_t k = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
        memcpy(out, s.data(), k);
        out[k] = 0;
    }
    return (int)s.size();
}

int vosk_b200_resample(const float *in, int n, float rate_in, float *out, int cap) {
    vb::LinearResampler rs(rate_in, 16000.0f, std::min(rate_in / 2, 8000.0f), 6);
    std::vector<float> vin(in, in + n), vout;
    rs.resample_flush(vin, &vout);
    for (size_t i = 0; i < vout.size() && (int)i < cap; i++) out[i] = vout[i];
    return (int)vout.size();
}

int vosk_b200_model_check(const char *model_dir, char *out, int cap) {
    try {
        vb::Model m;
        m.load(model_dir ? model_dir : "model");
        vb::Config cfg;
        m.apply_conf(&cfg);
        double wsum = 0;
        for (float w : m.graph.arc_w) wsum += w;
        int n_eps = 0;
        for (int p : m.graph.arc_pdf) n_eps += p < 0;
        snprintf(out, cap, "states=%d arcs=%d eps=%d start=%d wsum=%.3f ops=%zu context=%d pdfs=%d words=%zu beam=%g max_active=%d fpc=%d gauss=%d",
                 m.graph.num_states, m.graph.num_arcs, n_eps, m.graph.start, wsum, m.ops.size(), m.context, m.num_pdfs, m.words.size(),
                 cfg.beam, cfg.max_active, cfg.frames_per_chunk, m.num_gauss);
        return 0;
    } catch (const std::exception &e) {
        snprintf(out, cap, "error: %s", e.what());
        return -1;
    }
}

// Host-only: one object of a loaded model directory as doubles (format-independent view used by the loader tests).
int64_t vosk_b200_model_tensor(const char *model_dir, const char *name, double *out, int64_t cap) {
    try {
        static std::string cached_dir;
        static std::unique_ptr<vb::Model> cached;
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        const std::string dir = model_dir ? model_dir : "model";
        if (!cached || cached_dir != dir) {
            cached.reset();
            std::unique_ptr<vb::Model> m(new vb::Model);
            m->load(dir);
            cached = std::move(m);
            cached_dir = dir;
        }
        const vb::Model &m = *cached;
        std::vector<double> v;
        auto from_tensor = [&](const vb::Tensor *t) {
            if (!t) return;
            const int64_t n = t->numel();
            v.resize((size_t)n);
            for (int64_t i = 0; i < n; i++)
                v[(size_t)i] = t->dtype == 0 ? (double)t->f32()[i] : t->dtype == 1 ? (double)t->i32()[i] : t->dtype == 2 ? t->f64()[i] : (double)t->data[(size_t)i];
        };
        auto opt = [](const vb::TensorMap &tm, const char *k) -> const vb::Tensor * {
            auto it = tm.find(k);
            return it == tm.end() ? nullptr : &it->second;
        };
        const std::string nm = name ? name : "";
        if (nm == "meta") {
            v = {(double)m.ops.size(), (double)m.context, (double)m.num_pdfs, (double)m.ivec_dim, (double)m.bypass_scale, (double)m.prior_offset,
                 (double)m.num_gauss};
        } else if (nm == "tid2pdf") {
            v.assign(m.tid2pdf.begin(), m.tid2pdf.end());
        } else if (nm == "tid2phone") {
            v.assign(m.tid2phone.begin(), m.tid2phone.end());
        } else if (nm.compare(0, 3, "iv.") == 0) {
            const std::string k = nm.substr(3);
            if (k == "lda") from_tensor(opt(m.iv_lda, "lda"));
            else if (k == "cmvn") from_tensor(opt(m.iv_cmvn, "stats"));
            else if (k == "M" || k == "sigma_inv") from_tensor(opt(m.iv_ie, k.c_str()));
            else from_tensor(opt(m.iv_dubm, k.c_str()));
        } else if (nm.compare(0, 2, "op") == 0 && nm.find('.') != std::string::npos) {
            const size_t dot = nm.find('.');
            const int i = std::stoi(nm.substr(2, dot - 2));
            if (i < 0 || i >= (int)m.ops.size()) throw std::runtime_error("no such op: " + nm);
            const vb::AmOp &op = m.ops[(size_t)i];
            const std::string k = nm.substr(dot + 1);
            if (k == "meta") {
                v = {(double)op.in_node, (double)op.byp_node, (double)op.uses_ivec, (double)op.relu_bn, (double)op.K, (double)op.N, (double)op.offs.size()};
                for (int o : op.offs) v.push_back(o);
            } else if (k == "w") from_tensor(op.W);
            else if (k == "b") from_tensor(op.b);
            else if (k == "bn_scale") from_tensor(op.bn_s);
            else if (k == "bn_offset") from_tensor(op.bn_o);
            else throw std::runtime_error("unknown object: " + nm);
        } else {
            throw std::runtime_error("unknown object: " + nm);
        }
        for (int64_t i = 0; i < (int64_t)v.size() && i < cap; i++) out[i] = v[(size_t)i];
        return (int64_t)v.size();
    } catch (const std::exception &e) {
        g_last_error = e.what();
        return -1;
    }
}

// Host-only: the lattice chain of PushLattice on an explicit raw lattice.  stage 0: result text (JSON); 1: determinized
// lattice; 2: word-aligned lattice, as text lines "A src dst word graph acoustic tid,tid,..", "F state graph acoustic tids",
// "S start".
int vosk_b200_lattice_result(const char *model_dir, int n_states, int start, int n_links, const int *src, const int *dst, const int *arc,
                             const float *acoustic, int n_final, const int *final_state, const float *final_cost, float lattice_beam,
                             int stage, char *out, int cap) {
    try {
        static std::string cached_dir;
        static std::unique_ptr<vb::Model> cached;
        static std::mutex mu;
        std::lock_guard<std::mutex> lk(mu);
        const std::string dir = model_dir ? model_dir : "model";
        if (!cached || cached_dir != dir) {
            cached.reset(new vb::Model);
            cached->load(dir);
            cached_dir = dir;
        }
        const vb::Model &m = *cached;
        vb::RawLattice raw;
        raw.n_states = n_states;
        raw.start = start;
        raw.src.assign(src, src + n_links);
        raw.dst.assign(dst, dst + n_links);
        raw.arc.assign(arc, arc + n_links);
        raw.acoustic.assign(acoustic, acoustic + n_links);
        raw.final_state.assign(final_state, final_state + n_final);
        raw.final_cost.assign(final_cost, final_cost + n_final);
        std::string text;
        if (stage == 0) text = vb::result_json(m, vb::lattice_to_words(raw, m, lattice_beam), 0.0f);
        else if (stage == 3) text = vb::result_nlsml(m, vb::lattice_to_words(raw, m, lattice_beam));
        else if (stage == 4) {  // timing / size statistics of the chain on this lattice (repeated to warm the scratch)
            vb::LatticeStats st, one;
            for (int rep = 0; rep < 12; rep++) {  // best of 12 (the first repetitions size the scratch)
                vb::lattice_to_words(raw, m, lattice_beam, 0.9, &one);
                if (rep == 0) st = one;
                st.ms_det = std::min(st.ms_det, one.ms_det);
                st.ms_align = std::min(st.ms_align, one.ms_align);
                st.ms_mbr = std::min(st.ms_mbr, one.ms_mbr);
            }
            char b[512];
            snprintf(b, sizeof b, "{\"raw_states\": %d, \"raw_arcs\": %d, \"det1_states\": %d, \"det1_arcs\": %d, \"det_states\": %d, \"det_arcs\": %d, "
                     "\"ali_states\": %d, \"ali_arcs\": %d, \"mbr_iters\": %d, \"mbr_q\": %d, \"ms_det\": %.3f, \"ms_align\": %.3f, \"ms_mbr\": %.3f}",
                     st.raw_states, st.raw_arcs, st.det1_states, st.det1_arcs, st.det_states, st.det_arcs, st.ali_states, st.ali_arcs, st.mbr_iters,
                     st.mbr_q, st.ms_det, st.ms_align, st.ms_mbr);
            text = b;
        } else text = vb::lattice_debug_text(raw, m, lattice_beam, stage >= 10 ? stage - 10 : stage, stage < 10);
        if (out && cap > 0) {
            size_t k = text.size() < (size_t)cap - 1 ? text.size() : (size_t)cap - 1;
            memcpy(out, text.data(), k);
            out[k] = 0;
        }
        return (int)text.size();
    } catch (const std::exception &e) {
        if (out && cap > 0) snprintf(out, cap, "error: %s", e.what());
        return -1;
    }
}


// Native multi-threaded feeder (measurement helper): drives n streams through the reference ABI calls — one
// vosk_batch_recognizer_new per stream, accept_waveform in bytes_per_call pieces fed round robin inside each feeder thread as
// the reference's driver does [REF python/example/test_gpu_batch.py:27-51], finish_stream, then vosk_batch_model_wait and
// front_result / pop — from `threads` host threads (stream i belongs to thread i % threads).
int vosk_b200_feed_streams(VoskBatchModel *model, const int16_t *const *samples, const int *lengths, int n, int bytes_per_call, int threads,
                           char **results) {
    return vosk_b200_feed_streams_passes(model, samples, lengths, n, bytes_per_call, threads, 1, results, nullptr);
}
// The same with the n streams fed `passes` times over (new recognizers every pass) before the one vosk_batch_model_wait: the
// accept calls never block, so the streams of pass p + 1 queue up behind pass p and start as channels come free, the lattice
// chain of one pass runs beside the search of the next.  results = the last pass's texts; *mismatches (may be NULL) = streams
// of earlier passes whose text differs from it.
int vosk_b200_feed_streams_passes(VoskBatchModel *model, const int16_t *const *samples, const int *lengths, int n, int bytes_per_call, int threads,
                                  int passes, char **results, int *mismatches) {
    if (!model || !samples || !lengths || n <= 0 || bytes_per_call < 2 || passes < 1) return -1;
    try {
        std::vector<VoskBatchRecognizer *> recs((size_t)n * passes, nullptr);
        threads = std::max(1, std::min(threads, n));
        for (int p = 0; p < passes; p++) {
            VoskBatchRecognizer **rp = recs.data() + (size_t)p * n;
            for (int i = 0; i < n; i++) {
                rp[i] = vosk_batch_recognizer_new(model, 16000.0f);
                if (!rp[i]) throw std::runtime_error("recognizer creation failed");
            }
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; t++)
                pool.emplace_back([&, t] {
                    const int per = bytes_per_call / 2;
                    std::vector<long long> pos(n, 0);
                    bool any = true;
                    while (any) {
                        any = false;
                        for (int i = t; i < n; i += threads) {
                            if (pos[i] < 0) continue;
                            const long long left = (long long)lengths[i] - pos[i];
                            if (left <= 0) {
                                vosk_batch_recognizer_finish_stream(rp[i]);
                                pos[i] = -1;
                                continue;
                            }
                            const int take = (int)std::min<long long>(left, per);
                            vosk_batch_recognizer_accept_waveform(rp[i], (const char *)(samples[i] + pos[i]), take * 2);
                            pos[i] += take;
                            any = true;
                        }
                    }
                    for (int i = t; i < n; i += threads)
                        if (pos[i] >= 0) vosk_batch_recognizer_finish_stream(rp[i]);
                });
            for (auto &th : pool) th.join();
        }
        vosk_batch_model_wait(model);
        std::vector<std::string> last(n);
        int bad = 0;
        for (int p = passes - 1; p >= 0; p--)
            for (int i = 0; i < n; i++) {
                VoskBatchRecognizer *r = recs[(size_t)p * n + i];
                std::string all;
                for (;;) {
                    const char *t = vosk_batch_recognizer_front_result(r);
                    if (!t || !*t) break;
                    all += t;
                    vosk_batch_recognizer_pop(r);
                }
                if (p == passes - 1) last[i] = std::move(all);
                else if (all != last[i]) bad++;
                vosk_batch_recognizer_free(r);
            }
        if (results)
            for (int i = 0; i < n; i++) {
                results[i] = (char *)malloc(last[i].size() + 1);
                memcpy(results[i], last[i].c_str(), last[i].size() + 1);
            }
        if (mismatches) *mismatches = bad;
        return 0;
    } catch (const std::exception &e) {
        vb::log_msg(-1, "feed_streams: %s", e.what());
        return -1;
    }
}
void vosk_b200_free(void *p) { free(p); }

// ---- the reference's CPU recognizer / speaker / grammar API [REF src/vosk_api.h:58-285]: outside the accelerated path.
// Exported so that eager binders (JNA Native.register, cgo, P/Invoke) resolve every symbol of the reference header; each
// call logs once per function and returns the "not available" value, as the reference does for its batch half without CUDA
// [REF src/vosk_api.cc:198-282]. ----
static void cpu_api_stub(const char *fn) {
    vb::log_msg(-1, "%s: CPU API not built (this libvosk.so implements the batch / GPU path only)", fn);
}
VoskModel *vosk_model_new(const char *) { cpu_api_stub("vosk_model_new"); return nullptr; }
void vosk_model_free(VoskModel *) {}
int vosk_model_find_word(VoskModel *, const char *) { cpu_api_stub("vosk_model_find_word"); return -1; }
VoskSpkModel *vosk_spk_model_new(const char *) { cpu_api_stub("vosk_spk_model_new"); return nullptr; }
void vosk_spk_model_free(VoskSpkModel *) {}
VoskRecognizer *vosk_recognizer_new(VoskModel *, float) { cpu_api_stub("vosk_recognizer_new"); return nullptr; }
VoskRecognizer *vosk_recognizer_new_spk(VoskModel *, float, VoskSpkModel *) { cpu_api_stub("vosk_recognizer_new_spk"); return nullptr; }
VoskRecognizer *vosk_recognizer_new_grm(VoskModel *, float, const char *) { cpu_api_stub("vosk_recognizer_new_grm"); return nullptr; }
void vosk_recognizer_set_spk_model(VoskRecognizer *, VoskSpkModel *) { cpu_api_stub("vosk_recognizer_set_spk_model"); }
void vosk_recognizer_set_max_alternatives(VoskRecognizer *, int) { cpu_api_stub("vosk_recognizer_set_max_alternatives"); }
void vosk_recognizer_set_words(VoskRecognizer *, int) { cpu_api_stub("vosk_recognizer_set_words"); }
void vosk_recognizer_set_partial_words(VoskRecognizer *, int) { cpu_api_stub("vosk_recognizer_set_partial_words"); }
void vosk_recognizer_set_nlsml(VoskRecognizer *, int) { cpu_api_stub("vosk_recognizer_set_nlsml"); }
int vosk_recognizer_accept_waveform(VoskRecognizer *, const char *, int) { cpu_api_stub("vosk_recognizer_accept_waveform"); return -1; }
int vosk_recognizer_accept_waveform_s(VoskRecognizer *, const short *, int) { cpu_api_stub("vosk_recognizer_accept_waveform_s"); return -1; }
int vosk_recognizer_accept_waveform_f(VoskRecognizer *, const float *, int) { cpu_api_stub("vosk_recognizer_accept_waveform_f"); return -1; }
const char *vosk_recognizer_result(VoskRecognizer *) { cpu_api_stub("vosk_recognizer_result"); return ""; }
const char *vosk_recognizer_partial_result(VoskRecognizer *) { cpu_api_stub("vosk_recognizer_partial_result"); return ""; }
const char *vosk_recognizer_final_result(VoskRecognizer *) { cpu_api_stub("vosk_recognizer_final_result"); return ""; }
void vosk_recognizer_reset(VoskRecognizer *) {}
void vosk_recognizer_free(VoskRecognizer *) {}

}  // extern "C"
