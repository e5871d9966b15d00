// batch_model.cc — see batch_model.h.  Replaces [REF src/batch_model.cc].
#include "batch_model.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <sstream>

using namespace vb;

static void apply_option(Config *c, const std::string &k, const std::string &v) {
    auto I = [&](int *dst) { *dst = std::stoi(v); };
    auto F = [&](float *dst) { *dst = std::stof(v); };
    if (k == "frames-per-chunk") I(&c->frames_per_chunk);
    else if (k == "max-batch-size") I(&c->max_lanes);
    else if (k == "num-channels") I(&c->num_channels);
    else if (k == "beam") F(&c->beam);
    else if (k == "lattice-beam") F(&c->lattice_beam);
    else if (k == "max-active") I(&c->max_active);
    else if (k == "min-active") I(&c->min_active);
    else if (k == "tok-cap") I(&c->tok_cap);
    else if (k == "cand-cap") I(&c->cand_cap);
    else if (k == "hash-size") I(&c->hash_size);
    else if (k == "max-seconds") I(&c->max_seconds);
    else if (k == "log-tokens-per-frame") I(&c->log_tokens_per_frame);
    else if (k == "tensor-cores") I(&c->use_tensor_cores);
    else if (k == "pipeline-slots") I(&c->pipeline_slots);
    else if (k == "device-resample") I(&c->device_resample);
    else if (k == "heavy-tokens") I(&c->heavy_tokens);
    else if (k == "mid-tokens") I(&c->mid_tokens);
    else if (k == "load-decay") I(&c->load_decay_percent);
    else if (k == "mid-threads") I(&c->mid_threads);
    else if (k == "heavy-threads") I(&c->heavy_threads);
    else if (k == "light-threads") I(&c->light_threads);
    else if (k == "debug-capture") I(&c->debug_capture);
    else if (k == "lattice") I(&c->lattice);
    else if (k == "post-threads") I(&c->post_threads);
    else if (k == "batcher-sleep") I(&c->batcher_sleep);
    else if (k == "partials") I(&c->partials);
    else if (k == "fe-priority") I(&c->fe_priority);
    else if (k == "fe-split") I(&c->fe_split);
    else if (k == "endpoint-rule5-seconds") F(&c->endpoint_rule5_seconds);
    else if (k == "log-links-per-frame") I(&c->log_links_per_frame);
    else if (k == "lat-tok-cap") I(&c->lat_tok_cap);
    else if (k == "lat-link-cap") I(&c->lat_link_cap);
    else if (k == "model-conf") I(&c->model_conf);
    else if (k == "acoustic-scale") F(&c->acoustic_scale);
    else if (k == "devices") {}  // handled by the caller
    else throw std::runtime_error("unknown batch option '" + k + "'");
}

static std::vector<std::pair<std::string, std::string>> split_options(const std::string &s) {
    std::vector<std::pair<std::string, std::string>> out;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, ',')) {
        if (item.empty()) continue;
        size_t eq = item.find('=');
        if (eq == std::string::npos) throw std::runtime_error("bad option '" + item + "' (want key=value)");
        out.emplace_back(item.substr(0, eq), item.substr(eq + 1));
    }
    return out;
}

BatchModel::BatchModel(const std::string &model_dir, const std::string &options) {
    model_.load(model_dir);
    // hard-coded reference values are the defaults of vb::Config [REF src/batch_model.cc:69-88]; the feature / i-vector conf
    // files, then (only with model-conf=1) model.conf, then VOSK_BATCH_OPTIONS, then explicit options refine them.
    std::string devices = "0";
    if (const char *e = getenv("VOSK_BATCH_DEVICES")) devices = e;
    std::string all = options;
    if (const char *e = getenv("VOSK_BATCH_OPTIONS")) all = std::string(e) + "," + all;
    const auto opts = split_options(all);
    for (auto &kv : opts)
        if (kv.first == "model-conf") apply_option(&cfg_, kv.first, kv.second);
    model_.apply_conf(&cfg_);
    for (auto &kv : opts) {
        if (kv.first == "devices") devices = kv.second;
        apply_option(&cfg_, kv.first, kv.second);
    }
    // tensor-cores: 1 = fp16 hi/lo operand split (default), 2 = TF32 hi/lo split — two operand formats of the one tcgen05 kernel.
    // 0 selects the plain fp32 FFMA kernel that the parity tests use as an arithmetic reference: only with the test taps on.
    if (cfg_.use_tensor_cores < 0 || cfg_.use_tensor_cores > 2 || (cfg_.use_tensor_cores == 0 && !cfg_.debug_capture))
        throw std::runtime_error("tensor-cores must be 1 or 2 (0, the fp32 reference kernel, needs debug-capture=1)");
    std::vector<int> devs;
    if (devices == "all") {
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) throw std::runtime_error("no CUDA device visible");
        for (int i = 0; i < n; i++) devs.push_back(i);
    } else {
        std::stringstream ss(devices);
        std::string item;
        while (std::getline(ss, item, ':')) devs.push_back(std::stoi(item));
    }
    if (devs.empty()) devs.push_back(0);
    log_msg(0, "Decoding params beam=%g max-active=%d lattice-beam=%g", cfg_.beam, cfg_.max_active, cfg_.lattice_beam);
    for (int d : devs) {
        Config c = cfg_;
        c.device = d;
        c.num_engines = (int)devs.size();
        engines_.emplace_back(new Engine(model_, c));
    }
    samples_per_chunk_ = engines_[0]->samples_per_chunk();  // [REF src/batch_model.cc:98]
}

BatchModel::~BatchModel() {
    // engines (and their worker threads) go first; the model they borrow outlives them
    engines_.clear();
}

uint64_t BatchModel::GetID(BatchRecognizer *) { return last_id_.fetch_add(1); }

void BatchModel::WaitForCompletion() {
    for (auto &e : engines_) e->wait();
}
