// vb_model.h — host-side model: the files BatchModel::BatchModel() loads [REF src/batch_model.cc:26-67,75-77].
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "vb_common.h"

namespace vb {

struct Tensor {
    int dtype = 0;  // 0 f32, 1 i32, 2 f64, 3 u8
    std::vector<int64_t> shape;
    std::vector<uint8_t> data;
    int64_t numel() const {
        int64_t n = 1;
        for (auto d : shape) n *= d;
        return n;
    }
    const float *f32() const { return reinterpret_cast<const float *>(data.data()); }
    const int32_t *i32() const { return reinterpret_cast<const int32_t *>(data.data()); }
    const double *f64() const { return reinterpret_cast<const double *>(data.data()); }
};
using TensorMap = std::map<std::string, Tensor>;

TensorMap read_vbt(const std::string &path);
std::map<std::string, std::string> read_conf(const std::string &path);  // --key=value lines

// decoding graph in the canonical CSR order: per state emitting arcs first (file order), then epsilons
struct Graph {
    int num_states = 0, num_arcs = 0, start = 0;
    std::vector<float> final_cost;
    std::vector<int32_t> e_begin;    // [S+1]
    std::vector<int32_t> eps_begin;  // [S]
    std::vector<float> arc_w;
    std::vector<int32_t> arc_next, arc_pdf, arc_ilabel, arc_olabel;
    bool has_negative_eps = false;
};
Graph read_graph(const std::string &fst_path, const std::vector<int32_t> &tid2pdf);

struct AmOp {
    std::string name;
    int in_node, byp_node;
    std::vector<int> offs;
    bool uses_ivec, relu_bn;
    int K, N;
    const Tensor *W, *b, *bn_s, *bn_o;
};

struct Model {
    std::string dir;
    std::map<std::string, std::string> conf;
    TensorMap am;
    int feat_dim = 40, ivec_dim = 40, hidden = 0, bottleneck = 0, prefinal_small = 0, prefinal_big = 0, num_pdfs = 0;
    std::vector<int> strides;
    float bypass_scale = 0.75f;
    int context = 0;  // frames of left == right model context
    std::vector<AmOp> ops;
    std::vector<int> node_dim;  // node 0 = MFCC input, node i+1 = output of op i
    std::vector<int32_t> tid2pdf, tid2phone;
    // per transition-id: bit 0 self-loop (TransitionModel::IsSelfLoop), bit 1 final = enters the HMM's final state (IsFinal),
    // bit 2 leaves HMM state 0 (TransitionIdToHmmState == 0); index 0 unused
    std::vector<uint8_t> tid_flags;
    Graph graph;
    std::vector<std::string> words;
    std::vector<int> phone_type;  // 0 none, 1 nonword, 2 begin, 3 end, 4 internal, 5 singleton
    // i-vector extractor
    TensorMap iv_lda, iv_dubm, iv_ie, iv_cmvn;
    int num_gauss = 0;
    float prior_offset = 0.f;

    void load(const std::string &model_dir);  // throws std::runtime_error on a missing/corrupt file
    void load_vbt_am(const std::string &mdl);    // am/final.mdl as the generator's tensor container (collapsed network)
    void load_kaldi_am(const std::string &mdl);  // am/final.mdl as Kaldi's TransitionModel + nnet3 file (vb_kaldi.cc)
    void pad_dimensions();  // i-vector dim -> multiple of 4, output width -> multiple of 16 (exact: zero columns / rows)
    // feature / i-vector conf files always; decoding parameters, batch sizes and silence endpointing of model.conf only
    // when cfg->model_conf is set (the reference's batch path never reads model.conf)
    void apply_conf(Config *cfg) const;
};

}  // namespace vb
