// vb_kaldi.h — readers for Kaldi's on-disk formats of the model files BatchModel::BatchModel() loads
// [REF src/batch_model.cc:28-67,76-77]: final.mdl (TransitionModel + nnet3 AmNnetSimple), final.mat, final.dubm,
// final.ie, global_cmvn.stats.  Host only.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "vb_model.h"

namespace vb {

// true when the file starts with Kaldi's binary marker "\0B"
bool kaldi_is_binary(const std::string &path);
// true when the file is one of ours (VBT1 container)
bool file_is_vbt(const std::string &path);

// A Kaldi Matrix<float|double> file, binary ("\0B" + FM/DM) or text (" [ a b\n c d ]").
struct KaldiMatrix {
    int rows = 0, cols = 0;
    std::vector<double> v;
};
KaldiMatrix read_kaldi_matrix_file(const std::string &path);

// DiagGmm -> tensors gconsts[G], weights[G], means_invvars[G][F], inv_vars[G][F] (f32); gconsts are recomputed when absent.
TensorMap read_kaldi_dubm(const std::string &path);
// IvectorExtractor -> M[G][F][D], sigma_inv[G][F][F] (full symmetric), w[G], prior_offset[1] (f32)
TensorMap read_kaldi_ie(const std::string &path);

// One op of the collapsed acoustic model, in the engine's terms (vb_model.h AmOp): out = epilogue(W * splice(in_node, offs) [ivector])
struct KaldiOp {
    std::string name;
    int in_node = 0, byp_node = -1;
    std::vector<int> offs;
    bool uses_ivec = false, relu = false;
    int K = 0, N = 0;
    float byp_scale = 0.f;
    std::vector<float> W, b, bn_s, bn_o;  // b empty = no bias; bn_* present whenever relu is set (identity when the net has none)
};
struct KaldiAm {
    int feat_dim = 0, ivec_dim = 0, num_pdfs = 0, left_context = 0, right_context = 0;
    std::vector<KaldiOp> ops;            // op i produces node i + 1; node 0 is the "input" node
    std::vector<int32_t> tid2pdf, tid2phone;  // index = transition-id (entry 0 unused: -1 / 0)
    std::vector<uint8_t> tid_flags;           // bit 0 self-loop, bit 1 enters the final HMM state, bit 2 leaves HMM state 0
};
// Reads final.mdl and compiles the nnet3 graph feeding output-node "output" into the collapsed op chain: every linear,
// test-mode batchnorm, identity (dropout / no-op / spec-augment) component and every Append / Offset / Sum / Scale
// descriptor is folded into the weights of the next affine (what CollapseModel + the nnet3 compiler do
// [REF src/batch_model.cc:46-48]).  Throws std::runtime_error on anything it cannot express.
KaldiAm read_kaldi_final_mdl(const std::string &path);

}  // namespace vb
