// vb_model.cc — model directory loader (host).  File set and relative paths as hard-coded by the
// reference [REF src/batch_model.cc:28-37,75-77]; conf syntax = Kaldi ParseOptions "--key=value"
// [REF src/batch_model.cc:26-28].  HCLG.fst: OpenFst binary, "vector" or "const", StdArc.
#include "vb_model.h"

#include "vb_kaldi.h"

#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

namespace vb {

static std::vector<uint8_t> slurp(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<uint8_t> buf((size_t)n);
    if (n && !f.read(reinterpret_cast<char *>(buf.data()), n)) throw std::runtime_error("cannot read " + path);
    return buf;
}

TensorMap read_vbt(const std::string &path) {
    std::vector<uint8_t> buf = slurp(path);
    if (buf.size() < 8 || memcmp(buf.data(), "VBT1", 4) != 0) throw std::runtime_error("not a VBT1 container: " + path);
    size_t p = 4;
    auto rd32 = [&]() {
        if (p + 4 > buf.size()) throw std::runtime_error("truncated " + path);
        uint32_t v;
        memcpy(&v, &buf[p], 4);
        p += 4;
        return v;
    };
    auto rd64 = [&]() {
        if (p + 8 > buf.size()) throw std::runtime_error("truncated " + path);
        uint64_t v;
        memcpy(&v, &buf[p], 8);
        p += 8;
        return v;
    };
    static const int item[4] = {4, 4, 8, 1};
    TensorMap out;
    uint32_t n = rd32();
    for (uint32_t i = 0; i < n; i++) {
        uint32_t ln = rd32();
        if (p + ln > buf.size()) throw std::runtime_error("truncated " + path);
        std::string name((const char *)&buf[p], ln);
        p += ln;
        Tensor t;
        t.dtype = (int)rd32();
        if (t.dtype < 0 || t.dtype > 3) throw std::runtime_error("bad dtype in " + path);
        uint32_t nd = rd32();
        for (uint32_t d = 0; d < nd; d++) t.shape.push_back((int64_t)rd64());
        size_t bytes = (size_t)t.numel() * item[t.dtype];
        if (p + bytes > buf.size()) throw std::runtime_error("truncated tensor " + name + " in " + path);
        t.data.assign(buf.begin() + p, buf.begin() + p + bytes);
        p += bytes;
        out.emplace(name, std::move(t));
    }
    return out;
}

std::map<std::string, std::string> read_conf(const std::string &path) {
    std::map<std::string, std::string> out;
    std::ifstream f(path);
    std::string line;
    while (std::getline(f, line)) {
        size_t h = line.find('#');
        if (h != std::string::npos) line.resize(h);
        size_t a = line.find_first_not_of(" \t\r");
        if (a == std::string::npos || line.compare(a, 2, "--") != 0) continue;
        size_t e = line.find_last_not_of(" \t\r");
        line = line.substr(a + 2, e - a - 1);
        size_t eq = line.find('=');
        if (eq == std::string::npos) out[line] = "true";
        else out[line.substr(0, eq)] = line.substr(eq + 1);
    }
    return out;
}

namespace {
struct FileArc {
    int32_t ilabel, olabel;
    float weight;
    int32_t next;
};
}  // namespace

Graph read_graph(const std::string &path, const std::vector<int32_t> &tid2pdf) {
    std::vector<uint8_t> buf = slurp(path);
    size_t p = 0;
    auto need = [&](size_t n) {
        if (p + n > buf.size()) throw std::runtime_error("truncated FST " + path);
    };
    auto rdi32 = [&]() { need(4); int32_t v; memcpy(&v, &buf[p], 4); p += 4; return v; };
    auto rdi64 = [&]() { need(8); int64_t v; memcpy(&v, &buf[p], 8); p += 8; return v; };
    auto rdstr = [&]() { int32_t n = rdi32(); need((size_t)n); std::string s((const char *)&buf[p], n); p += n; return s; };
    if (rdi32() != 2125659606) throw std::runtime_error("bad FST magic in " + path);
    std::string ftype = rdstr(), atype = rdstr();
    if (atype != "standard") throw std::runtime_error("unsupported arc type " + atype);
    int32_t version = rdi32(), flags = rdi32();
    rdi64();  // properties
    int64_t start = rdi64(), ns = rdi64(), na = rdi64();
    if (flags & 3) throw std::runtime_error("FST with embedded symbol tables is not supported: " + path);
    std::vector<float> fin((size_t)ns);
    std::vector<int64_t> row((size_t)ns + 1, 0);
    std::vector<FileArc> arcs;
    if (ftype == "const") {
        if (version == 1 && (p % 16)) p += 16 - p % 16;
        need((size_t)ns * 20);
        for (int64_t s = 0; s < ns; s++) {
            uint32_t pos, narcs;
            memcpy(&fin[s], &buf[p], 4);
            memcpy(&pos, &buf[p + 4], 4);
            memcpy(&narcs, &buf[p + 8], 4);
            row[s] = pos;
            (void)narcs;
            p += 20;
        }
        row[ns] = na;
        if (version == 1 && (p % 16)) p += 16 - p % 16;
        need((size_t)na * 16);
        arcs.resize((size_t)na);
        memcpy(arcs.data(), &buf[p], (size_t)na * 16);
    } else if (ftype == "vector") {
        arcs.reserve((size_t)na);
        for (int64_t s = 0; s < ns; s++) {
            need(12);
            memcpy(&fin[s], &buf[p], 4);
            p += 4;
            int64_t n = rdi64();
            need((size_t)n * 16);
            size_t old = arcs.size();
            arcs.resize(old + (size_t)n);
            memcpy(&arcs[old], &buf[p], (size_t)n * 16);
            p += (size_t)n * 16;
            row[s + 1] = (int64_t)arcs.size();
        }
        na = (int64_t)arcs.size();
    } else {
        throw std::runtime_error("unsupported FST type " + ftype);
    }
    if (na >= (1ll << 31) || ns >= (1ll << 31)) throw std::runtime_error("graph too large for 32-bit arc ids");
    Graph g;
    g.num_states = (int)ns;
    g.num_arcs = (int)na;
    g.start = (int)start;
    g.final_cost = fin;
    g.e_begin.resize(ns + 1);
    g.eps_begin.resize(ns);
    g.arc_w.resize(na);
    g.arc_next.resize(na);
    g.arc_pdf.resize(na);
    g.arc_ilabel.resize(na);
    g.arc_olabel.resize(na);
    int64_t w = 0;
    for (int64_t s = 0; s < ns; s++) {
        g.e_begin[s] = (int32_t)w;
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) g.eps_begin[s] = (int32_t)w;
            for (int64_t a = row[s]; a < row[s + 1]; a++) {
                const FileArc &fa = arcs[a];
                if ((fa.ilabel == 0) != (pass == 1)) continue;
                if (fa.ilabel < 0 || fa.ilabel >= (int)tid2pdf.size()) throw std::runtime_error("ilabel out of range in " + path);
                if (fa.next < 0 || fa.next >= ns) throw std::runtime_error("nextstate out of range in " + path);
                g.arc_w[w] = fa.weight;
                g.arc_next[w] = fa.next;
                g.arc_ilabel[w] = fa.ilabel;
                g.arc_olabel[w] = fa.olabel;
                g.arc_pdf[w] = fa.ilabel ? tid2pdf[fa.ilabel] : -1;
                if (!fa.ilabel && fa.weight < 0) g.has_negative_eps = true;
                w++;
            }
        }
    }
    g.e_begin[ns] = (int32_t)na;
    return g;
}

static const Tensor *find(const TensorMap &m, const std::string &k, bool required = true) {
    auto it = m.find(k);
    if (it == m.end()) {
        if (required) throw std::runtime_error("model tensor missing: " + k);
        return nullptr;
    }
    return &it->second;
}

static Tensor make_f32(const std::vector<float> &v, std::vector<int64_t> shape) {
    Tensor t;
    t.dtype = 0;
    t.shape = std::move(shape);
    t.data.resize(v.size() * 4);
    if (!v.empty()) memcpy(t.data.data(), v.data(), v.size() * 4);
    return t;
}

// final.mdl in Kaldi's own format (TransitionModel + nnet3): compiled into the same op chain (vb_kaldi.cc)
void Model::load_kaldi_am(const std::string &mdl) {
    KaldiAm k = read_kaldi_final_mdl(mdl);
    feat_dim = k.feat_dim;
    ivec_dim = k.ivec_dim;
    num_pdfs = k.num_pdfs;
    if (feat_dim != kNumCeps) throw std::runtime_error("feat-dim must be 40");
    if (k.left_context != k.right_context)
        throw std::runtime_error("asymmetric model context (" + std::to_string(k.left_context) + ", " + std::to_string(k.right_context) + ") is not supported");
    context = k.left_context;
    if ((int)k.ops.size() + 1 > kMaxNodes) throw std::runtime_error("too many layers");
    node_dim.push_back(feat_dim);
    bool have_scale = false;
    for (size_t i = 0; i < k.ops.size(); i++) {
        KaldiOp &o = k.ops[i];
        const std::string base = "op" + std::to_string(i);
        AmOp op;
        op.name = o.name;
        op.in_node = o.in_node;
        op.byp_node = o.byp_node;
        op.offs = o.offs;
        op.uses_ivec = o.uses_ivec;
        op.relu_bn = o.relu;
        op.K = o.K;
        op.N = o.N;
        if ((int)op.offs.size() > kMaxOffsets) throw std::runtime_error("too many time offsets in " + o.name);
        op.W = &(am[base + ".w"] = make_f32(o.W, {o.N, o.K}));
        op.b = o.b.empty() ? nullptr : &(am[base + ".b"] = make_f32(o.b, {o.N}));
        op.bn_s = o.relu ? &(am[base + ".bn_scale"] = make_f32(o.bn_s, {o.N})) : nullptr;
        op.bn_o = o.relu ? &(am[base + ".bn_offset"] = make_f32(o.bn_o, {o.N})) : nullptr;
        if (o.byp_node >= 0) {
            if (have_scale && o.byp_scale != bypass_scale) throw std::runtime_error("layers with different bypass scales are not supported");
            bypass_scale = o.byp_scale;
            have_scale = true;
            if (node_dim[o.byp_node] != o.N) throw std::runtime_error("bypass dimension mismatch in " + o.name);
        }
        ops.push_back(op);
        node_dim.push_back(o.N);
    }
    tid2pdf = std::move(k.tid2pdf);
    tid2phone = std::move(k.tid2phone);
    tid_flags = std::move(k.tid_flags);
}

void Model::load_vbt_am(const std::string &mdl) {
    am = read_vbt(mdl);
    const Tensor *cfgt = find(am, "config");
    std::map<std::string, std::string> c;
    {
        std::istringstream ss(std::string((const char *)cfgt->data.data(), cfgt->data.size()));
        std::string line;
        while (std::getline(ss, line)) {
            size_t sp = line.find(' ');
            if (sp != std::string::npos) c[line.substr(0, sp)] = line.substr(sp + 1);
        }
    }
    if (c["arch"] != "tdnnf") throw std::runtime_error("unsupported acoustic model arch '" + c["arch"] + "'");
    feat_dim = std::stoi(c["feat-dim"]);
    ivec_dim = std::stoi(c["ivector-dim"]);
    hidden = std::stoi(c["hidden-dim"]);
    bottleneck = std::stoi(c["bottleneck-dim"]);
    prefinal_small = std::stoi(c["prefinal-small"]);
    prefinal_big = std::stoi(c["prefinal-big"]);
    num_pdfs = std::stoi(c["num-pdfs"]);
    bypass_scale = std::stof(c["bypass-scale"]);
    if (feat_dim != kNumCeps) throw std::runtime_error("feat-dim must be 40");
    {
        std::istringstream ss(c["tdnnf-strides"]);
        int s;
        while (ss >> s) strides.push_back(s);
    }
    context = 2;
    for (int s : strides) context += s;
    // op list of the collapsed network  [REF training/local/chain/run_tdnn.sh:98-129]
    auto add = [&](const std::string &name, int in_node, std::vector<int> offs, bool relu_bn, int byp, bool iv,
                   const std::string &bn_name) {
        AmOp op;
        op.name = name;
        op.in_node = in_node;
        op.byp_node = byp;
        op.offs = offs;
        op.uses_ivec = iv;
        op.relu_bn = relu_bn;
        op.W = find(am, name + ".w");
        op.b = find(am, name + ".b", false);
        op.bn_s = relu_bn ? find(am, bn_name + ".bn_scale") : nullptr;
        op.bn_o = relu_bn ? find(am, bn_name + ".bn_offset") : nullptr;
        op.N = (int)op.W->shape[0];
        op.K = (int)op.W->shape[1];
        int expectK = node_dim[in_node] * (int)offs.size() + (iv ? ivec_dim : 0);
        if (op.K != expectK) throw std::runtime_error("weight shape mismatch for " + name);
        ops.push_back(op);
        node_dim.push_back(op.N);
    };
    node_dim.push_back(feat_dim);
    add("tdnn1", 0, {-2, -1, 0, 1, 2}, true, -1, true, "tdnn1");
    int cur = 1;
    for (size_t k = 0; k < strides.size(); k++) {
        int s = strides[k];
        std::string nm = "tdnnf" + std::to_string(k + 2);
        add(nm + ".linear", cur, s ? std::vector<int>{-s, 0} : std::vector<int>{0}, false, -1, false, "");
        add(nm + ".affine", cur + 1, s ? std::vector<int>{0, s} : std::vector<int>{0}, true, cur, false, nm);
        cur += 2;
    }
    add("prefinal_l", cur, {0}, false, -1, false, "");
    add("prefinal.affine", cur + 1, {0}, true, -1, false, "prefinal");
    add("prefinal.linear", cur + 2, {0}, false, -1, false, "");
    add("output", cur + 3, {0}, false, -1, false, "");
    if ((int)node_dim.size() > kMaxNodes) throw std::runtime_error("too many layers");
    if (node_dim.back() != num_pdfs) throw std::runtime_error("output dim != num-pdfs");
    const Tensor *t2p = find(am, "tid2pdf"), *t2ph = find(am, "tid2phone");
    tid2pdf.assign(t2p->i32(), t2p->i32() + t2p->numel());
    tid2phone.assign(t2ph->i32(), t2ph->i32() + t2ph->numel());
    // the container's transition model is the chain topology of the generator: one emitting HMM state per phone,
    // tid = 2*tstate+1 its self-loop, 2*tstate+2 the forward transition into the final state
    tid_flags.assign(tid2phone.size(), 0);
    for (size_t tid = 1; tid < tid_flags.size(); tid++) tid_flags[tid] = (uint8_t)(4 | ((tid % 2) == 1 ? 1 : 2));
}

void Model::load(const std::string &d) {
    dir = d;
    conf = read_conf(d + "/conf/model.conf");
    // am/final.mdl: Kaldi's TransitionModel + nnet3 file, or the generator's tensor container
    const std::string mdl = d + "/am/final.mdl";
    if (!std::ifstream(mdl, std::ios::binary)) throw std::runtime_error("cannot open " + mdl);
    if (file_is_vbt(mdl)) load_vbt_am(mdl);
    else if (kaldi_is_binary(mdl)) load_kaldi_am(mdl);
    else throw std::runtime_error("unrecognised acoustic model file (neither Kaldi binary nor VBT1): " + mdl);
    graph = read_graph(d + "/graph/HCLG.fst", tid2pdf);
    for (int32_t pdf : graph.arc_pdf)
        if (pdf >= num_pdfs) throw std::runtime_error("pdf id out of range in graph");
    {
        std::ifstream f(d + "/graph/words.txt");
        if (!f) throw std::runtime_error("cannot open " + d + "/graph/words.txt");
        std::string w;
        long id;
        while (f >> w >> id) {
            if (id < 0) continue;
            if ((size_t)id >= words.size()) words.resize(id + 1);
            words[id] = w;
        }
    }
    {
        std::ifstream f(d + "/graph/phones/word_boundary.int");
        int ph;
        std::string kind;
        while (f >> ph >> kind) {
            if (ph < 0) continue;
            if ((size_t)ph >= phone_type.size()) phone_type.resize(ph + 1, 0);
            phone_type[ph] = kind == "nonword" ? 1 : kind == "begin" ? 2 : kind == "end" ? 3 : kind == "internal" ? 4 : kind == "singleton" ? 5 : 0;
        }
    }
    // i-vector extractor files [REF src/model.cc:251-256]: Kaldi formats or the generator's container, file by file
    auto matrix_tensor = [](const KaldiMatrix &km, bool f64) {
        Tensor t;
        t.dtype = f64 ? 2 : 0;
        t.shape = {km.rows, km.cols};
        t.data.resize(km.v.size() * (f64 ? 8 : 4));
        if (f64) memcpy(t.data.data(), km.v.data(), km.v.size() * 8);
        else
            for (size_t i = 0; i < km.v.size(); i++) reinterpret_cast<float *>(t.data.data())[i] = (float)km.v[i];
        return t;
    };
    const std::string ivd = d + "/ivector/";
    if (file_is_vbt(ivd + "final.mat")) iv_lda = read_vbt(ivd + "final.mat");
    else iv_lda["lda"] = matrix_tensor(read_kaldi_matrix_file(ivd + "final.mat"), false);
    iv_dubm = file_is_vbt(ivd + "final.dubm") ? read_vbt(ivd + "final.dubm") : read_kaldi_dubm(ivd + "final.dubm");
    iv_ie = file_is_vbt(ivd + "final.ie") ? read_vbt(ivd + "final.ie") : read_kaldi_ie(ivd + "final.ie");
    if (file_is_vbt(ivd + "global_cmvn.stats")) iv_cmvn = read_vbt(ivd + "global_cmvn.stats");
    else iv_cmvn["stats"] = matrix_tensor(read_kaldi_matrix_file(ivd + "global_cmvn.stats"), true);
    {
        const Tensor *lda = find(iv_lda, "lda"), *st = find(iv_cmvn, "stats");
        if (lda->shape.size() != 2 || lda->shape[0] != feat_dim) throw std::runtime_error("final.mat: unexpected shape");
        if (st->shape.size() != 2 || st->shape[0] != 2 || st->shape[1] != feat_dim + 1) throw std::runtime_error("global_cmvn.stats: unexpected shape");
    }
    num_gauss = (int)find(iv_dubm, "gconsts")->numel();
    prior_offset = find(iv_ie, "prior_offset")->f32()[0];
    if (find(iv_ie, "M")->shape[2] != ivec_dim) throw std::runtime_error("i-vector dim mismatch");
    pad_dimensions();
}

// The kernels move rows in 16-byte pieces and tile outputs by 16: an i-vector dimension that is not a multiple of 4 (30 is
// common) and an output layer that is not a multiple of 16 wide are padded here, exactly: extra i-vector dimensions get zero
// columns in the extractor (their prior keeps them at 0) and zero columns in the weights that read them; extra outputs get
// zero rows (no pdf id refers to them).  Any other odd layer size is rejected.
void Model::pad_dimensions() {
    auto padded_rows = [&](const Tensor &w, int n_new) {
        Tensor t = w;
        t.shape[0] = n_new;
        t.data.resize((size_t)n_new * (size_t)(w.numel() / w.shape[0]) * 4, 0);
        return t;
    };
    const int D = ivec_dim, Dp = (D + 3) & ~3;
    if (Dp != D && D > 0) {
        const Tensor &M = *find(iv_ie, "M");
        const int64_t G = M.shape[0], F = M.shape[1];
        Tensor t;
        t.dtype = 0;
        t.shape = {G, F, Dp};
        t.data.assign((size_t)(G * F * Dp) * 4, 0);
        for (int64_t r = 0; r < G * F; r++) memcpy(t.data.data() + (size_t)r * Dp * 4, M.data.data() + (size_t)r * D * 4, (size_t)D * 4);
        iv_ie["M"] = std::move(t);
        for (AmOp &op : ops) {
            if (!op.uses_ivec) continue;
            const int Kn = op.K + (Dp - D);  // the i-vector block is the tail of the spliced input
            Tensor w;
            w.dtype = 0;
            w.shape = {op.N, Kn};
            w.data.assign((size_t)op.N * Kn * 4, 0);
            for (int n = 0; n < op.N; n++) memcpy(w.data.data() + (size_t)n * Kn * 4, op.W->data.data() + (size_t)n * op.K * 4, (size_t)op.K * 4);
            op.W = &(am[op.name + ".w(padded)"] = std::move(w));
            op.K = Kn;
        }
        ivec_dim = Dp;
    }
    if (!ops.empty() && ops.back().N % 16) {
        AmOp &op = ops.back();
        const int Nn = (op.N + 15) & ~15;
        op.W = &(am[op.name + ".w(rows padded)"] = padded_rows(*op.W, Nn));
        if (op.b) op.b = &(am[op.name + ".b(padded)"] = padded_rows(*op.b, Nn));
        if (op.bn_s) {
            op.bn_s = &(am[op.name + ".bn_scale(padded)"] = padded_rows(*op.bn_s, Nn));
            op.bn_o = &(am[op.name + ".bn_offset(padded)"] = padded_rows(*op.bn_o, Nn));
        }
        op.N = Nn;
        node_dim.back() = Nn;
    }
    for (const AmOp &op : ops)
        if (op.N % 16 || op.K % 4) throw std::runtime_error("layer " + op.name + ": dimensions must be multiples of 16 (outputs) / 4 (inputs)");
}

void Model::apply_conf(Config *cfg) const {
    auto geti = [&](const std::map<std::string, std::string> &m, const char *k, int *v) {
        auto it = m.find(k);
        if (it != m.end() && !it->second.empty()) *v = std::stoi(it->second);
    };
    auto getf = [&](const std::map<std::string, std::string> &m, const char *k, float *v) {
        auto it = m.find(k);
        if (it != m.end() && !it->second.empty()) *v = std::stof(it->second);
    };
    // The reference's batch path hard-codes its decoding parameters [REF src/batch_model.cc:69-88] and never reads model.conf
    // (only the CPU recognizer does [REF src/model.cc:130-145]); model.conf values and its endpointing rules are opt-in.
    if (cfg->model_conf) {
        getf(conf, "beam", &cfg->beam);
        getf(conf, "lattice-beam", &cfg->lattice_beam);
        geti(conf, "max-active", &cfg->max_active);
        geti(conf, "min-active", &cfg->min_active);
        geti(conf, "frames-per-chunk", &cfg->frames_per_chunk);
        geti(conf, "max-batch-size", &cfg->max_lanes);
        geti(conf, "num-channels", &cfg->num_channels);
        // endpointing [REF src/model.cc:142-145]: Kaldi's OnlineEndpointConfig keys
        auto it = conf.find("endpoint.silence-phones");
        if (it != conf.end()) snprintf(cfg->endpoint_silence_phones, sizeof cfg->endpoint_silence_phones, "%s", it->second.c_str());
        for (int r = 0; r < 4; r++) {
            const std::string pre = "endpoint.rule" + std::to_string(r + 1) + ".";
            auto b = conf.find(pre + "must-contain-nonsilence");
            if (b != conf.end()) cfg->ep_must_contain_nonsilence[r] = b->second == "true" || b->second == "1";
            getf(conf, (pre + "min-trailing-silence").c_str(), &cfg->ep_min_trailing_silence[r]);
            getf(conf, (pre + "min-utterance-length").c_str(), &cfg->ep_min_utterance_length[r]);
            auto m = conf.find(pre + "max-relative-cost");
            if (m != conf.end() && !m->second.empty()) cfg->ep_max_relative_cost[r] = m->second == "inf" ? INFINITY : std::stof(m->second);
        }
        getf(conf, "endpoint.rule5.min-utterance-length", &cfg->endpoint_rule5_seconds);
    }
    // conf/mfcc.conf [REF src/batch_model.cc:75-76]: the mel band edges are configurable, the feature geometry the kernels are
    // built for (40 mel bins, 40 cepstra, no energy, 16 kHz, 25 ms / 10 ms frames, lifter 22) is checked
    {
        auto mf = read_conf(dir + "/conf/mfcc.conf");
        getf(mf, "low-freq", &cfg->mfcc_low_freq);
        getf(mf, "high-freq", &cfg->mfcc_high_freq);
        auto expect = [&](const char *k, double want) {
            auto it = mf.find(k);
            if (it == mf.end() || it->second.empty()) return;
            const double v = it->second == "true" ? 1 : it->second == "false" ? 0 : std::stod(it->second);
            if (std::fabs(v - want) > 1e-9) throw std::runtime_error(std::string("conf/mfcc.conf: --") + k + "=" + it->second + " is not supported (expected " + std::to_string(want) + ")");
        };
        expect("num-mel-bins", 40);
        expect("num-ceps", 40);
        expect("use-energy", 0);
        expect("sample-frequency", 16000);
        expect("frame-length", 25);
        expect("frame-shift", 10);
        expect("cepstral-lifter", 22);
        expect("dither", 0);
        const double hi = cfg->mfcc_high_freq > 0 ? cfg->mfcc_high_freq : 8000.0 + cfg->mfcc_high_freq;
        if (cfg->mfcc_low_freq < 0 || hi <= cfg->mfcc_low_freq || hi > 8000.0) throw std::runtime_error("conf/mfcc.conf: bad --low-freq / --high-freq");
    }
    auto iv = read_conf(dir + "/conf/ivector.conf");
    geti(iv, "num-gselect", &cfg->num_gselect);
    getf(iv, "min-post", &cfg->min_post);
    getf(iv, "posterior-scale", &cfg->posterior_scale);
    getf(iv, "max-count", &cfg->max_count);
    {  // the i-vector front end splices +-3 frames (the LDA matrix has 7 x 40 + 1 columns): splice.conf must say the same
        auto sp = read_conf(dir + "/ivector/splice.conf");
        int l = 3, r = 3;
        geti(sp, "left-context", &l);
        geti(sp, "right-context", &r);
        if (l != 3 || r != 3) throw std::runtime_error("ivector/splice.conf: only --left-context=3 --right-context=3 is supported");
    }
    auto cm = read_conf(dir + "/ivector/online_cmvn.conf");
    geti(cm, "cmn-window", &cfg->cmn_window);
    geti(cm, "global-frames", &cfg->global_frames);
}

}  // namespace vb
