// vb_kaldi.cc — Kaldi on-disk formats for the model files of the batch path (host only).
//
// The reference loads these through Kaldi: ReadKaldiObject(final.mdl) = TransitionModel + AmNnetSimple, then
// SetBatchnormTestMode / SetDropoutTestMode / CollapseModel [REF src/batch_model.cc:39-48]; the i-vector extractor
// files are named by ivector.conf [REF src/batch_model.cc:77], [REF src/model.cc:251-256].  Kaldi is not part of the
// reference tree, so the layouts are restated from its published I/O conventions:
//   binary marker "\0B"; tokens = ASCII + one space; basic types = size byte + little-endian value; bool 'T'/'F';
//   Matrix "FM "/"DM " rows cols data (or a CompressedMatrix "CM "/"CM2 "/"CM3 "); Vector "FV "/"DV " dim data; SpMatrix "FP "/"DP " rows + packed lower triangle;
//   integer vector = size byte, int32 count, raw data;
//   nnet3 = "<Nnet3>" + config lines up to a blank line + "<NumComponents>" + {"<ComponentName>" name <Type> .. </Type>}.
// Component bodies are read as self-describing token/value sequences, so optional fields of different Kaldi versions
// are tolerated; only the fields the forward pass needs are interpreted.
//
// The nnet3 graph is compiled by symbolic linear folding (the job of CollapseModel + the nnet3 compiler): every node's
// value is either a lazy linear expression over already materialised engine nodes (descriptors Append / Offset / Sum /
// Scale, fixed affines, test-mode batchnorm, identities), or a pending affine op (weights over (source, time offset) blocks
// + optional ReLU, batchnorm and scaled bypass), which becomes an engine op when the next weight-bearing component, a
// bypass or the output node consumes it.
#include "vb_kaldi.h"

#include <cctype>
#include <cmath>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>

namespace vb {

namespace {

std::vector<uint8_t> slurp_file(const std::string &path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::streamsize n = f.tellg();
    f.seekg(0);
    std::vector<uint8_t> buf((size_t)n);
    if (n && !f.read(reinterpret_cast<char *>(buf.data()), n)) throw std::runtime_error("cannot read " + path);
    return buf;
}

struct KVal {
    int kind = 0;  // 0 four raw bytes, 1 eight raw bytes, 2 bool, 3 matrix, 4 vector, 5 integer vector
    uint8_t raw[8] = {};
    bool b = false;
    KaldiMatrix m;
    std::vector<double> v;
    std::vector<int32_t> iv;
    int32_t as_int() const {
        int32_t x;
        memcpy(&x, raw, 4);
        return x;
    }
    double as_real() const {
        if (kind == 1) {
            double d;
            memcpy(&d, raw, 8);
            return d;
        }
        float f;
        memcpy(&f, raw, 4);
        return f;
    }
};

struct KComp {
    std::string type;
    std::map<std::string, std::vector<KVal>> f;
    const KVal &get(const std::string &tok, int kind, const std::string &who) const {
        auto it = f.find(tok);
        if (it == f.end() || it->second.empty() || it->second[0].kind != kind)
            throw std::runtime_error("component " + who + " (" + type + ") lacks " + tok);
        return it->second[0];
    }
    bool has(const std::string &tok) const { return f.count(tok) && !f.at(tok).empty(); }
};

struct KReader {
    std::vector<uint8_t> buf;
    size_t p = 0;
    std::string path;
    explicit KReader(const std::string &pth) : buf(slurp_file(pth)), path(pth) {}
    [[noreturn]] void fail(const std::string &what) const {
        throw std::runtime_error(path + ": " + what + " (byte " + std::to_string(p) + ")");
    }
    void need(size_t n) const {
        if (p + n > buf.size()) fail("truncated file");
    }
    int peek() const { return p < buf.size() ? buf[p] : -1; }
    void binary_header() {
        need(2);
        if (buf[0] != 0 || buf[1] != 'B') fail("not a Kaldi binary file");
        p = 2;
    }
    std::string token() {
        while (p < buf.size() && isspace(buf[p])) p++;
        size_t a = p;
        while (p < buf.size() && !isspace(buf[p])) p++;
        if (p == a) fail("token expected");
        std::string s((const char *)&buf[a], p - a);
        if (p < buf.size()) p++;  // the single separator
        return s;
    }
    void expect(const std::string &t) {
        std::string s = token();
        if (s != t) fail("expected " + t + ", found " + s);
    }
    int32_t i32() {
        need(5);
        if (buf[p] != 4) fail("int32 expected");
        int32_t v;
        memcpy(&v, &buf[p + 1], 4);
        p += 5;
        return v;
    }
    double real() {
        need(1);
        if (buf[p] == 4) {
            need(5);
            float v;
            memcpy(&v, &buf[p + 1], 4);
            p += 5;
            return v;
        }
        if (buf[p] == 8) {
            need(9);
            double v;
            memcpy(&v, &buf[p + 1], 8);
            p += 9;
            return v;
        }
        fail("floating-point value expected");
    }
    // "FM " / "DM " / "FV " / "DV " / "FP " / "DP ": returns the two tag characters
    std::string tag() {
        need(3);
        std::string t((const char *)&buf[p], 2);
        if (buf[p + 2] != ' ') fail("matrix/vector tag expected");
        if (t[0] == 'C') fail("compressed matrices (CM) are not supported");
        if ((t[0] != 'F' && t[0] != 'D') || (t[1] != 'M' && t[1] != 'V' && t[1] != 'P')) fail("unknown object tag " + t);
        p += 3;
        return t;
    }
    void data(bool dbl, size_t n, std::vector<double> *out) {
        const size_t item = dbl ? 8 : 4;
        if (n > (buf.size() - p) / item) fail("truncated matrix/vector data");
        out->resize(n);
        if (dbl) {
            memcpy(out->data(), &buf[p], n * 8);
        } else {
            for (size_t i = 0; i < n; i++) {
                float x;
                memcpy(&x, &buf[p + 4 * i], 4);
                (*out)[i] = x;
            }
        }
        p += n * item;
    }
    // CompressedMatrix ("CM " one byte per element with per-column percentile headers, stored by columns; "CM2 " two bytes per
    // element; "CM3 " one byte per element), as Kaldi's compressed-matrix.cc lays it out: min_value, range (float), rows, cols
    // (int32) raw, then the data.  The arithmetic of the expansion is float, as Kaldi's.
    bool compressed_matrix(KaldiMatrix *m) {
        int format = 0;
        if (buf.size() - p >= 3 && memcmp(&buf[p], "CM ", 3) == 0) format = 1, p += 3;
        else if (buf.size() - p >= 4 && memcmp(&buf[p], "CM2 ", 4) == 0) format = 2, p += 4;
        else if (buf.size() - p >= 4 && memcmp(&buf[p], "CM3 ", 4) == 0) format = 3, p += 4;
        else return false;
        need(16);
        float min_value, range;
        int32_t rows, cols;
        memcpy(&min_value, &buf[p], 4);
        memcpy(&range, &buf[p + 4], 4);
        memcpy(&rows, &buf[p + 8], 4);
        memcpy(&cols, &buf[p + 12], 4);
        p += 16;
        if (rows < 0 || cols < 0) fail("negative matrix size");
        const size_t n = (size_t)rows * cols;
        m->rows = rows;
        m->cols = cols;
        m->v.resize(n);
        const float inc16 = 1.52590218966964e-05f;  // 1 / 65535
        if (format == 1) {
            if ((size_t)cols * 8 > buf.size() - p || n > buf.size() - p - (size_t)cols * 8) fail("truncated compressed matrix");
            const unsigned char *hdr = &buf[p], *bytes = &buf[p + (size_t)cols * 8];
            for (int c = 0; c < cols; c++) {
                uint16_t pc[4];
                memcpy(pc, hdr + (size_t)c * 8, 8);
                const float p0 = min_value + range * inc16 * pc[0], p25 = min_value + range * inc16 * pc[1], p75 = min_value + range * inc16 * pc[2],
                            p100 = min_value + range * inc16 * pc[3];
                for (int r = 0; r < rows; r++) {
                    const unsigned char v = bytes[(size_t)c * rows + r];
                    float x;
                    if (v <= 64) x = p0 + (p25 - p0) * v * (1.0f / 64.0f);
                    else if (v <= 192) x = p25 + (p75 - p25) * (v - 64) * (1.0f / 128.0f);
                    else x = p75 + (p100 - p75) * (v - 192) * (1.0f / 63.0f);
                    m->v[(size_t)r * cols + c] = x;
                }
            }
            p += (size_t)cols * 8 + n;
        } else if (format == 2) {
            if (n > (buf.size() - p) / 2) fail("truncated compressed matrix");
            for (size_t i = 0; i < n; i++) {
                uint16_t v;
                memcpy(&v, &buf[p + 2 * i], 2);
                m->v[i] = min_value + range * inc16 * v;
            }
            p += 2 * n;
        } else {
            if (n > buf.size() - p) fail("truncated compressed matrix");
            for (size_t i = 0; i < n; i++) m->v[i] = min_value + range * (1.0f / 255.0f) * buf[p + i];
            p += n;
        }
        return true;
    }
    void matrix(KaldiMatrix *m) {
        if (compressed_matrix(m)) return;
        std::string t = tag();
        if (t[1] != 'M') fail("matrix expected");
        m->rows = i32();
        m->cols = i32();
        if (m->rows < 0 || m->cols < 0) fail("negative matrix size");
        data(t[0] == 'D', (size_t)m->rows * m->cols, &m->v);
    }
    void vector(std::vector<double> *v) {
        std::string t = tag();
        if (t[1] != 'V') fail("vector expected");
        int32_t n = i32();
        if (n < 0) fail("negative vector size");
        data(t[0] == 'D', (size_t)n, v);
    }
    // symmetric packed matrix -> full [n][n]
    void packed(int *n_out, std::vector<double> *full) {
        std::string t = tag();
        if (t[1] != 'P') fail("packed matrix expected");
        int32_t n = i32();
        if (n < 0) fail("negative packed size");
        std::vector<double> pk;
        data(t[0] == 'D', (size_t)n * (n + 1) / 2, &pk);
        full->assign((size_t)n * n, 0.0);
        size_t k = 0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++, k++) (*full)[(size_t)i * n + j] = (*full)[(size_t)j * n + i] = pk[k];
        *n_out = n;
    }
    std::vector<int32_t> intvec() {
        need(5);
        if (buf[p] != 4) fail("int32 vector expected");
        int32_t n;
        memcpy(&n, &buf[p + 1], 4);
        p += 5;
        if (n < 0 || (size_t)n > (buf.size() - p) / 4) fail("truncated integer vector");
        std::vector<int32_t> v((size_t)n);
        if (n) memcpy(v.data(), &buf[p], (size_t)n * 4);
        p += (size_t)n * 4;
        return v;
    }
    // one component: "<Type>" {token values...} "</Type>"
    KComp component() {
        static const std::set<std::string> int_vectors = {"<TimeOffsets>"};
        const std::string open = token();
        if (open.size() < 3 || open[0] != '<' || open.back() != '>') fail("component type expected, found " + open);
        KComp c;
        c.type = open.substr(1, open.size() - 2);
        const std::string close = "</" + c.type + ">";
        for (;;) {
            std::string tok = token();
            if (tok == close) break;
            if (tok == open) continue;  // (older files repeat the opening tag)
            if (tok.empty() || tok[0] != '<') fail("token expected inside " + c.type + ", found " + tok);
            std::vector<KVal> &vals = c.f[tok];
            while (peek() != '<') {
                KVal v;
                const int ch = peek();
                if (ch == 4 && int_vectors.count(tok)) {
                    v.kind = 5;
                    v.iv = intvec();
                } else if (ch == 4) {
                    need(5);
                    v.kind = 0;
                    memcpy(v.raw, &buf[p + 1], 4);
                    p += 5;
                } else if (ch == 8) {
                    need(9);
                    v.kind = 1;
                    memcpy(v.raw, &buf[p + 1], 8);
                    p += 9;
                } else if ((ch == 'F' || ch == 'D' || ch == 'C') && p + 2 < buf.size() && buf[p + 2] == ' ' &&
                           (buf[p + 1] == 'M' || buf[p + 1] == 'V' || buf[p + 1] == 'P' || buf[p + 1] == '2' || buf[p + 1] == '3')) {
                    if (buf[p + 1] == 'M' || ch == 'C') {
                        v.kind = 3;
                        matrix(&v.m);
                    } else if (buf[p + 1] == 'V') {
                        v.kind = 4;
                        vector(&v.v);
                    } else {
                        int n;
                        v.kind = 3;
                        packed(&n, &v.m.v);
                        v.m.rows = v.m.cols = n;
                    }
                } else if (ch == 'T' || ch == 'F') {
                    v.kind = 2;
                    v.b = ch == 'T';
                    p++;
                } else {
                    fail("unparseable value after " + tok + " in " + c.type);
                }
                vals.push_back(std::move(v));
            }
        }
        return c;
    }
};

}  // namespace

bool kaldi_is_binary(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    char h[2] = {1, 1};
    f.read(h, 2);
    return f.gcount() == 2 && h[0] == 0 && h[1] == 'B';
}

bool file_is_vbt(const std::string &path) {
    std::ifstream f(path, std::ios::binary);
    char h[4] = {};
    f.read(h, 4);
    return f.gcount() == 4 && memcmp(h, "VBT1", 4) == 0;
}

KaldiMatrix read_kaldi_matrix_file(const std::string &path) {
    KaldiMatrix m;
    if (kaldi_is_binary(path)) {
        KReader r(path);
        r.binary_header();
        r.matrix(&m);
        return m;
    }
    // text: " [ a b c\n d e f ]"
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open " + path);
    std::string all((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    size_t lb = all.find('['), rb = all.rfind(']');
    if (lb == std::string::npos || rb == std::string::npos || rb < lb) throw std::runtime_error(path + ": not a Kaldi matrix");
    std::istringstream body(all.substr(lb + 1, rb - lb - 1));
    std::string line;
    while (std::getline(body, line)) {
        std::istringstream ls(line);
        std::vector<double> row;
        std::string w;
        while (ls >> w) {
            try {
                size_t used = 0;
                row.push_back(std::stod(w, &used));
                if (used != w.size()) throw std::invalid_argument(w);
            } catch (const std::exception &) {
                throw std::runtime_error(path + ": bad number '" + w + "'");
            }
        }
        if (row.empty()) continue;
        if (m.rows && (int)row.size() != m.cols) throw std::runtime_error(path + ": ragged matrix rows");
        m.cols = (int)row.size();
        m.rows++;
        m.v.insert(m.v.end(), row.begin(), row.end());
    }
    return m;
}

static Tensor f32_tensor(const std::vector<double> &v, std::vector<int64_t> shape) {
    Tensor t;
    t.dtype = 0;
    t.shape = std::move(shape);
    t.data.resize(v.size() * 4);
    float *o = reinterpret_cast<float *>(t.data.data());
    for (size_t i = 0; i < v.size(); i++) o[i] = (float)v[i];
    return t;
}

TensorMap read_kaldi_dubm(const std::string &path) {
    KReader r(path);
    r.binary_header();
    r.expect("<DiagGMM>");
    std::vector<double> gconsts, weights;
    KaldiMatrix miv, iv;
    for (;;) {
        std::string tok = r.token();
        if (tok == "</DiagGMM>") break;
        if (tok == "<GCONSTS>") r.vector(&gconsts);
        else if (tok == "<WEIGHTS>") r.vector(&weights);
        else if (tok == "<MEANS_INVVARS>") r.matrix(&miv);
        else if (tok == "<INV_VARS>") r.matrix(&iv);
        else r.fail("unexpected token " + tok + " in DiagGMM");
    }
    const int G = miv.rows, F = miv.cols;
    if (G <= 0 || iv.rows != G || iv.cols != F || (int)weights.size() != G) r.fail("inconsistent DiagGMM sizes");
    // gconst_g = log w_g - 0.5 * (F log 2pi - sum log inv_var + sum mean^2 * inv_var)   (DiagGmm::ComputeGconsts)
    std::vector<double> gc((size_t)G);
    for (int g = 0; g < G; g++) {
        double s = F * std::log(2.0 * M_PI);
        for (int d = 0; d < F; d++) {
            const double ivd = iv.v[(size_t)g * F + d], m_iv = miv.v[(size_t)g * F + d];
            s += -std::log(ivd) + m_iv * m_iv / ivd;
        }
        gc[g] = std::log(weights[g]) - 0.5 * s;
    }
    if ((int)gconsts.size() != G) gconsts = gc;  // (Kaldi recomputes them on read in any case)
    TensorMap out;
    out["gconsts"] = f32_tensor(gconsts, {G});
    out["weights"] = f32_tensor(weights, {G});
    out["means_invvars"] = f32_tensor(miv.v, {G, F});
    out["inv_vars"] = f32_tensor(iv.v, {G, F});
    return out;
}

TensorMap read_kaldi_ie(const std::string &path) {
    KReader r(path);
    r.binary_header();
    r.expect("<IvectorExtractor>");
    r.expect("<w>");
    KaldiMatrix w;
    r.matrix(&w);
    r.expect("<w_vec>");
    std::vector<double> w_vec;
    r.vector(&w_vec);
    r.expect("<M>");
    const int G = r.i32();
    if (G <= 0) r.fail("empty i-vector extractor");
    std::vector<double> M, S;
    int F = 0, D = 0;
    for (int g = 0; g < G; g++) {
        KaldiMatrix m;
        r.matrix(&m);
        if (g == 0) {
            F = m.rows;
            D = m.cols;
        } else if (m.rows != F || m.cols != D) {
            r.fail("M matrices of different sizes");
        }
        M.insert(M.end(), m.v.begin(), m.v.end());
    }
    r.expect("<SigmaInv>");
    for (int g = 0; g < G; g++) {
        int n;
        std::vector<double> full;
        r.packed(&n, &full);
        if (n != F) r.fail("SigmaInv size mismatch");
        S.insert(S.end(), full.begin(), full.end());
    }
    r.expect("<IvectorOffset>");
    const double off = r.real();
    r.expect("</IvectorExtractor>");
    if (w.rows != 0) r.fail("i-vector extractors with weight projections (<w>) are not supported");
    TensorMap out;
    out["M"] = f32_tensor(M, {G, F, D});
    out["sigma_inv"] = f32_tensor(S, {G, F, F});
    out["w"] = f32_tensor(w_vec, {(int64_t)w_vec.size()});
    out["prior_offset"] = f32_tensor({off}, {1});
    return out;
}

// ---------------------------------------------------------------------------------------------------------------
// nnet3 graph -> collapsed op chain
// ---------------------------------------------------------------------------------------------------------------
namespace {

constexpr int kIvecSrc = -2;
enum { IDENT = 0, DIAG = 1, DENSE = 2 };

struct Term {
    int src = 0, off = 0, row0 = 0, nrows = 0, kind = IDENT;
    double scale = 1.0;
    std::vector<double> m;  // DIAG: [nrows]; DENSE: [nrows][srcdim]
};
struct LinExpr {
    int dim = 0;
    std::vector<double> bias;
    std::vector<Term> terms;
};
using WMap = std::map<std::pair<int, int>, std::vector<double>>;  // (src, off) -> [N][srcdim]
struct Pending {
    std::string name;
    int N = 0;
    std::shared_ptr<const WMap> w;
    std::vector<double> bias;
    bool has_bias = false, relu = false, bn = false;
    std::vector<double> bn_s, bn_o;
    int byp = -1;
    double byp_scale = 0;
    std::shared_ptr<int> mat_id = std::make_shared<int>(-1);  // shared by the aliases of this op
};
struct Value {
    int kind = 0;  // 0 lazy, 1 pending, 2 materialised node
    LinExpr lazy;
    std::shared_ptr<Pending> pend;
    int mat = -1;
};
struct DV {  // descriptor value with a not-yet-applied scale
    Value v;
    double scale = 1.0;
};

struct ConfLine {
    std::string type;
    std::map<std::string, std::string> kv;
};

ConfLine parse_conf_line(const std::string &line) {
    ConfLine c;
    size_t sp = line.find(' ');
    c.type = line.substr(0, sp);
    if (sp == std::string::npos) return c;
    // keys are identifiers followed by '=' at parenthesis depth 0, preceded by a space
    std::vector<std::pair<size_t, size_t>> keys;  // (key start, '=' position)
    int depth = 0;
    for (size_t i = sp; i < line.size(); i++) {
        const char ch = line[i];
        if (ch == '(') depth++;
        else if (ch == ')') depth--;
        else if (ch == '=' && depth == 0) {
            size_t a = i;
            while (a > 0 && (isalnum((unsigned char)line[a - 1]) || line[a - 1] == '-' || line[a - 1] == '_')) a--;
            if (a < i && a > 0 && line[a - 1] == ' ') keys.push_back({a, i});
        }
    }
    for (size_t k = 0; k < keys.size(); k++) {
        size_t vend = k + 1 < keys.size() ? keys[k + 1].first : line.size();
        std::string val = line.substr(keys[k].second + 1, vend - keys[k].second - 1);
        size_t a = val.find_first_not_of(" \t"), b = val.find_last_not_of(" \t\r");
        val = a == std::string::npos ? "" : val.substr(a, b - a + 1);
        c.kv[line.substr(keys[k].first, keys[k].second - keys[k].first)] = val;
    }
    return c;
}

struct Compiler {
    std::string path;
    std::map<std::string, ConfLine> nodes;     // node name -> its config line
    std::map<std::string, KComp> comps;
    std::map<std::string, Value> memo;
    std::vector<int> node_dim;                  // engine nodes
    int ivec_dim = 0;
    std::vector<KaldiOp> ops;

    [[noreturn]] void fail(const std::string &what) const { throw std::runtime_error(path + ": " + what); }
    int srcdim(int src) const { return src == kIvecSrc ? ivec_dim : node_dim[src]; }

    static LinExpr ref(int src, int dim) {
        LinExpr e;
        e.dim = dim;
        e.bias.assign(dim, 0.0);
        Term t;
        t.src = src;
        t.nrows = dim;
        e.terms.push_back(t);
        return e;
    }
    std::vector<double> dense(const Term &t) const {
        const int sd = srcdim(t.src);
        if (t.kind == DENSE) return t.m;
        std::vector<double> d((size_t)t.nrows * sd, 0.0);
        for (int i = 0; i < t.nrows; i++) d[(size_t)i * sd + i] = t.kind == IDENT ? t.scale : t.m[i];
        return d;
    }
    static void scale_expr(LinExpr *e, double s) {
        if (s == 1.0) return;
        for (double &b : e->bias) b *= s;
        for (Term &t : e->terms) {
            if (t.kind == IDENT) t.scale *= s;
            else
                for (double &x : t.m) x *= s;
        }
    }
    void rowscale(LinExpr *e, const std::vector<double> &s, const std::vector<double> &o) const {
        for (Term &t : e->terms) {
            const int sd = srcdim(t.src);
            if (t.kind == IDENT) {
                t.m.resize(t.nrows);
                for (int i = 0; i < t.nrows; i++) t.m[i] = t.scale * s[t.row0 + i];
                t.kind = DIAG;
            } else if (t.kind == DIAG) {
                for (int i = 0; i < t.nrows; i++) t.m[i] *= s[t.row0 + i];
            } else {
                for (int i = 0; i < t.nrows; i++)
                    for (int j = 0; j < sd; j++) t.m[(size_t)i * sd + j] *= s[t.row0 + i];
            }
        }
        for (int i = 0; i < e->dim; i++) e->bias[i] = e->bias[i] * s[i] + o[i];
    }
    static LinExpr append(const std::vector<LinExpr> &parts) {
        LinExpr e;
        for (const LinExpr &p : parts) {
            for (Term t : p.terms) {
                t.row0 += e.dim;
                e.terms.push_back(std::move(t));
            }
            e.bias.insert(e.bias.end(), p.bias.begin(), p.bias.end());
            e.dim += p.dim;
        }
        return e;
    }
    // rows [a, a+n) of an expression (dim-range-node)
    LinExpr range(const LinExpr &x, int a, int n) const {
        if (a < 0 || n <= 0 || a + n > x.dim) fail("dim-range outside its input");
        LinExpr e;
        e.dim = n;
        e.bias.assign(x.bias.begin() + a, x.bias.begin() + a + n);
        for (const Term &t : x.terms) {
            const int lo = std::max(a, t.row0), hi = std::min(a + n, t.row0 + t.nrows);
            if (lo >= hi) continue;
            const int sd = srcdim(t.src);
            std::vector<double> d = dense(t);
            Term u;
            u.src = t.src;
            u.off = t.off;
            u.row0 = lo - a;
            u.nrows = hi - lo;
            u.kind = DENSE;
            u.m.assign(d.begin() + (size_t)(lo - t.row0) * sd, d.begin() + (size_t)(hi - t.row0) * sd);
            e.terms.push_back(std::move(u));
        }
        return e;
    }
    // y = W x + b for a lazy x, as a lazy expression (fixed affines)
    LinExpr lazy_affine(const KaldiMatrix &W, const std::vector<double> &b, const LinExpr &x) const {
        if (W.cols != x.dim) fail("affine input dimension mismatch");
        LinExpr e;
        e.dim = W.rows;
        e.bias.assign(W.rows, 0.0);
        for (int n = 0; n < W.rows; n++) {
            double s = b.empty() ? 0.0 : b[n];
            for (int k = 0; k < W.cols; k++) s += W.v[(size_t)n * W.cols + k] * x.bias[k];
            e.bias[n] = s;
        }
        for (const Term &t : x.terms) {
            const int sd = srcdim(t.src);
            Term u;
            u.src = t.src;
            u.off = t.off;
            u.row0 = 0;
            u.nrows = W.rows;
            u.kind = DENSE;
            u.m.assign((size_t)W.rows * sd, 0.0);
            add_block(W, t, &u.m);
            e.terms.push_back(std::move(u));
        }
        return e;
    }
    // acc[N][srcdim] += W[:, row0 : row0 + nrows] * T
    void add_block(const KaldiMatrix &W, const Term &t, std::vector<double> *acc) const {
        const int sd = srcdim(t.src), N = W.rows, K = W.cols;
        if (t.kind == IDENT || t.kind == DIAG) {
            for (int n = 0; n < N; n++) {
                const double *wr = &W.v[(size_t)n * K + t.row0];
                double *ar = &(*acc)[(size_t)n * sd];
                for (int i = 0; i < t.nrows; i++) ar[i] += wr[i] * (t.kind == IDENT ? t.scale : t.m[i]);
            }
        } else {
            for (int n = 0; n < N; n++) {
                const double *wr = &W.v[(size_t)n * K + t.row0];
                double *ar = &(*acc)[(size_t)n * sd];
                for (int i = 0; i < t.nrows; i++) {
                    const double wv = wr[i];
                    if (wv == 0.0) continue;
                    const double *mr = &t.m[(size_t)i * sd];
                    for (int j = 0; j < sd; j++) ar[j] += wv * mr[j];
                }
            }
        }
    }

    // ---- values ----
    static void normalise(Value *v) {
        if (v->kind == 1 && *v->pend->mat_id >= 0) {
            v->mat = *v->pend->mat_id;
            v->kind = 2;
            v->pend.reset();
        }
    }
    int materialise(Value *v) {
        normalise(v);
        if (v->kind == 2) return v->mat;
        if (v->kind != 1) fail("internal: cannot materialise a lazy expression");
        const Pending &p = *v->pend;
        KaldiOp op;
        op.name = p.name;
        op.N = p.N;
        std::set<int> srcs, offs;
        bool iv = false;
        for (const auto &kv : *p.w) {
            if (kv.first.first == kIvecSrc) iv = true;
            else {
                srcs.insert(kv.first.first);
                offs.insert(kv.first.second);
            }
        }
        if (srcs.size() != 1) fail("component " + p.name + " reads " + std::to_string(srcs.size()) + " source layers (one is supported)");
        op.in_node = *srcs.begin();
        op.offs.assign(offs.begin(), offs.end());
        op.uses_ivec = iv;
        const int sd = node_dim[op.in_node];
        op.K = sd * (int)op.offs.size() + (iv ? ivec_dim : 0);
        op.W.assign((size_t)op.N * op.K, 0.f);
        for (const auto &kv : *p.w) {
            int col0, width;
            if (kv.first.first == kIvecSrc) {
                col0 = sd * (int)op.offs.size();
                width = ivec_dim;
            } else {
                col0 = sd * (int)(std::lower_bound(op.offs.begin(), op.offs.end(), kv.first.second) - op.offs.begin());
                width = sd;
            }
            for (int n = 0; n < op.N; n++)
                for (int j = 0; j < width; j++) op.W[(size_t)n * op.K + col0 + j] += (float)kv.second[(size_t)n * width + j];
        }
        bool any_bias = false;
        for (double b : p.bias) any_bias |= b != 0.0;
        if (any_bias) {
            op.b.resize(op.N);
            for (int n = 0; n < op.N; n++) op.b[n] = (float)p.bias[n];
        }
        op.relu = p.relu;
        if (p.relu) {
            op.bn_s.assign(op.N, 1.f);
            op.bn_o.assign(op.N, 0.f);
            if (p.bn)
                for (int n = 0; n < op.N; n++) {
                    op.bn_s[n] = (float)p.bn_s[n];
                    op.bn_o[n] = (float)p.bn_o[n];
                }
        }
        op.byp_node = p.byp;
        op.byp_scale = (float)p.byp_scale;
        ops.push_back(std::move(op));
        node_dim.push_back(p.N);
        const int id = (int)node_dim.size() - 1;
        *p.mat_id = id;
        normalise(v);
        return id;
    }
    LinExpr as_lazy(DV d) {
        normalise(&d.v);
        LinExpr e;
        if (d.v.kind == 0) e = d.v.lazy;
        else {
            const int id = materialise(&d.v);
            e = ref(id, node_dim[id]);
        }
        scale_expr(&e, d.scale);
        return e;
    }
    // engine node id of a value that is (or can become) exactly one materialised layer, else -1
    int as_node(DV d) {
        normalise(&d.v);
        if (d.scale != 1.0) return -1;
        if (d.v.kind == 0) {
            const LinExpr &e = d.v.lazy;
            if (e.terms.size() != 1) return -1;
            const Term &t = e.terms[0];
            for (double b : e.bias)
                if (b != 0.0) return -1;
            if (t.kind != IDENT || t.scale != 1.0 || t.off != 0 || t.src < 0 || t.row0 != 0 || t.nrows != e.dim) return -1;
            return t.src;
        }
        return materialise(&d.v);
    }
    int dim_of(const Value &v) const { return v.kind == 0 ? v.lazy.dim : v.kind == 1 ? v.pend->N : node_dim[v.mat]; }

    // ---- descriptors ----
    static std::vector<std::string> lex(const std::string &s) {
        std::vector<std::string> out;
        size_t i = 0;
        while (i < s.size()) {
            const char ch = s[i];
            if (isspace((unsigned char)ch)) {
                i++;
            } else if (ch == '(' || ch == ')' || ch == ',') {
                out.emplace_back(1, ch);
                i++;
            } else {
                size_t a = i;
                while (i < s.size() && !isspace((unsigned char)s[i]) && s[i] != '(' && s[i] != ')' && s[i] != ',') i++;
                out.push_back(s.substr(a, i - a));
            }
        }
        return out;
    }
    DV descriptor(const std::string &text) {
        std::vector<std::string> tk = lex(text);
        size_t pos = 0;
        DV d = desc(tk, &pos);
        if (pos != tk.size()) fail("trailing text in descriptor '" + text + "'");
        return d;
    }
    const std::string &tok_at(const std::vector<std::string> &tk, size_t pos) const {
        if (pos >= tk.size()) fail("descriptor ends early");
        return tk[pos];
    }
    void eat(const std::vector<std::string> &tk, size_t *pos, const char *what) const {
        if (tok_at(tk, *pos) != what) fail(std::string("descriptor: expected '") + what + "', found '" + tk[*pos] + "'");
        ++*pos;
    }
    DV desc(const std::vector<std::string> &tk, size_t *pos) {
        const std::string head = tok_at(tk, *pos);
        ++*pos;
        if (*pos >= tk.size() || tk[*pos] != "(") {  // a node name
            DV d;
            d.v = eval_node(head);
            return d;
        }
        ++*pos;  // '('
        DV out;
        if (head == "Append" || head == "Sum") {
            std::vector<DV> args;
            for (;;) {
                args.push_back(desc(tk, pos));
                if (tok_at(tk, *pos) == ",") {
                    ++*pos;
                    continue;
                }
                break;
            }
            if (head == "Append") {
                std::vector<LinExpr> parts;
                for (DV &a : args) parts.push_back(as_lazy(a));
                out.v.lazy = append(parts);
            } else {
                out = sum(args);
            }
        } else if (head == "Offset") {
            DV a = desc(tk, pos);
            eat(tk, pos, ",");
            const int n = std::stoi(tok_at(tk, *pos));
            ++*pos;
            if (tok_at(tk, *pos) == ",") {  // optional x offset
                ++*pos;
                if (std::stoi(tok_at(tk, *pos)) != 0) fail("Offset with an x offset is not supported");
                ++*pos;
            }
            LinExpr e = as_lazy(a);
            for (Term &t : e.terms)
                if (t.src != kIvecSrc) t.off += n;
            out.v.lazy = std::move(e);
        } else if (head == "Scale") {
            const double s = std::stod(tok_at(tk, *pos));
            ++*pos;
            eat(tk, pos, ",");
            out = desc(tk, pos);
            out.scale *= s;
        } else if (head == "ReplaceIndex" || head == "Round") {
            out = desc(tk, pos);
            while (tok_at(tk, *pos) == ",") *pos += 2;  // (", t, 0" / ", 10": which frame's i-vector a chunk uses is the engine's business)
            LinExpr e = as_lazy(out);
            for (const Term &t : e.terms)
                if (t.src != kIvecSrc) fail(head + " is supported on the ivector input only");
            out = DV();
            out.v.lazy = std::move(e);
        } else if (head == "IfDefined") {
            out = desc(tk, pos);
        } else if (head == "Failover") {
            out = desc(tk, pos);
            eat(tk, pos, ",");
            (void)desc(tk, pos);
        } else {
            fail("unsupported descriptor function " + head);
        }
        eat(tk, pos, ")");
        return out;
    }
    DV sum(std::vector<DV> &args) {
        for (DV &a : args) normalise(&a.v);
        if (args.size() == 2) {  // out = f(affine) + c * other  ->  the op's scaled bypass
            for (int k = 0; k < 2; k++) {
                DV &y = args[k], &x = args[1 - k];
                if (y.v.kind == 1 && y.scale == 1.0 && y.v.pend->relu && y.v.pend->byp < 0) {
                    DV xs = x;
                    const double c = xs.scale;
                    xs.scale = 1.0;
                    if (dim_of(xs.v) != y.v.pend->N) fail("Sum of different dimensions");
                    const int node = as_node(xs);
                    if (node < 0) break;
                    auto p = std::make_shared<Pending>(*y.v.pend);
                    p->mat_id = std::make_shared<int>(-1);
                    p->byp = node;
                    p->byp_scale = c;
                    DV out;
                    out.v.kind = 1;
                    out.v.pend = std::move(p);
                    return out;
                }
            }
        }
        LinExpr e = as_lazy(args[0]);
        for (size_t i = 1; i < args.size(); i++) {
            LinExpr b = as_lazy(args[i]);
            if (b.dim != e.dim) fail("Sum of different dimensions");
            for (int j = 0; j < e.dim; j++) e.bias[j] += b.bias[j];
            for (Term &t : b.terms) e.terms.push_back(std::move(t));
        }
        DV out;
        out.v.lazy = std::move(e);
        return out;
    }

    // ---- components ----
    DV weighty(const std::string &name, const KaldiMatrix &W, const std::vector<double> &b, const std::vector<int32_t> &time_offsets, DV in) {
        LinExpr x = as_lazy(in);
        LinExpr sp;
        if (time_offsets.empty() || (time_offsets.size() == 1 && time_offsets[0] == 0)) {
            sp = std::move(x);
        } else {
            std::vector<LinExpr> parts;
            for (int32_t o : time_offsets) {
                LinExpr c = x;
                for (Term &t : c.terms)
                    if (t.src != kIvecSrc) t.off += o;
                parts.push_back(std::move(c));
            }
            sp = append(parts);
        }
        if (W.cols != sp.dim) fail("component " + name + ": weight has " + std::to_string(W.cols) + " columns, input has dimension " + std::to_string(sp.dim));
        if (!b.empty() && (int)b.size() != W.rows) fail("component " + name + ": bias size mismatch");
        auto p = std::make_shared<Pending>();
        p->name = name;
        p->N = W.rows;
        auto w = std::make_shared<WMap>();
        for (const Term &t : sp.terms) {
            const int off = t.src == kIvecSrc ? 0 : t.off;
            std::vector<double> &acc = (*w)[{t.src, off}];
            if (acc.empty()) acc.assign((size_t)W.rows * srcdim(t.src), 0.0);
            add_block(W, t, &acc);
        }
        p->w = w;
        p->bias.assign(W.rows, 0.0);
        for (int n = 0; n < W.rows; n++) {
            double s = b.empty() ? 0.0 : b[n];
            for (int k = 0; k < W.cols; k++) s += W.v[(size_t)n * W.cols + k] * sp.bias[k];
            p->bias[n] = s;
        }
        DV out;
        out.v.kind = 1;
        out.v.pend = std::move(p);
        return out;
    }
    DV batchnorm(const KComp &c, const std::string &name, DV in) {
        const int dim = c.get("<Dim>", 0, name).as_int();
        const int block = c.has("<BlockDim>") ? c.get("<BlockDim>", 0, name).as_int() : dim;
        const double eps = c.get("<Epsilon>", 0, name).as_real(), target = c.get("<TargetRms>", 0, name).as_real();
        const double count = c.has("<Count>") ? c.f.at("<Count>")[0].as_real() : 1.0;
        const std::vector<double> &mean = c.get("<StatsMean>", 4, name).v, &var = c.get("<StatsVar>", 4, name).v;
        if (block <= 0 || dim % block || (count != 0.0 && ((int)mean.size() != block || (int)var.size() != block))) fail("batchnorm " + name + ": inconsistent sizes");
        std::vector<double> s(dim, 1.0), o(dim, 0.0);
        if (count != 0.0)
            for (int i = 0; i < dim; i++) {
                s[i] = target / std::sqrt(var[i % block] + eps);
                o[i] = -mean[i % block] * s[i];
            }
        normalise(&in.v);
        if (dim_of(in.v) != dim) fail("batchnorm " + name + ": input dimension mismatch");
        if (in.v.kind == 1 && in.scale == 1.0 && in.v.pend->byp < 0) {
            auto p = std::make_shared<Pending>(*in.v.pend);
            p->mat_id = std::make_shared<int>(-1);
            if (p->relu && !p->bn) {
                p->bn = true;
                p->bn_s = s;
                p->bn_o = o;
            } else if (p->relu) {
                for (int i = 0; i < dim; i++) {
                    p->bn_o[i] = p->bn_o[i] * s[i] + o[i];
                    p->bn_s[i] *= s[i];
                }
            } else {  // no nonlinearity in between: fold into the rows of the weights
                auto w = std::make_shared<WMap>(*p->w);
                for (auto &kv : *w) {
                    const int sd = srcdim(kv.first.first);
                    for (int n = 0; n < dim; n++)
                        for (int j = 0; j < sd; j++) kv.second[(size_t)n * sd + j] *= s[n];
                }
                p->w = w;
                for (int n = 0; n < dim; n++) p->bias[n] = p->bias[n] * s[n] + o[n];
            }
            DV out;
            out.v.kind = 1;
            out.v.pend = std::move(p);
            return out;
        }
        LinExpr e = as_lazy(in);
        rowscale(&e, s, o);
        DV out;
        out.v.lazy = std::move(e);
        return out;
    }
    DV apply(const KComp &c, const std::string &name, DV in) {
        static const std::set<std::string> identity = {"NoOpComponent", "DropoutComponent", "GeneralDropoutComponent",
                                                        "SpecAugmentTimeMaskComponent"};
        static const std::vector<int32_t> none;
        static const std::vector<double> nobias;
        const std::string &t = c.type;
        if (identity.count(t)) return in;  // (test mode [REF src/batch_model.cc:46-47])
        if (t == "FixedAffineComponent") {
            DV out;
            out.v.lazy = lazy_affine(c.get("<LinearParams>", 3, name).m, c.get("<BiasParams>", 4, name).v, as_lazy(in));
            return out;
        }
        if (t == "AffineComponent" || t == "NaturalGradientAffineComponent")
            return weighty(name, c.get("<LinearParams>", 3, name).m, c.get("<BiasParams>", 4, name).v, none, in);
        if (t == "LinearComponent") return weighty(name, c.get("<Params>", 3, name).m, nobias, none, in);
        if (t == "TdnnComponent")
            return weighty(name, c.get("<LinearParams>", 3, name).m, c.get("<BiasParams>", 4, name).v, c.get("<TimeOffsets>", 5, name).iv, in);
        if (t == "BatchNormComponent") return batchnorm(c, name, in);
        if (t == "RectifiedLinearComponent") {
            normalise(&in.v);
            if (in.v.kind != 1 || in.scale != 1.0 || in.v.pend->relu || in.v.pend->byp >= 0)
                fail("ReLU " + name + " does not directly follow an affine component (unsupported topology)");
            auto p = std::make_shared<Pending>(*in.v.pend);
            p->mat_id = std::make_shared<int>(-1);
            p->relu = true;
            DV out;
            out.v.kind = 1;
            out.v.pend = std::move(p);
            return out;
        }
        fail("unsupported nnet3 component type " + t + " (" + name + ")");
    }

    Value eval_node(const std::string &name) {
        auto m = memo.find(name);
        if (m != memo.end()) {
            normalise(&m->second);
            return m->second;
        }
        auto it = nodes.find(name);
        if (it == nodes.end()) fail("undefined node " + name);
        const ConfLine &ln = it->second;
        auto kv = [&](const char *k) -> const std::string & {
            auto f = ln.kv.find(k);
            if (f == ln.kv.end()) fail("config line of node " + name + " lacks " + k);
            return f->second;
        };
        Value v;
        if (ln.type == "input-node") {
            if (name == "input") {
                v.kind = 2;
                v.mat = 0;
            } else if (name == "ivector") {
                v.lazy = ref(kIvecSrc, ivec_dim);
            } else {
                fail("unknown input node " + name);
            }
        } else if (ln.type == "component-node") {
            auto c = comps.find(kv("component"));
            if (c == comps.end()) fail("node " + name + " uses undefined component " + kv("component"));
            DV in = descriptor(kv("input"));
            DV out = apply(c->second, name, in);
            if (out.scale != 1.0) {
                LinExpr e = as_lazy(out);
                out = DV();
                out.v.lazy = std::move(e);
            }
            v = out.v;
        } else if (ln.type == "dim-range-node") {
            DV in;
            in.v = eval_node(kv("input-node"));
            v.lazy = range(as_lazy(in), std::stoi(kv("dim-offset")), std::stoi(kv("dim")));
        } else {
            fail("node " + name + " of type " + ln.type + " cannot be an input");
        }
        memo[name] = v;
        return v;
    }
};

}  // namespace

KaldiAm read_kaldi_final_mdl(const std::string &path) {
    KReader r(path);
    r.binary_header();
    KaldiAm am;
    // ---- TransitionModel ----
    r.expect("<TransitionModel>");
    r.expect("<Topology>");
    std::vector<int32_t> phones = r.intvec(), phone2idx = r.intvec();
    int32_t n_entries = r.i32();
    bool is_hmm = true;
    if (n_entries == -1) {  // extended format: every state also carries a self-loop pdf class
        is_hmm = false;
        n_entries = r.i32();
    }
    if (n_entries < 0) r.fail("bad topology");
    struct TopoState {
        std::vector<int32_t> dst;
    };
    std::vector<std::vector<TopoState>> entries((size_t)n_entries);
    for (auto &e : entries) {
        const int32_t ns = r.i32();
        if (ns < 0) r.fail("bad topology entry");
        e.resize((size_t)ns);
        for (auto &st : e) {
            r.i32();  // forward pdf class
            if (!is_hmm) r.i32();
            const int32_t nt = r.i32();
            if (nt < 0) r.fail("bad topology state");
            for (int32_t k = 0; k < nt; k++) {
                st.dst.push_back(r.i32());
                r.real();
            }
        }
    }
    r.expect("</Topology>");
    const std::string tup = r.token();
    if (tup != "<Tuples>" && tup != "<Triples>") r.fail("expected <Tuples> or <Triples>, found " + tup);
    const int32_t n_tuples = r.i32();
    if (n_tuples < 0) r.fail("bad tuple count");
    am.tid2pdf.assign(1, -1);
    am.tid2phone.assign(1, 0);
    am.tid_flags.assign(1, 0);
    int max_pdf = -1;
    for (int32_t i = 0; i < n_tuples; i++) {
        const int32_t phone = r.i32(), hs = r.i32(), fwd = r.i32();
        const int32_t self = tup == "<Tuples>" ? r.i32() : fwd;
        if (phone < 0 || phone >= (int)phone2idx.size() || phone2idx[phone] < 0 || phone2idx[phone] >= n_entries) r.fail("tuple with an unknown phone");
        const auto &e = entries[(size_t)phone2idx[phone]];
        if (hs < 0 || hs >= (int)e.size()) r.fail("tuple with an unknown HMM state");
        for (int32_t d : e[(size_t)hs].dst) {
            am.tid2phone.push_back(phone);
            am.tid2pdf.push_back(d == hs ? self : fwd);
            // TransitionModel::IsSelfLoop / IsFinal (the transition enters the topology's last, non-emitting state) / HMM state 0
            am.tid_flags.push_back((uint8_t)((d == hs ? 1 : 0) | (d + 1 == (int32_t)e.size() ? 2 : 0) | (hs == 0 ? 4 : 0)));
        }
        max_pdf = std::max(max_pdf, std::max(fwd, self));
    }
    r.expect(tup == "<Tuples>" ? "</Tuples>" : "</Triples>");
    r.expect("<LogProbs>");
    std::vector<double> log_probs;
    r.vector(&log_probs);
    r.expect("</LogProbs>");
    r.expect("</TransitionModel>");
    // ---- nnet3 ----
    r.expect("<Nnet3>");
    Compiler cc;
    cc.path = path;
    std::string output_input;
    {
        auto line = [&]() {
            size_t a = r.p;
            while (r.p < r.buf.size() && r.buf[r.p] != '\n') r.p++;
            std::string s((const char *)&r.buf[a], r.p - a);
            if (r.p < r.buf.size()) r.p++;
            while (!s.empty() && (s.back() == '\r' || s.back() == ' ')) s.pop_back();
            return s;
        };
        if (!line().empty()) r.fail("expected a newline after <Nnet3>");
        for (;;) {
            if (r.p >= r.buf.size()) r.fail("unterminated nnet3 config section");
            std::string s = line();
            if (s.empty()) break;
            ConfLine c = parse_conf_line(s);
            auto nm = c.kv.find("name");
            if (nm == c.kv.end()) r.fail("config line without a name: " + s);
            if (c.type == "output-node") {
                if (nm->second == "output") output_input = c.kv.count("input") ? c.kv["input"] : "";
                continue;
            }
            if (c.type == "input-node") {
                const int dim = c.kv.count("dim") ? std::stoi(c.kv["dim"]) : 0;
                if (nm->second == "input") am.feat_dim = dim;
                if (nm->second == "ivector") am.ivec_dim = dim;
            }
            cc.nodes[nm->second] = c;
        }
    }
    r.expect("<NumComponents>");
    const int32_t ncomp = r.i32();
    if (ncomp < 0) r.fail("bad component count");
    for (int32_t i = 0; i < ncomp; i++) {
        r.expect("<ComponentName>");
        const std::string name = r.token();
        cc.comps[name] = r.component();
    }
    r.expect("</Nnet3>");
    std::vector<double> priors;
    if (r.p < r.buf.size()) {  // AmNnetSimple trailer
        r.expect("<LeftContext>");
        r.i32();
        r.expect("<RightContext>");
        r.i32();
        r.expect("<Priors>");
        r.vector(&priors);
    }
    if (output_input.empty()) r.fail("no output-node named 'output'");
    if (am.feat_dim <= 0) r.fail("no input-node named 'input'");
    cc.ivec_dim = am.ivec_dim;
    cc.node_dim.push_back(am.feat_dim);
    DV out = cc.descriptor(output_input);
    Compiler::normalise(&out.v);
    if (out.scale != 1.0 || out.v.kind == 0) r.fail("the output node must read an affine component directly");
    cc.materialise(&out.v);
    am.ops = std::move(cc.ops);
    if (am.ops.empty()) r.fail("empty network");
    am.num_pdfs = am.ops.back().N;
    if (max_pdf >= am.num_pdfs) r.fail("transition model refers to pdf " + std::to_string(max_pdf) + " but the network has " + std::to_string(am.num_pdfs) + " outputs");
    if (!priors.empty()) {  // DecodableAmNnetSimple subtracts the log priors from the output
        if ((int)priors.size() != am.num_pdfs) r.fail("priors size mismatch");
        KaldiOp &last = am.ops.back();
        if (last.relu) r.fail("priors on a nonlinear output are not supported");
        if (last.b.empty()) last.b.assign(last.N, 0.f);
        for (int n = 0; n < last.N; n++) last.b[n] -= (float)std::log(priors[n]);
    }
    // model context: first / last input frame the output at t = 0 depends on
    const int nn = (int)am.ops.size() + 1;
    std::vector<int> t_lo(nn, 1 << 30), t_hi(nn, -(1 << 30));
    t_lo[nn - 1] = t_hi[nn - 1] = 0;
    for (int o = nn - 2; o >= 0; o--) {
        const KaldiOp &op = am.ops[o];
        if (t_lo[o + 1] > t_hi[o + 1]) continue;
        const int mn = op.offs.empty() ? 0 : op.offs.front(), mx = op.offs.empty() ? 0 : op.offs.back();
        t_lo[op.in_node] = std::min(t_lo[op.in_node], t_lo[o + 1] + mn);
        t_hi[op.in_node] = std::max(t_hi[op.in_node], t_hi[o + 1] + mx);
        if (op.byp_node >= 0) {
            t_lo[op.byp_node] = std::min(t_lo[op.byp_node], t_lo[o + 1]);
            t_hi[op.byp_node] = std::max(t_hi[op.byp_node], t_hi[o + 1]);
        }
    }
    am.left_context = -t_lo[0];
    am.right_context = t_hi[0];
    return am;
}

}  // namespace vb
