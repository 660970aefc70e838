// spada_host.hpp -- compiled host side above the C ABI: the reference's own interface for the hot
// path, restated in C++ because the image has no Rust toolchain (the Rust -sys crate that a
// maintainer would use instead is in ../rust/).  Same names, argument meaning and error behaviour
// as the reference:
//   frontend.rs:8-85    OmegaConfig, Cli (positional grammar, case-insensitive enums, -p), parse_config
//   py2rust.rs:62-97    load_mm_mat        (native Matrix Market reader instead of embedded scipy)
//   py2rust.rs:5-60     load_pickled_gemms (python3 helper process, the pickle is Python's format)
//   gemm.rs:26-91       GEMM, GEMM::new / from_mat (square => A x A, else A x A^T), Display
//   storage.rs:22-324   Element, CsrRow (+ Display), CsrMatStorage::init_with_gemm
//   simulator.rs:431-1062  Simulator::new / execute / get_exec_result / get_*_stat
//   preprocessing.rs:76-89 sort_by_length
// All arithmetic of C = A x B happens behind include/spada_b200.h on the GPU; nothing here computes
// products (the transpose and the loaders only move values).
#pragma once
#include <algorithm>
#include <array>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include <unistd.h>

#include "../../include/spada_b200.h"

namespace spada_host {

// ---- Rust `{:?}` formatting of the values the reference prints --------------------------------
inline std::string debug_f64(double x) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
    if (x == 0.0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);  // shortest round-trip
    std::string s(buf, res.ptr);
    size_t epos = s.find('e');
    std::string mant = s.substr(0, epos);
    int exp = std::atoi(s.c_str() + epos + 1);
    bool neg = mant[0] == '-';
    if (neg) mant = mant.substr(1);
    std::string digits;
    for (char c : mant)
        if (c != '.') digits.push_back(c);
    double ax = std::fabs(x);
    std::string out;
    if (ax >= 1e-4 && ax < 1e16) {
        // decimal notation with at least one fractional digit
        if (exp >= 0) {
            std::string ip = digits.substr(0, std::min<size_t>(digits.size(), (size_t)exp + 1));
            while ((int)ip.size() < exp + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)exp + 1 ? digits.substr(exp + 1) : "0";
            out = ip + "." + fp;
        } else {
            out = "0." + std::string((size_t)(-exp - 1), '0') + digits;
        }
    } else {
        out = digits.substr(0, 1);
        if (digits.size() > 1) out += "." + digits.substr(1);
        out += "e" + std::to_string(exp);
    }
    return (neg ? "-" : "") + out;
}
template <typename It>
std::string debug_list_f64(It b, It e) {
    std::string s = "[";
    for (It i = b; i != e; ++i) s += (i == b ? "" : ", ") + debug_f64(*i);
    return s + "]";
}
template <typename It>
std::string debug_list_int(It b, It e) {
    std::string s = "[";
    for (It i = b; i != e; ++i) s += (i == b ? "" : ", ") + std::to_string(*i);
    return s + "]";
}

// ---- sprs::CsMat<f64> stand-in: CSR with usize indices ------------------------------------------
struct CsrMat {
    size_t rows = 0, cols = 0;
    std::vector<uint64_t> indptr, indices;
    std::vector<double> data;
    size_t nnz() const { return data.size(); }
};

// structural transpose as CSR (gemm.rs:44-46 `transpose_into().to_csr()`): counting transpose, the
// rows of the result come out with ascending column ids
inline CsrMat transpose(const CsrMat& a) {
    CsrMat t;
    t.rows = a.cols;
    t.cols = a.rows;
    t.indptr.assign(t.rows + 1, 0);
    t.indices.resize(a.nnz());
    t.data.resize(a.nnz());
    for (uint64_t c : a.indices) t.indptr[c + 1]++;
    for (size_t j = 0; j < t.rows; ++j) t.indptr[j + 1] += t.indptr[j];
    std::vector<uint64_t> cur(t.indptr.begin(), t.indptr.end() - 1);
    for (size_t i = 0; i < a.rows; ++i)
        for (uint64_t p = a.indptr[i]; p < a.indptr[i + 1]; ++p) {
            uint64_t d = cur[a.indices[p]]++;
            t.indices[d] = i;
            t.data[d] = a.data[p];
        }
    return t;
}

// ---- Matrix Market reader: what `scipy.io.mmread(f).tocsr()` returns for coordinate files ---------
// field real / integer / pattern, symmetry general / symmetric / skew-symmetric.  Array-format and
// complex files are errors (the reference panics on them: mmread gives an ndarray without .tocsr,
// py2rust.rs:74).  Entries are decimal text parsed with strtod (correctly rounded, like Python).
// Duplicate coordinates are summed (COO -> CSR); columns inside a row come out ascending.
inline CsrMat read_matrix_market(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("FileNotFoundError: [Errno 2] No such file or directory: '" + path + "'");
    std::string line;
    if (!std::getline(f, line)) throw std::runtime_error("ValueError: empty Matrix Market file");
    std::string lower = line;
    std::transform(lower.begin(), lower.end(), lower.begin(), ::tolower);
    std::istringstream hs(lower);
    std::string banner, object, format, field, symmetry;
    hs >> banner >> object >> format >> field >> symmetry;
    if (banner != "%%matrixmarket" || object != "matrix") throw std::runtime_error("ValueError: not a Matrix Market matrix file");
    if (format != "coordinate")
        throw std::runtime_error("AttributeError: 'numpy.ndarray' object has no attribute 'tocsr'");  // as upstream
    if (field != "real" && field != "integer" && field != "pattern" && field != "double")
        throw std::runtime_error("TypeError: unsupported Matrix Market field '" + field + "'");
    bool sym = symmetry == "symmetric" || symmetry == "hermitian", skew = symmetry == "skew-symmetric";
    if (!sym && !skew && symmetry != "general") throw std::runtime_error("ValueError: unknown symmetry '" + symmetry + "'");
    while (std::getline(f, line))
        if (!line.empty() && line[0] != '%' && line.find_first_not_of(" \t\r") != std::string::npos) break;
    size_t m = 0, n = 0, entries = 0;
    {
        std::istringstream ss(line);
        if (!(ss >> m >> n >> entries)) throw std::runtime_error("ValueError: bad Matrix Market size line");
    }
    std::vector<uint64_t> ri, ci;
    std::vector<double> v;
    ri.reserve(entries * (sym || skew ? 2 : 1));
    ci.reserve(ri.capacity());
    v.reserve(ri.capacity());
    // the rest of the file in one buffer, parsed with strtoull / strtod
    std::string rest((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const char* p = rest.c_str();
    char* end = nullptr;
    for (size_t e = 0; e < entries; ++e) {
        unsigned long long i = std::strtoull(p, &end, 10);
        if (end == p) throw std::runtime_error("ValueError: truncated Matrix Market file");
        p = end;
        unsigned long long j = std::strtoull(p, &end, 10);
        p = end;
        double val = 1.0;
        if (field != "pattern") {
            val = std::strtod(p, &end);
            if (end == p) throw std::runtime_error("ValueError: missing value in Matrix Market entry");
            p = end;
        }
        if (i < 1 || i > m || j < 1 || j > n) throw std::runtime_error("ValueError: Matrix Market index out of range");
        ri.push_back(i - 1);
        ci.push_back(j - 1);
        v.push_back(val);
        if ((sym || skew) && i != j) {
            ri.push_back(j - 1);
            ci.push_back(i - 1);
            v.push_back(skew ? -val : val);
        }
    }
    // COO -> canonical CSR: stable counting sort by row, then (column, arrival) sort inside each row, sum dups
    CsrMat a;
    a.rows = m;
    a.cols = n;
    a.indptr.assign(m + 1, 0);
    for (uint64_t r : ri) a.indptr[r + 1]++;
    for (size_t r = 0; r < m; ++r) a.indptr[r + 1] += a.indptr[r];
    std::vector<uint64_t> cur(a.indptr.begin(), a.indptr.end() - 1), cols(ri.size());
    std::vector<double> vals(ri.size());
    for (size_t e = 0; e < ri.size(); ++e) {
        uint64_t d = cur[ri[e]]++;
        cols[d] = ci[e];
        vals[d] = v[e];
    }
    std::vector<uint64_t> out_ptr(m + 1, 0);
    std::vector<size_t> order;
    for (size_t r = 0; r < m; ++r) {
        size_t s = a.indptr[r], e = a.indptr[r + 1];
        order.resize(e - s);
        std::iota(order.begin(), order.end(), s);
        std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return cols[x] < cols[y]; });
        for (size_t t = 0; t < order.size(); ++t) {
            size_t q = order[t];
            if (t > 0 && a.indices.back() == cols[q] && a.indices.size() > out_ptr[r]) {
                a.data.back() += vals[q];
            } else {
                a.indices.push_back(cols[q]);
                a.data.push_back(vals[q]);
            }
        }
        out_ptr[r + 1] = a.indices.size();
    }
    a.indptr = out_ptr;
    return a;
}

// py2rust.rs:62-97 (same two stdout lines as the embedded Python)
inline CsrMat load_mm_mat(const std::string& dir_path, const std::string& gemm_nm) {
    std::printf("---- Python Interface ----\n");
    std::printf("%% Load %s from %s\n", gemm_nm.c_str(), dir_path.c_str());
    std::fflush(stdout);
    std::string sep = (!dir_path.empty() && dir_path.back() == '/') ? "" : "/";
    return read_matrix_market(dir_path + sep + gemm_nm + ".mtx");
}

// ---- gemm.rs -----------------------------------------------------------------------------------
struct GEMM {
    std::string name;
    CsrMat a, b;
    bool b_is_a = false;  // square workloads: the reference clones A; here B aliases A
    static GEMM create(const std::string& name, CsrMat a, CsrMat b) {  // GEMM::new, gemm.rs:33-39
        GEMM g;
        g.name = name;
        g.a = std::move(a);
        g.b = std::move(b);
        return g;
    }
    static GEMM from_mat(const std::string& name, CsrMat mat) {  // gemm.rs:41-53
        GEMM g;
        g.name = name;
        if (mat.rows == mat.cols) {
            g.b_is_a = true;
        } else {
            g.b = transpose(mat);
        }
        g.a = std::move(mat);
        return g;
    }
    const CsrMat& B() const { return b_is_a ? a : b; }
    // gemm.rs:56-91, including its quirk of printing A's data/indices under "--B" (:79, :84)
    std::string display() const {
        const CsrMat& bb = B();
        auto head = [](size_t n) { return std::min<size_t>(n, 5); };
        std::ostringstream o;
        o << "---- " << name << " ----\n";
        o << "--A: (" << a.rows << ", " << a.cols << ")\n";
        o << "data: " << debug_list_f64(a.data.begin(), a.data.begin() + head(a.data.size())) << " .. \n";
        o << "indices: " << debug_list_int(a.indices.begin(), a.indices.begin() + head(a.indices.size())) << " ...\n";
        o << "indptr: " << debug_list_int(a.indptr.begin(), a.indptr.begin() + head(a.indptr.size())) << " ...\n";
        o << "--B: (" << bb.rows << ", " << bb.cols << ")\n";
        o << "data: " << debug_list_f64(a.data.begin(), a.data.begin() + head(bb.data.size())) << " ...\n";
        o << "indices: " << debug_list_int(a.indices.begin(), a.indices.begin() + head(bb.indices.size())) << " ...\n";
        o << "indptr: " << debug_list_int(bb.indptr.begin(), bb.indptr.begin() + head(bb.indptr.size())) << " ...\n";
        return o.str();
    }
};

// py2rust.rs:5-60: the pickle holds scipy/numpy objects, so a python3 helper (the same conversion rules
// as the reference's embedded snippet) turns the (A, B) pair into raw arrays this process reads back.
inline GEMM load_pickled_gemms(const std::string& gemm_fp, const std::string& gemm_nm) {
    std::string tmp = "/tmp/spada_b200_pkl_" + std::to_string((long)getpid()) + ".bin";
    std::string code =
        "import sys,pickle,numpy as np\n"
        "from scipy.sparse import coo_matrix,csr_matrix,csc_matrix\n"
        "fp,nm,out=sys.argv[1:4]\n"
        "print('---- Python Interface ----')\n"
        "print(f'% Load {nm} from', fp)\n"
        "def conv(x):\n"
        "    if isinstance(x,(csc_matrix,coo_matrix)): x=x.tocsr()\n"
        "    elif isinstance(x,np.ndarray): x=csr_matrix(x)\n"
        "    elif not isinstance(x,csr_matrix): raise TypeError('Unsupported matrix type: {}'.format(type(x)))\n"
        "    x=csr_matrix(x,dtype=np.float64); x.sum_duplicates(); x.sort_indices(); return x\n"
        "A,B=[conv(x) for x in pickle.load(open(fp,'rb'))[nm]]\n"
        "for t,m in (('A',A),('B',B)):\n"
        "    print(f'% -- {t} --'); print(f'% shape: {m.shape} data: {m.data[:5]}... indices: {m.indices[:5]}... indptr: {m.indptr[:5]}...')\n"
        "print('--- Return from Python Interface ---\\n')\n"
        "with open(out,'wb') as f:\n"
        "    for m in (A,B):\n"
        "        np.array([m.shape[0],m.shape[1],m.nnz],dtype='<u8').tofile(f)\n"
        "        m.indptr.astype('<u8').tofile(f); m.indices.astype('<u8').tofile(f); m.data.astype('<f8').tofile(f)\n";
    std::string script = tmp + ".py";
    {
        std::ofstream s(script);
        s << code;
    }
    std::fflush(stdout);
    std::string cmd = "python3 -W ignore '" + script + "' '" + gemm_fp + "' '" + gemm_nm + "' '" + tmp + "'";
    int rc = std::system(cmd.c_str());
    std::remove(script.c_str());
    if (rc != 0) throw std::runtime_error("load_pickled_gemms: python helper failed (see its message above)");
    std::ifstream f(tmp, std::ios::binary);
    auto read_mat = [&](CsrMat& m) {
        uint64_t hdr[3];
        f.read((char*)hdr, sizeof(hdr));
        m.rows = hdr[0];
        m.cols = hdr[1];
        m.indptr.resize(m.rows + 1);
        m.indices.resize(hdr[2]);
        m.data.resize(hdr[2]);
        f.read((char*)m.indptr.data(), 8 * m.indptr.size());
        f.read((char*)m.indices.data(), 8 * m.indices.size());
        f.read((char*)m.data.data(), 8 * m.data.size());
    };
    CsrMat a, b;
    read_mat(a);
    read_mat(b);
    f.close();
    std::remove(tmp.c_str());
    return GEMM::create(gemm_nm, std::move(a), std::move(b));
}

// ---- frontend.rs -------------------------------------------------------------------------------
struct OmegaConfig {
    std::string ss_filepath, nn_filepath;
    size_t pe_num = 0, at_num = 0, lane_num = 0, cache_size = 0, word_byte = 0;
    std::array<size_t, 2> block_shape{{0, 0}};
    size_t mem_latency = 0, cache_latency = 0;
    float freq = 0;
    size_t channel = 0;
    float bandwidth_per_channel = 0;
};

// flat JSON object with string / number / [number, number] values -- all OmegaConfig needs
inline OmegaConfig parse_config(const std::string& config_fp) {
    std::printf("%s\n", config_fp.c_str());  // frontend.rs:78
    std::fflush(stdout);
    std::ifstream f(config_fp);
    if (!f) throw std::runtime_error("No such file or directory (os error 2)");
    std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    std::map<std::string, std::string> kv;
    size_t i = 0;
    auto skip = [&]() { while (i < s.size() && (isspace((unsigned char)s[i]) || s[i] == ',')) ++i; };
    skip();
    if (i >= s.size() || s[i] != '{') throw std::runtime_error("expected value at line 1 column 1");
    ++i;
    while (true) {
        skip();
        if (i >= s.size()) throw std::runtime_error("EOF while parsing an object");
        if (s[i] == '}') break;
        if (s[i] != '"') throw std::runtime_error("key must be a string");
        size_t e = s.find('"', i + 1);
        std::string key = s.substr(i + 1, e - i - 1);
        i = s.find(':', e) + 1;
        skip();
        size_t st = i;
        if (s[i] == '"') {
            e = s.find('"', i + 1);
            kv[key] = s.substr(i + 1, e - i - 1);
            i = e + 1;
        } else if (s[i] == '[') {
            e = s.find(']', i);
            kv[key] = s.substr(i + 1, e - i - 1);
            i = e + 1;
        } else {
            while (i < s.size() && s[i] != ',' && s[i] != '}' && !isspace((unsigned char)s[i])) ++i;
            kv[key] = s.substr(st, i - st);
        }
    }
    auto need = [&](const char* k) -> const std::string& {
        auto it = kv.find(k);
        if (it == kv.end()) throw std::runtime_error(std::string("missing field `") + k + "`");
        return it->second;
    };
    OmegaConfig c;
    c.ss_filepath = need("ss_filepath");
    c.nn_filepath = need("nn_filepath");
    c.pe_num = std::stoull(need("pe_num"));
    c.at_num = std::stoull(need("at_num"));
    c.lane_num = std::stoull(need("lane_num"));
    c.cache_size = std::stoull(need("cache_size"));
    c.word_byte = std::stoull(need("word_byte"));
    {
        std::string bs = need("block_shape");
        std::replace(bs.begin(), bs.end(), ',', ' ');
        std::istringstream ss(bs);
        if (!(ss >> c.block_shape[0] >> c.block_shape[1])) throw std::runtime_error("invalid length, expected an array of length 2");
    }
    c.mem_latency = std::stoull(need("mem_latency"));
    c.cache_latency = std::stoull(need("cache_latency"));
    c.freq = std::stof(need("freq"));
    c.channel = std::stoull(need("channel"));
    c.bandwidth_per_channel = std::stof(need("bandwidth_per_channel"));
    return c;
}

struct Cli {
    std::string simulator, accelerator, category, workload, configuration;
    bool preprocess = false;
};
inline std::string match_enum(const std::string& v, std::initializer_list<const char*> variants) {
    std::string lv = v;
    std::transform(lv.begin(), lv.end(), lv.begin(), ::tolower);
    std::string all;
    for (const char* x : variants) {
        std::string lx = x;
        std::transform(lx.begin(), lx.end(), lx.begin(), ::tolower);
        if (lx == lv) return x;
        all += (all.empty() ? "" : ", ") + std::string(x);
    }
    throw std::invalid_argument("error: '" + v + "' isn't a valid value\n\t[possible values: " + all + "]");
}
inline Cli parse_args(int argc, char** argv) {  // frontend.rs:52-75
    Cli c;
    std::vector<std::string> pos;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        if (a == "-p" || a == "--preprocess") c.preprocess = true;
        else if (!a.empty() && a[0] == '-') throw std::invalid_argument("error: Found argument '" + a + "' which wasn't expected");
        else pos.push_back(a);
    }
    if (pos.size() != 5)
        throw std::invalid_argument(
            "error: The following required arguments were not provided:\n    <simulator> <accelerator> <category> "
            "<workload> <configuration>\n\nUSAGE:\n    spada-sim [FLAGS] <simulator> <accelerator> <category> <workload> "
            "<configuration>");
    c.simulator = match_enum(pos[0], {"AccurateSimu", "TrafficModel", "BReuseCounter"});
    c.accelerator = match_enum(pos[1], {"Ip", "Op", "MultiRow", "Spada"});
    c.category = match_enum(pos[2], {"SS", "NN"});
    c.workload = pos[3];
    c.configuration = pos[4];
    return c;
}

// ---- storage.rs --------------------------------------------------------------------------------
struct Element {
    std::array<size_t, 2> idx;
    double value;
};
struct CsrRow {
    size_t rowptr = 0;
    std::vector<double> data;
    std::vector<uint64_t> indptr;  // column ids (the reference's naming, storage.rs:34-39)
    static CsrRow new_from_data(size_t rowptr, std::vector<double> data, std::vector<uint64_t> indptr) {
        CsrRow r;
        r.rowptr = rowptr;
        r.data = std::move(data);
        r.indptr = std::move(indptr);
        return r;
    }
    size_t len() const { return indptr.size(); }
    std::string display() const {  // storage.rs:115-125
        size_t n = std::min<size_t>(data.size(), 5);
        return "rowptr: " + std::to_string(rowptr) + " indptr: " + debug_list_int(indptr.begin(), indptr.begin() + n) +
               " data: " + debug_list_f64(data.begin(), data.begin() + n);
    }
};
// ---- result sink (the reference prints ten rows of C and drops the rest, main.rs:113-116) ----------
// dump_result writes all of C as a Matrix Market coordinate file (1-based, %.17g: every f64 round-trips) and returns
// a one-line digest: nnz, sum of the values, FNV-1a-64 over the little-endian (indptr u64 | column ids u64 | values f64).
inline uint64_t fnv1a64(const void* p, size_t n, uint64_t h = 14695981039346656037ull) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) {
        h ^= b[i];
        h *= 1099511628211ull;
    }
    return h;
}
inline std::string dump_result(const std::string& path, const std::vector<CsrRow>& rows, size_t n_cols) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) throw std::runtime_error("cannot write " + path);
    size_t nnz = 0;
    for (const CsrRow& r : rows) nnz += r.len();
    std::fprintf(f, "%%%%MatrixMarket matrix coordinate real general\n%zu %zu %zu\n", rows.size(), n_cols, nnz);
    uint64_t h = 14695981039346656037ull, ptr = 0;
    double sum = 0.0;
    h = fnv1a64(&ptr, 8, h);
    for (const CsrRow& r : rows) {
        ptr += r.len();
        h = fnv1a64(&ptr, 8, h);
    }
    for (const CsrRow& r : rows) h = fnv1a64(r.indptr.data(), 8 * r.indptr.size(), h);
    for (const CsrRow& r : rows) {
        h = fnv1a64(r.data.data(), 8 * r.data.size(), h);
        for (size_t j = 0; j < r.len(); ++j) {
            std::fprintf(f, "%zu %llu %.17g\n", r.rowptr + 1, (unsigned long long)r.indptr[j] + 1, r.data[j]);
            sum += r.data[j];
        }
    }
    std::fclose(f);
    char buf[160];
    std::snprintf(buf, sizeof(buf), "C dumped: nnz %zu sum %.17g fnv1a64 %016llx", nnz, sum, (unsigned long long)h);
    return buf;
}

struct CsrMatStorage {  // storage.rs:150-160
    std::vector<double> data;
    std::vector<uint64_t> indptr, indices;
    std::array<size_t, 2> mat_shape{{0, 0}};  // [cols, rows] (sic, storage.rs:225, 236)
    bool remapped = false;
    std::vector<size_t> row_remap;
    size_t row_num() const { return indptr.size() - 1; }
    size_t get_ele_num(size_t s, size_t t) const { return (size_t)(indptr[t] - indptr[s]); }
    void reorder_row(std::vector<size_t> rowmap) {
        remapped = true;
        row_remap = std::move(rowmap);
    }
    static std::pair<CsrMatStorage, CsrMatStorage> init_with_gemm(const GEMM& g) {  // storage.rs:214-239
        auto mk = [](const CsrMat& m) {
            CsrMatStorage s;
            s.data = m.data;
            s.indptr = m.indptr;
            s.indices = m.indices;
            s.mat_shape = {m.cols, m.rows};
            return s;
        };
        return {mk(g.a), mk(g.B())};
    }
};
inline std::vector<size_t> sort_by_length(const CsrMatStorage& a) {  // preprocessing.rs:76-89
    std::vector<size_t> id(a.row_num());
    std::iota(id.begin(), id.end(), 0);
    std::stable_sort(id.begin(), id.end(), [&](size_t x, size_t y) { return a.get_ele_num(x, x + 1) < a.get_ele_num(y, y + 1); });
    return id;
}

// ---- simulator.rs: the operator interface of the hot path, backed by the engine ---------------------
class Simulator {
  public:
    Simulator(size_t /*pe_num*/, size_t /*at_num*/, size_t lane_num, size_t /*cache_size*/, size_t /*word_byte*/,
              size_t /*output_base_addr*/, std::array<size_t, 2> default_block_shape, CsrMatStorage& a_matrix,
              CsrMatStorage& b_matrix, const std::string& accelerator, size_t /*mem_latency*/ = 0,
              size_t /*cache_latency*/ = 0, float /*freq*/ = 1.f, size_t /*channel*/ = 1, float /*bw*/ = 1.f)
        : a_(a_matrix), b_(b_matrix) {
        spada_b200_opts o{};
        o.device = -1;
        o.accelerator = accelerator == "Ip" ? SPADA_B200_ACC_IP : accelerator == "Op" ? SPADA_B200_ACC_OP
                        : accelerator == "MultiRow" ? SPADA_B200_ACC_MULTIROW : SPADA_B200_ACC_SPADA;
        o.lane_num = (uint32_t)lane_num;
        o.block_shape[0] = (uint32_t)std::min<size_t>(default_block_shape[0], 0xffffffffu);
        o.block_shape[1] = (uint32_t)std::min<size_t>(default_block_shape[1], 0xffffffffu);
        o.flags = SPADA_B200_FLAG_VALIDATE;
        // SPADA_B200_GPUS=N: every product is sharded over N GPUs of this process (rows of A by equal product count, B
        // replicated, C gathered by the placement kernels' peer stores): spada_b200_group_* instead of one handle
        if (const char* e = std::getenv("SPADA_B200_GPUS")) n_gpus_ = std::max(1, std::atoi(e));
        if (n_gpus_ > 1) check(spada_b200_group_create(&o, (uint32_t)n_gpus_, &g_));
        else check(spada_b200_create(&o, &h_));  // the reference panics on every failure; so does this
    }
    ~Simulator() {
        if (r_) spada_b200_result_free(r_);
        if (h_) spada_b200_destroy(h_);
        if (g_) spada_b200_group_destroy(g_);
    }
    Simulator(const Simulator&) = delete;
    Simulator& operator=(const Simulator&) = delete;

    void execute() {  // simulator.rs:509-890
        spada_csr_view va{a_.mat_shape[1], a_.mat_shape[0], a_.data.size(), a_.indptr.data(), a_.indices.data(), a_.data.data()};
        spada_csr_view vb{b_.mat_shape[1], b_.mat_shape[0], b_.data.size(), b_.indptr.data(), b_.indices.data(), b_.data.data()};
        if (g_) check(spada_b200_group_spgemm(g_, &va, &vb, &r_));
        else check(spada_b200_spgemm(h_, &va, &vb, &r_));
        check(spada_b200_result_stats(r_, &st_));
        st_.nnz_a = a_.data.size();
    }
    std::vector<CsrRow> get_exec_result() {  // simulator.rs:1034-1062
        uint64_t rows, cols, nnz;
        check(spada_b200_result_shape(r_, &rows, &cols, &nnz));
        std::vector<uint64_t> ip(rows + 1), ix(nnz);
        std::vector<double> dx(nnz);
        check(spada_b200_result_copy(r_, ip.data(), ix.data(), dx.data()));
        std::vector<CsrRow> out;
        out.reserve(rows);
        for (uint64_t r = 0; r < rows; ++r)
            out.push_back(CsrRow::new_from_data(r, std::vector<double>(dx.begin() + ip[r], dx.begin() + ip[r + 1]),
                                                std::vector<uint64_t>(ix.begin() + ip[r], ix.begin() + ip[r + 1])));
        return out;
    }
    // simulator.rs:1008-1032 -- the cycle / traffic model is out of scope: analytic element counts with the
    // reference's accounting rules (storage.rs:313-315, :201-203)
    size_t get_exec_cycle() const { return 0; }
    std::array<size_t, 2> get_a_mat_stat() const { return {2 * (size_t)st_.nnz_a, 0}; }
    std::array<size_t, 2> get_b_mat_stat() const { return {2 * (size_t)st_.products, 0}; }
    std::array<size_t, 2> get_c_mat_stat() const { return {0, 2 * (size_t)st_.nnz_c + (size_t)st_.rows}; }
    std::array<size_t, 2> get_cache_stat() const { return {0, 0}; }
    const spada_b200_stats& engine_stats() const { return st_; }

  private:
    static void check(int rc) {
        if (rc != 0) throw std::runtime_error(std::string("spada_b200 error ") + std::to_string(rc) + ": " + spada_b200_last_error());
    }
    CsrMatStorage &a_, &b_;
    spada_b200_t* h_ = nullptr;
    spada_b200_group_t* g_ = nullptr;
    int n_gpus_ = 1;
    spada_b200_result_t* r_ = nullptr;
    spada_b200_stats st_{};
};

}  // namespace spada_host
