// On-disk contracts either side of the hot path (host code):
//   feature_cache/<sanitised path>.npy  -- C-order <f4 [n][60]                       (streamz-rs/src/lib.rs:550-579)
//   model.npz                           -- stored zip of npy members, w3/b3 per column (lib.rs:1081-1282)
// ndarray-npy 0.8.1 + zip 0.5.13 are not available here; this is a from-scratch reader/writer of the same container
// formats (NumPy .npy v1.0, ZIP "stored" entries), checked in tests against numpy.load / numpy.savez.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>

#include "common.cuh"
#include "mlp.cuh"

namespace szb {

static uint32_t crc32_of(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

// ---- npy ------------------------------------------------------------------------------------------------------------
static std::vector<uint8_t> npy_bytes(const char* descr, const std::vector<uint64_t>& shape, const void* data, size_t nbytes) {
    std::ostringstream h;
    h << "{'descr': '" << descr << "', 'fortran_order': False, 'shape': (";
    for (size_t i = 0; i < shape.size(); ++i) h << shape[i] << (shape.size() == 1 ? "," : (i + 1 < shape.size() ? ", " : ""));
    h << "), }";
    std::string hdr = h.str();
    const size_t pre = 10;
    size_t total = pre + hdr.size() + 1;
    const size_t pad = (64 - total % 64) % 64;
    hdr.append(pad, ' ');
    hdr.push_back('\n');
    std::vector<uint8_t> out;
    out.reserve(pre + hdr.size() + nbytes);
    const uint8_t magic[8] = { 0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0 };
    out.insert(out.end(), magic, magic + 8);
    out.push_back(uint8_t(hdr.size() & 0xFF));
    out.push_back(uint8_t(hdr.size() >> 8));
    out.insert(out.end(), hdr.begin(), hdr.end());
    const uint8_t* d = static_cast<const uint8_t*>(data);
    out.insert(out.end(), d, d + nbytes);
    return out;
}

struct NpyView {
    std::string descr;
    bool fortran = false;
    std::vector<uint64_t> shape;
    const uint8_t* data = nullptr;
    size_t nbytes = 0;
    // product of the dimensions, or UINT64_MAX when it (or count * itemsize) does not fit: callers compare against nbytes
    uint64_t count() const {
        uint64_t c = 1;
        for (auto s : shape) {
            if (s != 0 && c > (UINT64_MAX >> 4) / s) return UINT64_MAX;
            c *= s;
        }
        return c;
    }
    bool holds(uint64_t itemsize) const { const uint64_t c = count(); return c != UINT64_MAX && c <= nbytes / itemsize; }
};

static bool parse_npy(const uint8_t* p, size_t n, NpyView& v) {
    if (n < 10 || std::memcmp(p, "\x93NUMPY", 6) != 0) return false;
    const int major = p[6];
    size_t hlen, pre;
    if (major == 1) { hlen = p[8] | (size_t(p[9]) << 8); pre = 10; }
    else { if (n < 12) return false; hlen = p[8] | (size_t(p[9]) << 8) | (size_t(p[10]) << 16) | (size_t(p[11]) << 24); pre = 12; }
    if (pre + hlen > n) return false;
    const std::string h(reinterpret_cast<const char*>(p + pre), hlen);
    auto find_val = [&](const char* key) -> size_t {
        size_t k = h.find(key);
        if (k == std::string::npos) return k;
        k = h.find(':', k);
        return k == std::string::npos ? k : k + 1;
    };
    size_t a = find_val("'descr'");
    if (a == std::string::npos) return false;
    const size_t q0 = h.find('\'', a);
    if (q0 == std::string::npos) return false;
    const size_t q1 = h.find('\'', q0 + 1);
    if (q1 == std::string::npos) return false;
    v.descr = h.substr(q0 + 1, q1 - q0 - 1);
    a = find_val("'fortran_order'");
    if (a == std::string::npos) return false;
    const size_t fo = h.find_first_not_of(' ', a);
    if (fo == std::string::npos) return false;
    v.fortran = h.compare(fo, 4, "True") == 0;
    a = find_val("'shape'");
    if (a == std::string::npos) return false;
    const size_t p0 = h.find('(', a);
    if (p0 == std::string::npos) return false;
    const size_t p1 = h.find(')', p0);
    if (p1 == std::string::npos) return false;
    v.shape.clear();
    std::string dims = h.substr(p0 + 1, p1 - p0 - 1);
    std::stringstream ss(dims);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        size_t b = tok.find_first_not_of(' ');
        if (b == std::string::npos) continue;
        v.shape.push_back(std::strtoull(tok.c_str() + b, nullptr, 10));
    }
    v.data = p + pre + hlen;
    v.nbytes = n - pre - hlen;
    return true;
}

static bool read_file(const char* path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    const std::streamoff n = f.tellg();
    f.seekg(0);
    out.resize(size_t(n));
    if (n) f.read(reinterpret_cast<char*>(out.data()), n);
    return bool(f);
}
static bool write_file(const char* path, const std::vector<uint8_t>& d) {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) return false;
    if (!d.empty()) f.write(reinterpret_cast<const char*>(d.data()), std::streamsize(d.size()));
    return bool(f);
}

// ---- zip (stored entries only) ---------------------------------------------------------------------------------------
struct ZipWriter {
    std::vector<uint8_t> buf;
    struct Entry { std::string name; uint32_t crc; uint64_t size, offset; };
    std::vector<Entry> entries;
    // Classic (non-zip64) container: at most 65 535 members, every size and offset below 4 GiB.  A model with ~21 800
    // speakers (three members each) reaches the first limit; finish() reports it instead of writing a corrupt file.
    bool fits() const {
        if (entries.size() > 0xFFFFu || buf.size() > 0xFFFFFFFFull) return false;
        for (const Entry& e : entries)
            if (e.size > 0xFFFFFFFFull || e.offset > 0xFFFFFFFFull || e.name.size() > 0xFFFFu) return false;
        return true;
    }
    static void p16(std::vector<uint8_t>& b, uint16_t v) { b.push_back(uint8_t(v)); b.push_back(uint8_t(v >> 8)); }
    static void p32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back(uint8_t(v >> (8 * i))); }
    void add(const std::string& name, const std::vector<uint8_t>& data) {
        Entry e{ name, crc32_of(data.data(), data.size()), uint64_t(data.size()), uint64_t(buf.size()) };
        p32(buf, 0x04034b50u); p16(buf, 20); p16(buf, 0); p16(buf, 0); p16(buf, 0); p16(buf, 0x21);
        p32(buf, e.crc); p32(buf, uint32_t(e.size)); p32(buf, uint32_t(e.size)); p16(buf, uint16_t(name.size())); p16(buf, 0);
        buf.insert(buf.end(), name.begin(), name.end());
        buf.insert(buf.end(), data.begin(), data.end());
        entries.push_back(e);
    }
    bool finish() {
        if (!fits()) return false;
        const uint32_t cd_off = uint32_t(buf.size());
        for (const Entry& e : entries) {
            p32(buf, 0x02014b50u); p16(buf, 20); p16(buf, 20); p16(buf, 0); p16(buf, 0); p16(buf, 0); p16(buf, 0x21);
            p32(buf, e.crc); p32(buf, uint32_t(e.size)); p32(buf, uint32_t(e.size)); p16(buf, uint16_t(e.name.size())); p16(buf, 0);
            p16(buf, 0); p16(buf, 0); p16(buf, 0); p32(buf, 0); p32(buf, uint32_t(e.offset));
            buf.insert(buf.end(), e.name.begin(), e.name.end());
        }
        if (buf.size() > 0xFFFFFFFFull) return false;
        const uint32_t cd_size = uint32_t(buf.size()) - cd_off;
        p32(buf, 0x06054b50u); p16(buf, 0); p16(buf, 0); p16(buf, uint16_t(entries.size())); p16(buf, uint16_t(entries.size()));
        p32(buf, cd_size); p32(buf, cd_off); p16(buf, 0);
        return true;
    }
};

static uint16_t g16(const uint8_t* p) { return uint16_t(p[0] | (p[1] << 8)); }
static uint32_t g32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }

// name -> (pointer, size) of each stored member
static szb_status zip_index(const std::vector<uint8_t>& z, std::map<std::string, std::pair<const uint8_t*, size_t>>& out) {
    if (z.size() < 22) { set_error("npz: file too small"); return SZB_ERR_IO; }
    size_t eocd = std::string::npos;
    for (size_t i = z.size() - 22; ; --i) {
        if (g32(&z[i]) == 0x06054b50u) { eocd = i; break; }
        if (i == 0 || z.size() - i > 65557) break;
    }
    if (eocd == std::string::npos) { set_error("npz: end-of-central-directory not found"); return SZB_ERR_IO; }
    const uint32_t n = g16(&z[eocd + 10]);
    const uint64_t zs = z.size();
    uint64_t p = g32(&z[eocd + 16]);                          // every offset is widened before anything is added to it
    for (uint32_t i = 0; i < n; ++i) {
        if (p + 46 > zs || g32(&z[p]) != 0x02014b50u) { set_error("npz: bad central directory"); return SZB_ERR_IO; }
        const uint16_t method = g16(&z[p + 10]);
        const uint64_t csize = g32(&z[p + 20]), usize = g32(&z[p + 24]);
        const uint64_t nlen = g16(&z[p + 28]), xlen = g16(&z[p + 30]), clen = g16(&z[p + 32]);
        const uint64_t lho = g32(&z[p + 42]);
        if (p + 46 + nlen > zs) { set_error("npz: member name runs past the end of the file"); return SZB_ERR_IO; }
        const std::string name(reinterpret_cast<const char*>(&z[p + 46]), size_t(nlen));
        if (method != 0) { set_error("npz: member '%s' is compressed (method %u); only stored members are supported", name.c_str(), method); return SZB_ERR_UNSUPPORTED; }
        if (lho + 30 > zs || g32(&z[lho]) != 0x04034b50u) { set_error("npz: bad local header"); return SZB_ERR_IO; }
        const uint64_t data = lho + 30 + g16(&z[lho + 26]) + g16(&z[lho + 28]);
        if (data > zs || csize > zs - data || csize != usize) { set_error("npz: truncated member '%s'", name.c_str()); return SZB_ERR_IO; }
        out[name] = { &z[data], size_t(usize) };
        p += 46 + nlen + xlen + clen;
    }
    return SZB_OK;
}

}  // namespace szb

using namespace szb;

// Nothing may throw across the C ABI (std::bad_alloc from a crafted header, std::out_of_range from a string operation).
#define SZB_GUARD_IO(body)                                                       \
    try { body }                                                                 \
    catch (const std::exception& e) { set_error("%s: %s", __func__, e.what()); return SZB_ERR_IO; } \
    catch (...) { set_error("%s: unknown exception", __func__); return SZB_ERR_IO; }

static szb_status npy_write_impl(const char* path, const float* data, uint64_t rows, uint64_t cols);
static szb_status npy_read_impl(const char* path, float* data, uint64_t cap_elems, uint64_t* rows, uint64_t* cols);
static szb_status net_save_impl(szb_net* net, const char* path, uint32_t sample_rate, uint32_t bits);
static szb_status net_load_impl(szb_ctx* ctx, const char* path, szb_net** out, uint32_t* sample_rate, uint32_t* bits);

extern "C" {

// lib.rs:550-555: "feature_cache/" + path with '/' and '\\' replaced by '_' + ".npy"
szb_status szb_feature_cache_path(const char* audio_path, char* out, size_t cap) {
    SZB_REQUIRE(audio_path && out, "szb_feature_cache_path: NULL argument");
    std::string s(audio_path);
    for (char& c : s) if (c == '/' || c == '\\') c = '_';
    s = "feature_cache/" + s + ".npy";
    SZB_REQUIRE(cap > s.size(), "szb_feature_cache_path: capacity %zu <= %zu", cap, s.size());
    std::memcpy(out, s.c_str(), s.size() + 1);
    return SZB_OK;
}

szb_status szb_npy_write_f32(const char* path, const float* data, uint64_t rows, uint64_t cols) {
    SZB_GUARD_IO(return npy_write_impl(path, data, rows, cols);)
}
szb_status szb_npy_read_f32(const char* path, float* data, uint64_t cap_elems, uint64_t* rows, uint64_t* cols) {
    SZB_GUARD_IO(return npy_read_impl(path, data, cap_elems, rows, cols);)
}
szb_status szb_net_save(szb_net* net, const char* path, uint32_t sample_rate, uint32_t bits) {
    SZB_GUARD_IO(return net_save_impl(net, path, sample_rate, bits);)
}
szb_status szb_net_load(szb_ctx* ctx, const char* path, szb_net** out, uint32_t* sample_rate, uint32_t* bits) {
    SZB_GUARD_IO(return net_load_impl(ctx, path, out, sample_rate, bits);)
}

}  // extern "C"

static szb_status npy_write_impl(const char* path, const float* data, uint64_t rows, uint64_t cols) {
    SZB_REQUIRE(path && (data || rows * cols == 0), "szb_npy_write_f32: NULL argument");
    SZB_REQUIRE(cols == 0 || rows <= (UINT64_MAX >> 4) / cols, "szb_npy_write_f32: %llu x %llu elements overflow",
                (unsigned long long)rows, (unsigned long long)cols);
    const auto bytes = npy_bytes("<f4", { rows, cols }, data, size_t(rows * cols) * 4);
    if (!write_file(path, bytes)) { set_error("cannot write %s", path); return SZB_ERR_IO; }
    return SZB_OK;
}

static szb_status npy_read_impl(const char* path, float* data, uint64_t cap_elems, uint64_t* rows, uint64_t* cols) {
    SZB_REQUIRE(path && rows && cols, "szb_npy_read_f32: NULL argument");
    std::vector<uint8_t> raw;
    if (!read_file(path, raw)) { set_error("cannot read %s", path); return SZB_ERR_IO; }
    NpyView v;
    if (!parse_npy(raw.data(), raw.size(), v)) { set_error("%s is not a valid .npy file", path); return SZB_ERR_IO; }
    if (v.descr != "<f4" || v.fortran || v.shape.size() != 2) {
        set_error("%s: expected C-order <f4 2-D array, got descr '%s' ndim %zu", path, v.descr.c_str(), v.shape.size());
        return SZB_ERR_IO;
    }
    *rows = v.shape[0];
    *cols = v.shape[1];
    if (!v.holds(4)) { set_error("%s: truncated", path); return SZB_ERR_IO; }
    if (!data) return SZB_OK;  // size query
    SZB_REQUIRE(cap_elems >= v.count(), "szb_npy_read_f32: capacity %llu < %llu", (unsigned long long)cap_elems, (unsigned long long)v.count());
    std::memcpy(data, v.data, size_t(v.count()) * 4);
    return SZB_OK;
}

extern "C" {

szb_status szb_net_record_training_file(szb_net* net, uint32_t speaker, const char* path) {
    SZB_REQUIRE(net && path, "szb_net_record_training_file: NULL argument");
    if (net->file_lists.size() <= speaker) net->file_lists.resize(size_t(speaker) + 1);   // lib.rs:856-858
    auto& l = net->file_lists[speaker];
    if (std::find(l.begin(), l.end(), std::string(path)) == l.end()) l.emplace_back(path);  // lib.rs:859-861
    return SZB_OK;
}

szb_status szb_net_file_list(const szb_net* net, uint32_t speaker, char* out, size_t cap, size_t* len) {
    SZB_REQUIRE(net && len, "szb_net_file_list: NULL argument");
    std::string joined;
    if (speaker < net->file_lists.size())
        for (size_t i = 0; i < net->file_lists[speaker].size(); ++i) joined += (i ? "\n" : "") + net->file_lists[speaker][i];
    *len = joined.size();
    if (!out) return SZB_OK;
    SZB_REQUIRE(cap > joined.size(), "szb_net_file_list: capacity %zu <= %zu", cap, joined.size());
    std::memcpy(out, joined.c_str(), joined.size() + 1);
    return SZB_OK;
}

// set_embeddings / embeddings (lib.rs:869-877): (embedding, mean similarity, std similarity) per speaker
szb_status szb_net_set_embeddings(szb_net* net, const float* emb, const float* mean_sims, const float* std_sims, uint32_t n, uint32_t dim) {
    SZB_REQUIRE(net && (n == 0 || (emb && mean_sims && std_sims && dim > 0)), "szb_net_set_embeddings: NULL argument");
    SZB_GUARD_IO(
        net->emb.assign(emb, emb + size_t(n) * dim);
        net->emb_mean.assign(mean_sims, mean_sims + n);
        net->emb_std.assign(std_sims, std_sims + n);
        net->emb_n = n;
        net->emb_dim = n ? dim : 0;
        return SZB_OK;)
}
szb_status szb_net_get_embeddings(const szb_net* net, float* emb, float* mean_sims, float* std_sims, uint32_t cap_n, uint32_t* n,
                                  uint32_t* dim) {
    SZB_REQUIRE(net && n && dim, "szb_net_get_embeddings: NULL argument");
    *n = net->emb_n;
    *dim = net->emb_dim;
    if (!emb && !mean_sims && !std_sims) return SZB_OK;                  // size query
    SZB_REQUIRE(cap_n >= net->emb_n, "szb_net_get_embeddings: capacity %u < %u", cap_n, net->emb_n);
    if (emb && net->emb_n) std::memcpy(emb, net->emb.data(), net->emb.size() * 4);
    if (mean_sims && net->emb_n) std::memcpy(mean_sims, net->emb_mean.data(), net->emb_mean.size() * 4);
    if (std_sims && net->emb_n) std::memcpy(std_sims, net->emb_std.data(), net->emb_std.size() * 4);
    return SZB_OK;
}
// The optional hidden encoding layer w4 [rows][n] / b4 [n] (lib.rs:752-754, set_output_layer lib.rs:835-848): carried, never
// evaluated by this library (SURVEY.md section 2 #16 is out of scope); n == 0 removes it.
szb_status szb_net_set_encoding_layer(szb_net* net, const float* w4, const float* b4, uint32_t rows, uint32_t n) {
    SZB_REQUIRE(net && (n == 0 || (w4 && b4 && rows > 0)), "szb_net_set_encoding_layer: NULL argument");
    SZB_GUARD_IO(
        net->w4.assign(w4, w4 + size_t(n ? rows : 0) * n);
        net->b4.assign(b4, b4 + n);
        net->w4_rows = n ? rows : 0;
        return SZB_OK;)
}
szb_status szb_net_get_encoding_layer(const szb_net* net, float* w4, float* b4, uint64_t cap_elems, uint32_t* rows, uint32_t* n) {
    SZB_REQUIRE(net && rows && n, "szb_net_get_encoding_layer: NULL argument");
    *rows = net->w4_rows;
    *n = uint32_t(net->b4.size());
    if (!w4 && !b4) return SZB_OK;                                       // size query
    SZB_REQUIRE(cap_elems >= net->w4.size(), "szb_net_get_encoding_layer: capacity %llu < %zu", (unsigned long long)cap_elems, net->w4.size());
    if (w4 && !net->w4.empty()) std::memcpy(w4, net->w4.data(), net->w4.size() * 4);
    if (b4 && !net->b4.empty()) std::memcpy(b4, net->b4.data(), net->b4.size() * 4);
    return SZB_OK;
}

}  // extern "C"

static szb_status net_save_impl(szb_net* net, const char* path, uint32_t sample_rate, uint32_t bits) {
    SZB_REQUIRE(net && path, "szb_net_save: NULL argument");
    const uint32_t I = net->n_in, H1 = net->h1, H2 = net->h2, C = net->n_out;
    std::vector<float> w1(size_t(I) * H1), b1(H1), w2(size_t(H1) * H2), b2(H2), w3(size_t(H2) * C), b3(C);
    SZB_TRY(szb_net_get_weights(net, w1.data(), b1.data(), w2.data(), b2.data(), w3.data(), b3.data()));
    ZipWriter z;
    // member order and names exactly as lib.rs:1084-1113 (bare names, no ".npy" suffix)
    z.add("w1", npy_bytes("<f4", { I, H1 }, w1.data(), w1.size() * 4));
    z.add("b1", npy_bytes("<f4", { H1 }, b1.data(), b1.size() * 4));
    z.add("w2", npy_bytes("<f4", { H1, H2 }, w2.data(), w2.size() * 4));
    z.add("b2", npy_bytes("<f4", { H2 }, b2.data(), b2.size() * 4));
    const int64_t sr = sample_rate, bt = bits, ns = C;
    z.add("sample_rate", npy_bytes("<i8", { 1 }, &sr, 8));
    z.add("bits", npy_bytes("<i8", { 1 }, &bt, 8));
    z.add("num_speakers", npy_bytes("<i8", { 1 }, &ns, 8));
    std::vector<float> col(H2);
    for (uint32_t k = 0; k < C; ++k) {                                   // lib.rs:1091-1098: one column per member
        for (uint32_t r = 0; r < H2; ++r) col[r] = w3[size_t(r) * C + k];
        z.add("w3_" + std::to_string(k + 1), npy_bytes("<f4", { H2 }, col.data(), col.size() * 4));
        z.add("b3_" + std::to_string(k + 1), npy_bytes("<f4", { 1 }, &b3[k], 4));
    }
    if (!net->b4.empty()) {                                              // lib.rs:1099-1108: w4_k / b4_k, one column each
        const uint32_t n4 = uint32_t(net->b4.size()), r4 = net->w4_rows;
        std::vector<float> c4(r4);
        for (uint32_t k = 0; k < n4; ++k) {
            for (uint32_t r = 0; r < r4; ++r) c4[r] = net->w4[size_t(r) * n4 + k];
            z.add("w4_" + std::to_string(k + 1), npy_bytes("<f4", { r4 }, c4.data(), c4.size() * 4));
            z.add("b4_" + std::to_string(k + 1), npy_bytes("<f4", { 1 }, &net->b4[k], 4));
        }
    }
    for (uint32_t i = 0; i < C; ++i) {                                   // lib.rs:1109-1113
        std::string joined;
        if (i < net->file_lists.size())
            for (size_t f = 0; f < net->file_lists[i].size(); ++f) joined += (f ? "\n" : "") + net->file_lists[i][f];
        z.add("speaker_" + std::to_string(i) + "_files", npy_bytes("|u1", { joined.size() }, joined.data(), joined.size()));
    }
    if (net->emb_n > 0) {                                                // lib.rs:1114-1127
        z.add("speaker_embeddings", npy_bytes("<f4", { net->emb_n, net->emb_dim }, net->emb.data(), net->emb.size() * 4));
        z.add("speaker_mean_sims", npy_bytes("<f4", { net->emb_n }, net->emb_mean.data(), net->emb_mean.size() * 4));
        z.add("speaker_std_sims", npy_bytes("<f4", { net->emb_n }, net->emb_std.data(), net->emb_std.size() * 4));
    }
    if (!z.finish()) {
        set_error("szb_net_save: %zu members / %zu bytes do not fit a classic zip (65 535 members, 4 GiB); zip64 is not written",
                  z.entries.size(), z.buf.size());
        return SZB_ERR_UNSUPPORTED;
    }
    if (!write_file(path, z.buf)) { set_error("cannot write %s", path); return SZB_ERR_IO; }
    return SZB_OK;
}

static szb_status net_load_impl(szb_ctx* ctx, const char* path, szb_net** out, uint32_t* sample_rate, uint32_t* bits) {
    SZB_REQUIRE(ctx && path && out, "szb_net_load: NULL argument");
    std::vector<uint8_t> raw;
    if (!read_file(path, raw)) { set_error("cannot read %s", path); return SZB_ERR_IO; }
    std::map<std::string, std::pair<const uint8_t*, size_t>> idx;
    SZB_TRY(zip_index(raw, idx));
    auto has = [&](const std::string& n) { return idx.count(n) || idx.count(n + ".npy"); };  // accept both spellings
    auto get = [&](const std::string& n, NpyView& v) -> bool {
        auto it = idx.find(n);
        if (it == idx.end()) it = idx.find(n + ".npy");
        if (it == idx.end()) return false;
        return parse_npy(it->second.first, it->second.second, v) && !v.fortran;
    };
    auto getf = [&](const std::string& n, std::vector<float>& dst, std::vector<uint64_t>* shape) -> bool {
        NpyView v;
        if (!get(n, v) || v.descr != "<f4" || !v.holds(4)) return false;
        dst.resize(size_t(v.count()));
        std::memcpy(dst.data(), v.data, dst.size() * 4);
        if (shape) *shape = v.shape;
        return true;
    };
    auto geti = [&](const std::string& n, int64_t& dst) -> bool {
        NpyView v;
        if (!get(n, v) || v.descr != "<i8" || v.nbytes < 8) return false;
        std::memcpy(&dst, v.data, 8);
        return true;
    };
    std::vector<float> w1, b1, w2, b2;
    std::vector<uint64_t> s1, s2;
    int64_t sr = 0, bt = 0;
    if (!geti("sample_rate", sr) || !geti("bits", bt) || !getf("w1", w1, &s1) || !getf("b1", b1, nullptr) || !getf("w2", w2, &s2) ||
        !getf("b2", b2, nullptr) || s1.size() != 2 || s2.size() != 2 || s1[1] != s2[0]) {
        set_error("%s: missing or malformed w1/b1/w2/b2/sample_rate/bits (lib.rs:1136-1141)", path);
        return SZB_ERR_IO;
    }
    const uint32_t I = uint32_t(s1[0]), H1 = uint32_t(s1[1]), H2 = uint32_t(s2[1]);
    std::vector<std::vector<float>> cols;
    std::vector<float> biases;
    for (uint32_t k = 1;; ++k) {                                          // lib.rs:1151-1166
        const std::string wn = "w3_" + std::to_string(k), bn = "b3_" + std::to_string(k);
        if (!has(wn) || !has(bn)) break;
        std::vector<float> c, b;
        if (!getf(wn, c, nullptr) || !getf(bn, b, nullptr) || c.size() != H2 || b.empty()) {
            set_error("%s: malformed %s / %s", path, wn.c_str(), bn.c_str());
            return SZB_ERR_IO;
        }
        cols.push_back(std::move(c));
        biases.push_back(b[0]);
    }
    uint32_t C = uint32_t(cols.size());
    std::vector<float> w3, b3;
    if (C > 0) {
        w3.assign(size_t(H2) * C, 0.f);
        b3 = biases;
        for (uint32_t k = 0; k < C; ++k)
            for (uint32_t r = 0; r < H2; ++r) w3[size_t(r) * C + k] = cols[k][r];
    } else if (has("w3")) {                                               // legacy dense pair, lib.rs:1199-1207
        std::vector<uint64_t> s3;
        if (!getf("w3", w3, &s3) || !getf("b3", b3, nullptr) || s3.size() != 2 || s3[0] != H2 || s3[1] < b3.size()) {
            set_error("%s: malformed legacy w3 / b3", path);
            return SZB_ERR_IO;
        }
        C = uint32_t(b3.size());
        if (s3[1] != C) {  // keep the first C columns
            std::vector<float> t(size_t(H2) * C);
            for (uint32_t r = 0; r < H2; ++r)
                for (uint32_t k = 0; k < C; ++k) t[size_t(r) * C + k] = w3[size_t(r) * s3[1] + k];
            w3.swap(t);
        }
    }
    int64_t ns = 0;
    const bool has_ns = geti("num_speakers", ns);                        // optional, lib.rs:1142-1147
    const uint32_t outputs = has_ns ? uint32_t(ns) : C;                   // lib.rs:1227-1233
    if (outputs == 0 || outputs > C) {
        // the reference would build a net whose forward slices w3[.., ..num_speakers] out of range and panic
        set_error("%s: num_speakers %u is not covered by the %u stored output columns", path, outputs, C);
        return SZB_ERR_IO;
    }
    if (outputs < C) {  // forward uses only the first `outputs` columns (lib.rs:884-885)
        std::vector<float> t(size_t(H2) * outputs);
        for (uint32_t r = 0; r < H2; ++r)
            for (uint32_t k = 0; k < outputs; ++k) t[size_t(r) * outputs + k] = w3[size_t(r) * C + k];
        w3.swap(t);
        b3.resize(outputs);
    }
    if (b1.size() != H1 || b2.size() != H2) { set_error("%s: bias shapes do not match", path); return SZB_ERR_IO; }
    SZB_TRY(szb_net_from_weights(ctx, I, H1, H2, outputs, w1.data(), b1.data(), w2.data(), b2.data(), w3.data(), b3.data(), out));
    for (uint32_t i = 0; i < outputs; ++i) {                              // we read both spellings
        NpyView v;
        if (get("speaker_" + std::to_string(i) + "_files", v) && v.descr == "|u1") {
            std::string text(reinterpret_cast<const char*>(v.data), size_t(std::min<uint64_t>(v.count(), v.nbytes)));
            std::stringstream ss(text);
            std::string line;
            while (std::getline(ss, line))
                if (!line.empty()) szb_net_record_training_file(*out, i, line.c_str());
        }
    }
    {   // optional hidden encoding layer, lib.rs:1168-1186, 1209-1226: columns w4_k with their own length
        std::vector<std::vector<float>> c4;
        std::vector<float> b4;
        for (uint32_t k = 1;; ++k) {
            const std::string wn = "w4_" + std::to_string(k), bn = "b4_" + std::to_string(k);
            if (!has(wn) || !has(bn)) break;
            std::vector<float> c, b;
            if (!getf(wn, c, nullptr) || !getf(bn, b, nullptr) || b.empty() || (!c4.empty() && c.size() != c4[0].size())) {
                set_error("%s: malformed %s / %s", path, wn.c_str(), bn.c_str());
                szb_net_destroy(*out); *out = nullptr;
                return SZB_ERR_IO;
            }
            c4.push_back(std::move(c));
            b4.push_back(b[0]);
        }
        if (!c4.empty()) {
            const uint32_t r4 = uint32_t(c4[0].size()), n4 = uint32_t(c4.size());
            std::vector<float> w4(size_t(r4) * n4);
            for (uint32_t k = 0; k < n4; ++k)
                for (uint32_t r = 0; r < r4; ++r) w4[size_t(r) * n4 + k] = c4[k][r];
            (*out)->w4.swap(w4); (*out)->b4.swap(b4); (*out)->w4_rows = r4;
        }
    }
    if (has("speaker_embeddings")) {                                       // lib.rs:1253-1264
        std::vector<float> e, m, sd;
        std::vector<uint64_t> es;
        if (!getf("speaker_embeddings", e, &es) || !getf("speaker_mean_sims", m, nullptr) || !getf("speaker_std_sims", sd, nullptr) ||
            es.size() != 2 || m.size() < es[0] || sd.size() < es[0]) {
            set_error("%s: malformed speaker_embeddings / speaker_mean_sims / speaker_std_sims", path);
            szb_net_destroy(*out); *out = nullptr;
            return SZB_ERR_IO;
        }
        m.resize(size_t(es[0])); sd.resize(size_t(es[0]));
        (*out)->emb.swap(e); (*out)->emb_mean.swap(m); (*out)->emb_std.swap(sd);
        (*out)->emb_n = uint32_t(es[0]); (*out)->emb_dim = es[0] ? uint32_t(es[1]) : 0;
    }
    if (sample_rate) *sample_rate = uint32_t(sr);
    if (bits) *bits = uint32_t(bt);
    return SZB_OK;
}
