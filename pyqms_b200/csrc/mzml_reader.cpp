// Native mzML ingest (include/qms_mzml.h): memory-mapped file, one tag scan to index the spectra, parallel decode of
// the binary arrays into the packed (N, 2) layout.  Host-only; built as libqms_mzml.so next to libqms_b200.so.
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <math.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <numeric>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/qms_mzml.h"

namespace {

// memmem with a string literal as the needle (its length without the terminating NUL)
#define FIND_LIT(hay, n, lit) ((const char*)memmem((hay), (n), (lit), sizeof(lit) - 1))

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

struct Span {
    const char* a = nullptr;
    const char* b = nullptr;
    bool empty() const { return a == b; }
    size_t size() const { return (size_t)(b - a); }
    bool equals(const char* s) const { return size() == strlen(s) && memcmp(a, s, size()) == 0; }
    std::string str() const { return std::string(a, b); }
};

enum Packing { PACK_FLOAT, PACK_INT32, PACK_INT64, PACK_NUMPRESS_LINEAR, PACK_NUMPRESS_PIC, PACK_NUMPRESS_SLOF };

struct ArrayInfo {
    Span text;            // base64 payload
    bool present = false, f32 = false, zlib = false;
    Packing packing = PACK_FLOAT;     // a numpress / integer accession wins over the float width, wherever it stands
    int64_t length = -1;  // arrayLength attribute, if given
};

struct Param {            // one cvParam that matters (accession + value + unitName)
    Span accession, value, unit;
};

struct SpectrumInfo {
    int64_t ordinal = 0;          // position among all <spectrum> elements of the file
    Span native_id;
    int32_t level = -1;
    bool has_rt = false, rt_seconds = false;
    double rt = 0.0;
    int64_t default_length = -1;
    ArrayInfo mz, intensity;
    bool unsupported = false;
};

enum ArrayKind { KIND_NONE, KIND_MZ, KIND_INTENSITY };

}  // namespace

struct qms_mzml {
    int fd = -1;
    const char* data = nullptr;       // the mapped file, or `inflated` for a gzip-compressed one
    size_t size = 0;
    bool mapped = false;
    std::vector<char> inflated;
    std::vector<SpectrumInfo> spectra;
    int scan_blocks = 0;              // blocks the tag scan was split into at the file's own spectrum offsets (0: sequential scan)
};

namespace {

// value of attribute `name` inside the tag text [a, b) (after the tag name); empty span if absent
Span attribute(const char* a, const char* b, const char* name) {
    const size_t n = strlen(name);
    for (const char* p = a; p + n + 2 < b; ++p) {
        if ((p == a || p[-1] == ' ' || p[-1] == '\t' || p[-1] == '\n' || p[-1] == '\r') && memcmp(p, name, n) == 0) {
            const char* q = p + n;
            while (q < b && (*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r')) ++q;
            if (q >= b || *q != '=') continue;
            ++q;
            while (q < b && (*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r')) ++q;
            if (q >= b || (*q != '"' && *q != '\'')) continue;
            const char quote = *q++;
            const char* e = (const char*)memchr(q, quote, (size_t)(b - q));
            if (!e) return Span{};
            return Span{q, e};
        }
    }
    return Span{};
}

int64_t to_int(Span s, int64_t fallback) {
    if (s.empty()) return fallback;
    char buf[32];
    const size_t n = std::min(s.size(), sizeof(buf) - 1);
    memcpy(buf, s.a, n);
    buf[n] = 0;
    char* end = nullptr;
    const long long v = strtoll(buf, &end, 10);
    return end == buf ? fallback : (int64_t)v;
}

bool to_double(Span s, double& out) {
    if (s.empty()) return false;
    char buf[64];
    const size_t n = std::min(s.size(), sizeof(buf) - 1);
    memcpy(buf, s.a, n);
    buf[n] = 0;
    char* end = nullptr;
    out = strtod(buf, &end);
    return end != buf;
}

// what one cvParam means for the spectrum / the binary array that is open
void apply_param(const Param& p, SpectrumInfo& s, ArrayInfo* array, ArrayKind* kind) {
    const Span& acc = p.accession;
    if (acc.equals("MS:1000511")) {
        s.level = (int32_t)to_int(p.value, -1);
    } else if (acc.equals("MS:1000016")) {
        if (!s.has_rt) {
            double v;
            if (to_double(p.value, v)) {
                s.has_rt = true;
                s.rt = v;
                // unitName 'second' / 'minute' (missing: minute), compared case-insensitively on its first three letters
                s.rt_seconds = p.unit.size() >= 3 && (p.unit.a[0] | 32) == 's' && (p.unit.a[1] | 32) == 'e' && (p.unit.a[2] | 32) == 'c';
            }
        }
    } else if (array) {
        if (acc.equals("MS:1000514")) *kind = KIND_MZ;
        else if (acc.equals("MS:1000515")) *kind = KIND_INTENSITY;
        else if (acc.equals("MS:1000521")) array->f32 = true;
        else if (acc.equals("MS:1000523")) array->f32 = false;
        else if (acc.equals("MS:1000574")) array->zlib = true;
        else if (acc.equals("MS:1000576")) array->zlib = false;
        else if (acc.equals("MS:1000519")) array->packing = PACK_INT32;
        else if (acc.equals("MS:1000522")) array->packing = PACK_INT64;
        else if (acc.equals("MS:1002312")) array->packing = PACK_NUMPRESS_LINEAR;
        else if (acc.equals("MS:1002313")) array->packing = PACK_NUMPRESS_PIC;
        else if (acc.equals("MS:1002314")) array->packing = PACK_NUMPRESS_SLOF;
        else if (acc.equals("MS:1002746") || acc.equals("MS:1002747") || acc.equals("MS:1002748")) {     // numpress, then zlib
            array->packing = acc.equals("MS:1002746") ? PACK_NUMPRESS_LINEAR : acc.equals("MS:1002747") ? PACK_NUMPRESS_PIC : PACK_NUMPRESS_SLOF;
            array->zlib = true;
        } else if (acc.equals("MS:1003088") || acc.equals("MS:1003089") || acc.equals("MS:1003090")) {   // truncation codecs
            s.unsupported = true;
        }
    }
}

typedef std::unordered_map<std::string, std::vector<Param>> ParamGroups;

// Tag scan of [begin, end).  Spectra found are appended to `out` with ordinals from `first_ordinal`; param-group
// definitions go to `define` (head of the file) and are looked up in `groups`.  With `stop_at_spectrum` the scan
// ends at the first <spectrum> tag and reports its position.
int scan_range(const char* begin, const char* end, const ParamGroups& groups, ParamGroups* define, int64_t first_ordinal,
               std::vector<SpectrumInfo>& out, bool stop_at_spectrum, const char** stopped_at) {
    const char* p = begin;
    std::vector<Param>* open_group = nullptr;
    bool in_spectrum = false;           // `out.back()` is open
    ArrayInfo array;
    ArrayKind kind = KIND_NONE;
    bool in_array = false;
    int64_t ordinal = first_ordinal;
    if (stopped_at) *stopped_at = nullptr;
    while (p < end) {
        const char* lt = (const char*)memchr(p, '<', (size_t)(end - p));
        if (!lt || lt + 1 >= end) break;
        const char* name = lt + 1;
        if (*name == '!' || *name == '?') {                // comments, declarations, processing instructions
            const char* gt = name[0] == '!' && name + 2 < end && name[1] == '-' && name[2] == '-'
                                 ? FIND_LIT(name, (size_t)(end - name), "-->")
                                 : (const char*)memchr(name, '>', (size_t)(end - name));
            if (!gt) break;
            p = gt + (*gt == '-' ? 3 : 1);
            continue;
        }
        const bool closing = *name == '/';
        if (closing) ++name;
        const char* q = name;
        while (q < end && *q != ' ' && *q != '>' && *q != '/' && *q != '\t' && *q != '\n' && *q != '\r') ++q;
        // end of the tag, quotes respected ('>' may stand inside an attribute value)
        const char* gt = q;
        char quote = 0;
        while (gt < end && (quote || *gt != '>')) {
            if (quote) {
                if (*gt == quote) quote = 0;
            } else if (*gt == '"' || *gt == '\'') {
                quote = *gt;
            }
            ++gt;
        }
        if (gt >= end) break;
        Span tag{name, q};
        const char* colon = (const char*)memchr(tag.a, ':', tag.size());     // namespace prefix
        if (colon) tag.a = colon + 1;
        const bool self_closing = gt > lt && gt[-1] == '/';
        SpectrumInfo* spectrum = in_spectrum ? &out.back() : nullptr;
        if (!closing) {
            if (tag.equals("cvParam")) {
                Param param{attribute(q, gt, "accession"), attribute(q, gt, "value"), attribute(q, gt, "unitName")};
                if (open_group) open_group->push_back(param);
                else if (spectrum) apply_param(param, *spectrum, in_array ? &array : nullptr, &kind);
            } else if (tag.equals("referenceableParamGroupRef")) {
                if (spectrum) {
                    auto it = groups.find(attribute(q, gt, "ref").str());
                    if (it != groups.end())
                        for (const Param& param : it->second) apply_param(param, *spectrum, in_array ? &array : nullptr, &kind);
                }
            } else if (tag.equals("referenceableParamGroup")) {
                open_group = (self_closing || !define) ? nullptr : &(*define)[attribute(q, gt, "id").str()];
            } else if (tag.equals("spectrum")) {
                if (stop_at_spectrum) {
                    if (stopped_at) *stopped_at = lt;
                    return QMS_MZML_OK;
                }
                out.emplace_back();
                spectrum = &out.back();
                spectrum->ordinal = ordinal++;
                spectrum->native_id = attribute(q, gt, "id");
                spectrum->default_length = to_int(attribute(q, gt, "defaultArrayLength"), -1);
                in_spectrum = !self_closing;
            } else if (tag.equals("binaryDataArray") && spectrum) {
                array = ArrayInfo{};
                array.length = to_int(attribute(q, gt, "arrayLength"), -1);
                kind = KIND_NONE;
                in_array = !self_closing;
            } else if (tag.equals("binary") && spectrum && in_array) {
                array.present = true;
                if (self_closing) {
                    array.text = Span{gt, gt};
                } else {
                    const char* close = FIND_LIT(gt + 1, (size_t)(end - gt - 1), "</");
                    if (!close) return fail(QMS_MZML_E_FORMAT, "unterminated <binary> element");
                    array.text = Span{gt + 1, close};
                    p = close;
                    continue;
                }
            }
        } else {
            if (tag.equals("binaryDataArray") && spectrum && in_array) {
                if (kind == KIND_MZ) spectrum->mz = array;
                else if (kind == KIND_INTENSITY) spectrum->intensity = array;
                in_array = false;
            } else if (tag.equals("spectrum")) {
                in_spectrum = false;
                in_array = false;
            } else if (tag.equals("referenceableParamGroup")) {
                open_group = nullptr;
            }
        }
        p = gt + 1;
    }
    return QMS_MZML_OK;
}

// byte offsets of the <spectrum> elements from the index of an indexedmzML file; empty if there is no usable index
std::vector<size_t> indexed_offsets(const qms_mzml* f) {
    std::vector<size_t> offsets;
    const size_t tail = std::min<size_t>(f->size, 4096);
    const char* tag = FIND_LIT(f->data + f->size - tail, tail, "<indexListOffset>");
    if (!tag) return offsets;
    const int64_t list_at = to_int(Span{tag + 17, std::min(tag + 17 + 24, f->data + f->size)}, -1);
    if (list_at <= 0 || (size_t)list_at >= f->size || memcmp(f->data + list_at, "<indexList", 10) != 0) return offsets;
    const char* p = f->data + list_at;
    const char* end = f->data + f->size;
    // the <index> element whose name attribute is "spectrum" (attribute order and quote character are the writer's choice)
    const char* index = nullptr;
    for (const char* q = p; (q = FIND_LIT(q, (size_t)(end - q), "<index")) != nullptr; q += 6) {
        if (q[6] != ' ' && q[6] != '\t' && q[6] != '\n' && q[6] != '\r') continue;         // <indexList, <indexListOffset
        const char* gt = (const char*)memchr(q, '>', (size_t)(end - q));
        if (!gt) break;
        const Span name = attribute(q + 6, gt, "name");
        if (name.size() == 8 && memcmp(name.a, "spectrum", 8) == 0) {
            index = gt;
            break;
        }
    }
    if (!index) return offsets;
    const char* index_end = FIND_LIT(index, (size_t)(end - index), "</index>");
    if (!index_end) return offsets;
    for (const char* o = index; (o = FIND_LIT(o, (size_t)(index_end - o), "<offset")) != nullptr;) {
        const char* gt = (const char*)memchr(o, '>', (size_t)(index_end - o));
        if (!gt) break;
        const int64_t at = to_int(Span{gt + 1, std::min(gt + 1 + 24, index_end)}, -1);
        // every offset must point at a <spectrum> tag in ascending order, or the index is not trusted at all
        if (at <= 0 || (size_t)at + 10 >= (size_t)list_at || memcmp(f->data + at, "<spectrum", 9) != 0 ||
            (!offsets.empty() && (size_t)at <= offsets.back()))
            return std::vector<size_t>();
        offsets.push_back((size_t)at);
        o = gt;
    }
    return offsets;
}

int build_index(qms_mzml* f, int n_threads) {
    const char* begin = f->data;
    const char* end = f->data + f->size;
    ParamGroups groups;
    const std::vector<size_t> offsets = indexed_offsets(f);
    int threads = n_threads > 0 ? n_threads : (int)(std::thread::hardware_concurrency() / 2);
    threads = std::max(1, std::min(threads, 16));
    if (offsets.size() < 64 || threads < 2) {               // no (trusted) index or nothing to gain: one sequential scan
        const ParamGroups none;
        std::vector<SpectrumInfo> head;
        const char* first = nullptr;
        int rc = scan_range(begin, end, none, &groups, 0, head, true, &first);
        if (rc != QMS_MZML_OK || !first) return rc;
        return scan_range(first, end, groups, nullptr, 0, f->spectra, false, nullptr);
    }
    // indexed file: the head (param groups) sequentially, then the spectra in contiguous blocks, one per thread
    {
        const ParamGroups none;
        std::vector<SpectrumInfo> head;
        int rc = scan_range(begin, begin + offsets[0], none, &groups, 0, head, true, nullptr);
        if (rc != QMS_MZML_OK) return rc;
    }
    const char* list_end = FIND_LIT(begin + offsets.back(), (size_t)(end - begin) - offsets.back(), "</spectrumList");
    const char* region_end = list_end ? list_end : end;
    threads = (int)std::min<size_t>((size_t)threads, offsets.size() / 32);
    std::vector<std::vector<SpectrumInfo>> parts((size_t)threads);
    std::vector<int> status((size_t)threads, QMS_MZML_OK);
    std::vector<std::string> messages((size_t)threads);
    auto work = [&](int t) {
        const size_t a = offsets.size() * (size_t)t / (size_t)threads, b = offsets.size() * (size_t)(t + 1) / (size_t)threads;
        const char* stop = b < offsets.size() ? begin + offsets[b] : region_end;
        status[t] = scan_range(begin + offsets[a], stop, groups, nullptr, (int64_t)a, parts[t], false, nullptr);
        if (status[t] != QMS_MZML_OK) messages[t] = g_error;
        if (status[t] == QMS_MZML_OK && parts[t].size() != b - a) {
            status[t] = QMS_MZML_E_FORMAT;
            messages[t] = "the spectrum index of the file does not match its content";
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& t : pool) t.join();
    for (int t = 0; t < threads; ++t)
        if (status[t] != QMS_MZML_OK) {
            if (status[t] == QMS_MZML_E_FORMAT && messages[t].find("index of the file") != std::string::npos) {
                // a wrong index is not an error of the data: fall back to the sequential scan
                f->spectra.clear();
                const ParamGroups none;
                ParamGroups again;
                std::vector<SpectrumInfo> head;
                const char* first = nullptr;
                int rc = scan_range(begin, end, none, &again, 0, head, true, &first);
                if (rc != QMS_MZML_OK || !first) return rc;
                return scan_range(first, end, again, nullptr, 0, f->spectra, false, nullptr);
            }
            return fail(status[t], "%s", messages[t].c_str());
        }
    for (auto& part : parts) f->spectra.insert(f->spectra.end(), part.begin(), part.end());
    f->scan_blocks = threads;
    return QMS_MZML_OK;
}

// ---- decoding --------------------------------------------------------------------------------------------------
struct Base64Table {
    int8_t value[256];
    Base64Table() {
        memset(value, -1, sizeof(value));
        const char* alphabet = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
        for (int i = 0; i < 64; ++i) value[(unsigned char)alphabet[i]] = (int8_t)i;
    }
};
const Base64Table g_base64;

// characters outside the alphabet (white space, padding) are skipped, like Python's base64.b64decode
void base64_decode(Span text, std::vector<unsigned char>& out) {
    out.resize(text.size() / 4 * 3 + 3);
    unsigned char* o = out.data();
    const unsigned char* p = (const unsigned char*)text.a;
    const unsigned char* end = (const unsigned char*)text.b;
    uint32_t acc = 0;
    int bits = 0;
    while (p < end) {
        // fast path: whole quartets of alphabet characters (the usual payload has no white space inside)
        if (bits == 0) {
            while (p + 4 <= end) {
                const int a = g_base64.value[p[0]], b = g_base64.value[p[1]], c = g_base64.value[p[2]], d = g_base64.value[p[3]];
                if ((a | b | c | d) < 0) break;
                const uint32_t v = ((uint32_t)a << 18) | ((uint32_t)b << 12) | ((uint32_t)c << 6) | (uint32_t)d;
                o[0] = (unsigned char)(v >> 16);
                o[1] = (unsigned char)(v >> 8);
                o[2] = (unsigned char)v;
                o += 3;
                p += 4;
            }
            if (p >= end) break;
        }
        const int v = g_base64.value[*p++];
        if (v < 0) continue;
        acc = (acc << 6) | (uint32_t)v;
        bits += 6;
        if (bits >= 8) {
            bits -= 8;
            *o++ = (unsigned char)(acc >> bits);
        }
    }
    out.resize((size_t)(o - out.data()));
}

struct Inflater {               // one zlib stream per thread, reset between arrays
    z_stream z;
    bool ready = false;
    ~Inflater() {
        if (ready) inflateEnd(&z);
    }
    bool run(const std::vector<unsigned char>& in, std::vector<unsigned char>& out, size_t expected) {
        if (in.empty()) {
            out.clear();
            return true;
        }
        if (!ready) {
            memset(&z, 0, sizeof(z));
            if (inflateInit(&z) != Z_OK) return false;
            ready = true;
        } else if (inflateReset(&z) != Z_OK) {
            return false;
        }
        out.resize(std::max<size_t>(expected, in.size() * 4) + 64);
        z.next_in = const_cast<unsigned char*>(in.data());
        z.avail_in = (uInt)in.size();
        for (;;) {
            z.next_out = out.data() + z.total_out;
            z.avail_out = (uInt)(out.size() - z.total_out);
            const int rc = inflate(&z, Z_NO_FLUSH);
            if (rc == Z_STREAM_END) break;
            if (rc != Z_OK && rc != Z_BUF_ERROR) return false;
            if (z.avail_out == 0) out.resize(out.size() * 2);
            else if (z.avail_in == 0) return false;        // truncated stream
        }
        out.resize(z.total_out);
        return true;
    }
};

struct Scratch {
    Inflater inflater;
    std::vector<unsigned char> raw, plain;
    std::vector<double> mz, intensity;
    std::vector<int64_t> order;
};

// ---- MS-Numpress (Teleman et al. 2014; the rules are spelled out in pyqms_b200/numpress.py) ----
struct HalfBytes {
    const unsigned char* data;
    size_t size, at;
    bool low = false;
    bool at_end() const { return at >= size || (at == size - 1 && low && (data[at] & 0xF) == 0); }
    bool next(unsigned& out) {
        if (at >= size) return false;
        if (low) {
            out = data[at++] & 0xFu;
            low = false;
        } else {
            out = data[at] >> 4;
            low = true;
        }
        return true;
    }
    // one integer: a head half byte (<= 8: that many leading zero half bytes, else head - 8 leading 0xF ones), then the rest,
    // least significant first
    bool integer(int32_t& out) {
        unsigned head;
        if (!next(head)) return false;
        unsigned lead = head <= 8 ? head : head - 8;
        uint32_t value = 0;
        if (head > 8)
            for (unsigned i = 0; i < lead; ++i) value |= 0xF0000000u >> (4 * i);
        for (unsigned i = 0; i + lead < 8; ++i) {
            unsigned hb;
            if (!next(hb)) return false;
            value |= (uint32_t)hb << (4 * i);
        }
        out = (int32_t)value;
        return true;
    }
};

bool numpress_fixed_point(const unsigned char* data, size_t size, double& out) {
    if (size < 8) return false;
    unsigned char raw[8];
    for (int i = 0; i < 8; ++i) raw[i] = data[7 - i];          // stored most significant byte first
    memcpy(&out, raw, 8);
    return true;
}

uint32_t le32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

bool numpress_decode(Packing kind, const unsigned char* data, size_t size, std::vector<double>& out) {
    out.clear();
    if (kind == PACK_NUMPRESS_PIC) {
        HalfBytes h{data, size, 0};
        while (!h.at_end()) {
            int32_t v;
            if (!h.integer(v)) return false;
            out.push_back((double)(uint32_t)v);
        }
        return true;
    }
    double fixed_point;
    if (!numpress_fixed_point(data, size, fixed_point)) return false;
    if (kind == PACK_NUMPRESS_SLOF) {
        if ((size - 8) % 2) return false;
        out.reserve((size - 8) / 2);
        for (size_t i = 8; i + 1 < size; i += 2) out.push_back(exp((double)((unsigned)data[i] | ((unsigned)data[i + 1] << 8)) / fixed_point) - 1.0);
        return true;
    }
    if (size == 8) return true;                                // linear prediction
    if (size < 12) return false;
    int64_t before = (int64_t)le32(data + 8);
    out.push_back((double)before / fixed_point);
    if (size == 12) return true;
    if (size < 16) return false;
    int64_t last = (int64_t)le32(data + 12);
    out.push_back((double)last / fixed_point);
    HalfBytes h{data, size, 16};
    while (!h.at_end()) {
        int32_t residual;
        if (!h.integer(residual)) return false;
        const int64_t value = last + (last - before) + (int64_t)residual;
        out.push_back((double)value / fixed_point);
        before = last;
        last = value;
    }
    return true;
}

// one binary array -> doubles; false on a broken payload
bool decode_array(const ArrayInfo& a, int64_t expected, Scratch& s, std::vector<double>& out) {
    out.clear();
    if (!a.present) return true;
    base64_decode(a.text, s.raw);
    const std::vector<unsigned char>* bytes = &s.raw;
    if (a.zlib) {
        const size_t guess = expected > 0 ? (size_t)expected * (a.f32 ? 4 : 8) : 0;
        if (!s.inflater.run(s.raw, s.plain, guess)) return false;
        bytes = &s.plain;
    }
    if (a.packing >= PACK_NUMPRESS_LINEAR) return numpress_decode(a.packing, bytes->data(), bytes->size(), out);
    if (a.packing == PACK_INT32 || a.packing == PACK_INT64) {
        const size_t wide = a.packing == PACK_INT32 ? 4 : 8;
        out.resize(bytes->size() / wide);
        for (size_t i = 0; i < out.size(); ++i) {
            if (wide == 4) {
                int32_t v;
                memcpy(&v, bytes->data() + 4 * i, 4);
                out[i] = (double)v;
            } else {
                int64_t v;
                memcpy(&v, bytes->data() + 8 * i, 8);
                out[i] = (double)v;
            }
        }
        return true;
    }
    const size_t width = a.f32 ? 4 : 8;
    const size_t n = bytes->size() / width;
    out.resize(n);
    if (a.f32) {
        for (size_t i = 0; i < n; ++i) {
            float v;
            memcpy(&v, bytes->data() + 4 * i, 4);
            out[i] = (double)v;
        }
    } else if (n) {
        memcpy(out.data(), bytes->data(), 8 * n);
    }
    return true;
}

int64_t expected_length(const SpectrumInfo& s, const ArrayInfo& a) { return a.length >= 0 ? a.length : s.default_length; }

// rows of a spectrum before decoding: exact for uncompressed arrays, the declared length for compressed ones
int64_t planned_rows(const SpectrumInfo& s) {
    if (!s.mz.present) return 0;
    if (!s.mz.zlib && s.mz.packing < PACK_NUMPRESS_LINEAR) {
        size_t symbols = 0;
        for (const char* p = s.mz.text.a; p < s.mz.text.b; ++p) symbols += g_base64.value[(unsigned char)*p] >= 0;
        const size_t width = s.mz.packing == PACK_INT32 ? 4 : s.mz.packing == PACK_INT64 ? 8 : s.mz.f32 ? 4 : 8;
        return (int64_t)(symbols * 6 / 8 / width);
    }
    const int64_t n = expected_length(s, s.mz);
    return n < 0 ? 0 : n;
}

// pymzml-style id: the number behind the first "scan=", "index=" or "spectrum=" of the native id, else ordinal + 1
int64_t spectrum_id(const SpectrumInfo& s) {
    static const char* keys[3] = {"scan=", "index=", "spectrum="};
    for (const char* p = s.native_id.a; p < s.native_id.b; ++p) {
        for (const char* key : keys) {
            const size_t n = strlen(key);
            if ((size_t)(s.native_id.b - p) > n && memcmp(p, key, n) == 0 && p[n] >= '0' && p[n] <= '9') {
                int64_t v = 0;
                for (const char* d = p + n; d < s.native_id.b && *d >= '0' && *d <= '9'; ++d) v = v * 10 + (*d - '0');
                return v;
            }
        }
    }
    return s.ordinal + 1;
}

bool selected(const SpectrumInfo& s, int32_t ms_level) { return ms_level <= 0 || s.level == ms_level; }

}  // namespace

extern "C" {

const char* qms_mzml_last_error(void) { return g_error.c_str(); }

int qms_mzml_open(const char* path, qms_mzml** out) {
    if (!path || !out) return fail(QMS_MZML_E_ARG, "qms_mzml_open: NULL argument");
    *out = nullptr;
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(QMS_MZML_E_IO, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) {
        close(fd);
        return fail(QMS_MZML_E_IO, "cannot stat %s (or the file is empty)", path);
    }
    void* map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) {
        close(fd);
        return fail(QMS_MZML_E_IO, "cannot map %s", path);
    }
    madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL);
    qms_mzml* f = new qms_mzml();
    f->fd = fd;
    f->data = (const char*)map;
    f->size = (size_t)st.st_size;
    f->mapped = true;
    if (f->size > 2 && (unsigned char)f->data[0] == 0x1f && (unsigned char)f->data[1] == 0x8b) {
        // gzip-compressed file (.mzML.gz): inflate it into memory once, index and decode from there
        z_stream z;
        memset(&z, 0, sizeof(z));
        if (inflateInit2(&z, 16 + MAX_WBITS) != Z_OK) {
            qms_mzml_close(f);
            return fail(QMS_MZML_E_IO, "zlib initialisation failed");
        }
        f->inflated.resize(f->size * 4 + (1u << 20));
        z.next_in = (Bytef*)f->data;
        size_t fed = 0;
        int zrc = Z_OK;
        while (zrc != Z_STREAM_END) {
            if (z.avail_in == 0 && fed < f->size) {
                const size_t n = std::min<size_t>(f->size - fed, 1u << 30);
                z.next_in = (Bytef*)(f->data + fed);
                z.avail_in = (uInt)n;
                fed += n;
            }
            if (z.total_out == f->inflated.size()) f->inflated.resize(f->inflated.size() * 2);
            z.next_out = (Bytef*)f->inflated.data() + z.total_out;
            z.avail_out = (uInt)std::min<size_t>(f->inflated.size() - z.total_out, 1u << 30);
            zrc = inflate(&z, Z_NO_FLUSH);
            if (zrc == Z_STREAM_END && (z.avail_in > 0 || fed < f->size)) {
                // another gzip member follows (bgzip, pigz, concatenated files): gzip.open reads them all, so do we
                const uLong so_far = z.total_out;
                if (inflateReset(&z) != Z_OK) break;
                z.total_out = so_far;
                zrc = Z_OK;
                continue;
            }
            if (zrc != Z_OK && zrc != Z_STREAM_END && zrc != Z_BUF_ERROR) break;
            if (zrc == Z_BUF_ERROR && z.avail_in == 0 && fed >= f->size) break;      // truncated
        }
        const size_t total = z.total_out;
        inflateEnd(&z);
        if (zrc != Z_STREAM_END) {
            qms_mzml_close(f);
            return fail(QMS_MZML_E_FORMAT, "%s: broken gzip stream", path);
        }
        f->inflated.resize(total);
        munmap((void*)f->data, f->size);
        f->mapped = false;
        f->data = f->inflated.data();
        f->size = total;
    }
    const int rc = build_index(f, 0);
    if (rc != QMS_MZML_OK) {
        qms_mzml_close(f);
        return rc;
    }
    *out = f;
    return QMS_MZML_OK;
}

void qms_mzml_close(qms_mzml* f) {
    if (!f) return;
    if (f->data && f->mapped) munmap((void*)f->data, f->size);
    if (f->fd >= 0) close(f->fd);
    delete f;
}

int qms_mzml_scan_blocks(qms_mzml* file) { return file ? file->scan_blocks : 0; }

int qms_numpress_decode(int32_t kind, const unsigned char* data, int64_t size, double* out, int64_t capacity, int64_t* count) {
    if (kind < 0 || kind > 2 || size < 0 || (size > 0 && !data) || !count) return fail(QMS_MZML_E_ARG, "qms_numpress_decode: bad argument");
    static thread_local std::vector<double> values;
    if (!numpress_decode((Packing)(PACK_NUMPRESS_LINEAR + kind), data, (size_t)size, values))
        return fail(QMS_MZML_E_FORMAT, "numpress stream is cut short");
    *count = (int64_t)values.size();
    if (!out || capacity < *count) return out ? fail(QMS_MZML_E_ARG, "qms_numpress_decode: room for %lld values, the stream holds %lld",
                                                     (long long)capacity, (long long)*count) : QMS_MZML_OK;
    if (!values.empty()) memcpy(out, values.data(), values.size() * sizeof(double));
    return QMS_MZML_OK;
}

int qms_mzml_count(qms_mzml* f, int32_t ms_level, int64_t* n_spectra, int64_t* n_peaks) {
    if (!f || !n_spectra || !n_peaks) return fail(QMS_MZML_E_ARG, "qms_mzml_count: NULL argument");
    int64_t spectra = 0, peaks = 0;
    for (const SpectrumInfo& s : f->spectra) {
        if (!selected(s, ms_level)) continue;
        if (s.unsupported) return fail(QMS_MZML_E_UNSUPPORTED, "spectrum %lld uses an array encoding this reader does not decode", (long long)s.ordinal);
        if (!s.mz.present || !s.intensity.present)
            return fail(QMS_MZML_E_FORMAT, "spectrum %lld: m/z and intensity arrays missing", (long long)s.ordinal);
        ++spectra;
        peaks += planned_rows(s);
    }
    *n_spectra = spectra;
    *n_peaks = peaks;
    return QMS_MZML_OK;
}

int qms_mzml_read(qms_mzml* f, int32_t ms_level, int32_t sort_peaks, int32_t n_threads, double* peaks, int64_t* offsets, int64_t* ids,
                  double* rt, uint8_t* rt_unit, int32_t* levels) {
    if (!f || !offsets) return fail(QMS_MZML_E_ARG, "qms_mzml_read: NULL argument");
    std::vector<const SpectrumInfo*> chosen;
    for (const SpectrumInfo& s : f->spectra)
        if (selected(s, ms_level)) chosen.push_back(&s);
    offsets[0] = 0;
    for (size_t i = 0; i < chosen.size(); ++i) {
        offsets[i + 1] = offsets[i] + planned_rows(*chosen[i]);
        if (ids) ids[i] = spectrum_id(*chosen[i]);
        if (rt) rt[i] = chosen[i]->has_rt ? chosen[i]->rt : 0.0;
        if (rt_unit) rt_unit[i] = chosen[i]->has_rt ? (chosen[i]->rt_seconds ? 2 : 1) : 0;
        if (levels) levels[i] = chosen[i]->level;
    }
    if (chosen.empty()) return QMS_MZML_OK;
    if (!peaks && offsets[chosen.size()] > 0) return fail(QMS_MZML_E_ARG, "qms_mzml_read: peaks is NULL");
    int threads = n_threads > 0 ? n_threads : (int)(std::thread::hardware_concurrency() / 2);
    threads = std::max(1, std::min(threads, 16));
    threads = (int)std::min<size_t>((size_t)threads, chosen.size());
    std::atomic<size_t> next{0};
    std::atomic<int> status{QMS_MZML_OK};
    std::atomic<long long> broken{-1};
    auto worker = [&]() {
        Scratch s;
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= chosen.size() || status.load() != QMS_MZML_OK) return;
            const SpectrumInfo& sp = *chosen[i];
            const int64_t rows = offsets[i + 1] - offsets[i];
            if (!decode_array(sp.mz, expected_length(sp, sp.mz), s, s.mz) ||
                !decode_array(sp.intensity, expected_length(sp, sp.intensity), s, s.intensity) ||
                (int64_t)s.mz.size() != rows || s.intensity.size() != s.mz.size()) {
                broken.store((long long)sp.ordinal);
                status.store(QMS_MZML_E_FORMAT);
                return;
            }
            double* out = peaks + 2 * offsets[i];
            bool ascending = true;
            for (int64_t k = 1; k < rows && ascending; ++k) ascending = !(s.mz[k] < s.mz[k - 1]);
            if (sort_peaks && !ascending) {
                s.order.resize((size_t)rows);
                std::iota(s.order.begin(), s.order.end(), (int64_t)0);
                std::stable_sort(s.order.begin(), s.order.end(), [&](int64_t a, int64_t b) { return s.mz[a] < s.mz[b]; });
                for (int64_t k = 0; k < rows; ++k) {
                    out[2 * k] = s.mz[s.order[k]];
                    out[2 * k + 1] = s.intensity[s.order[k]];
                }
            } else {
                for (int64_t k = 0; k < rows; ++k) {
                    out[2 * k] = s.mz[k];
                    out[2 * k + 1] = s.intensity[k];
                }
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    if (status.load() != QMS_MZML_OK)
        return fail(status.load(), "spectrum %lld: binary arrays are broken or do not have the declared length", broken.load());
    return QMS_MZML_OK;
}

}  // extern "C"
