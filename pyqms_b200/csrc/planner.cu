// Host-side planner: molecule strings -> element-count rows and Hill formulas (no GPU work).
//
// The library build starts from molecule strings (IL:314-347 hands every molecule to the chemical_composition
// package, one Python call per molecule).  At 10^5..10^6 molecules that string loop, not the device build, bounds the
// end-to-end envelopes/s, so the two common molecule kinds are parsed here in one pass over a packed text buffer:
//
//   "+C104H175N29O33S1", "+C(2)H(3)14N(1)O(1)", "+C(-6)13C(6)"      plain formulas (isotope prefix, signed counts)
//   "PEPTIDEK", "PEPTIDEK+PO3", "AC0DEFC0K"                          residues (letter + optional fixed-label index,
//                                                                    IL:1375-1677) + H2O (+ add-on formula)
//
// Anything else (modifications '#', unknown symbols or residues, odd syntax) is flagged
// and left to the host parser, which is the single authority for those.  Besides the counts the planner reports the
// order in which the symbols of a molecule were first touched: the envelope of a formula is the convolution of its
// element terms in exactly that order (IL:546-646), and the result bits depend on it.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "qms_internal.cuh"

namespace {

struct SymbolTable {
    std::unordered_map<std::string, int> index;
    int16_t plain[26][27];          // "C", "Cl": capital (+ one small letter) -> symbol, without a string or a hash
    explicit SymbolTable(int32_t n, const char* blob, const int32_t* off) {
        for (auto& row : plain)
            for (auto& slot : row) slot = -1;
        for (int32_t s = 0; s < n; ++s) {
            const char* name = blob + off[s];
            const int32_t len = off[s + 1] - off[s];
            index.emplace(std::string(name, name + len), s);
            if (s < 32767 && len >= 1 && len <= 2 && name[0] >= 'A' && name[0] <= 'Z' && (len == 1 || (name[1] >= 'a' && name[1] <= 'z')))
                plain[name[0] - 'A'][len == 1 ? 0 : 1 + (name[1] - 'a')] = (int16_t)s;
        }
    }
    int find(const char* a, const char* b) const {
        const size_t len = (size_t)(b - a);
        if (len >= 1 && len <= 2 && a[0] >= 'A' && a[0] <= 'Z' && (len == 1 || (a[1] >= 'a' && a[1] <= 'z')))
            return plain[a[0] - 'A'][len == 1 ? 0 : 1 + (a[1] - 'a')];
        auto it = index.find(std::string(a, b));
        return it == index.end() ? -1 : it->second;
    }
};

struct Row {
    int32_t* counts;
    int16_t* seen;
    int16_t next = 0;
    void add(int symbol, int64_t n) {
        if (seen[symbol] < 0) seen[symbol] = next++;      // a symbol keeps the slot of its first addition, even at count 0
        counts[symbol] += (int32_t)n;
    }
};

inline bool is_upper(char c) { return c >= 'A' && c <= 'Z'; }
inline bool is_lower(char c) { return c >= 'a' && c <= 'z'; }
inline bool is_digit(char c) { return c >= '0' && c <= '9'; }

// <isotope digits><Element>( "(" signed int ")" | digits* ) ... ; spaces are skipped.  false = not understood.
bool parse_formula(const char* p, const char* end, const SymbolTable& symbols, Row& row) {
    std::string key;
    while (p < end) {
        if (*p == ' ') {
            ++p;
            continue;
        }
        key.clear();
        while (p < end && (is_digit(*p) || *p == ' ')) {
            if (*p != ' ') key.push_back(*p);
            ++p;
        }
        if (p >= end || !is_upper(*p)) return false;
        key.push_back(*p++);
        while (p < end && is_lower(*p)) key.push_back(*p++);
        int64_t n = 1;
        while (p < end && *p == ' ') ++p;
        if (p < end && *p == '(') {
            ++p;
            bool negative = false;
            if (p < end && *p == '-') {
                negative = true;
                ++p;
            }
            if (p >= end || !is_digit(*p)) return false;
            n = 0;
            while (p < end && is_digit(*p)) {
                n = n * 10 + (*p++ - '0');
                if (n > 100000000) return false;
            }
            if (p >= end || *p != ')') return false;
            ++p;
            if (negative) n = -n;
        } else if (p < end && is_digit(*p)) {
            // bare count; the digits belong to this element even if an isotope-prefixed element follows ("C13C6" is
            // C13 + C6), like the greedy pattern of the host parser
            n = 0;
            while (p < end && (is_digit(*p) || *p == ' ')) {
                if (*p != ' ') {
                    n = n * 10 + (*p - '0');
                    if (n > 100000000) return false;
                }
                ++p;
            }
        }
        const int symbol = symbols.find(key.data(), key.data() + key.size());
        if (symbol < 0) return false;
        row.add(symbol, n);
    }
    return true;
}

}  // namespace

extern "C" {

int qms_parse_molecules(const char* text, const int64_t* str_off, int64_t n, int32_t n_symbols, const char* symbols_blob,
                        const int32_t* symbol_off, int32_t n_residues, const char* residue_blob, const int32_t* residue_off,
                        const int32_t* residue_counts, const int8_t* residue_order, int32_t* counts, int16_t* first_seen,
                        uint8_t* status) {
    if (!text || !str_off || n < 0 || n_symbols <= 0 || n_symbols > 127 || !symbols_blob || !symbol_off || n_residues < 0 ||
        (n_residues > 0 && (!residue_blob || !residue_off || !residue_counts || !residue_order)) || !counts || !first_seen || !status)
        return QMS_E_ARG;
    const SymbolTable symbols(n_symbols, symbols_blob, symbol_off);
    const SymbolTable residues(n_residues, residue_blob, residue_off);     // "A", "C", "C0", "K1", ...
    // letter (+ one index digit) -> residue without a string or a hash per residue; longer names go through the table above
    int16_t quick[26][11];
    for (auto& letter : quick)
        for (auto& slot : letter) slot = -1;
    for (int32_t r = 0; r < n_residues && r < 32767; ++r) {
        const char* name = residue_blob + residue_off[r];
        const int32_t len = residue_off[r + 1] - residue_off[r];
        if (len >= 1 && len <= 2 && is_upper(name[0]) && (len == 1 || is_digit(name[1])))
            quick[name[0] - 'A'][len == 1 ? 0 : 1 + (name[1] - '0')] = (int16_t)r;
    }
    const int h = symbols.find("H", "H" + 1), o = symbols.find("O", "O" + 1);
    for (int64_t m = 0; m < n; ++m) {
        const char* p = text + str_off[m];
        const char* end = text + str_off[m + 1];
        Row row{counts + m * n_symbols, first_seen + m * n_symbols};
        for (int s = 0; s < n_symbols; ++s) {
            row.counts[s] = 0;
            row.seen[s] = -1;
        }
        bool ok = p < end && memchr(p, '#', (size_t)(end - p)) == nullptr;
        if (ok && *p == '+') {
            ok = parse_formula(p + 1, end, symbols, row);
        } else if (ok) {
            const char* plus = (const char*)memchr(p, '+', (size_t)(end - p));
            const char* seq_end = plus ? plus : end;
            ok = seq_end > p && h >= 0 && o >= 0;
            for (const char* q = p; ok && q < seq_end;) {
                const char* name = q;                      // residue letter + optional fixed-label index
                if (!is_upper(*q)) {
                    ok = false;
                    break;
                }
                ++q;
                while (q < seq_end && is_digit(*q)) ++q;
                const int r = q - name <= 2 ? quick[name[0] - 'A'][q - name == 1 ? 0 : 1 + (name[1] - '0')] : residues.find(name, q);
                if (r < 0) {
                    ok = false;
                    break;
                }
                const int8_t* order = residue_order + (size_t)r * n_symbols;
                for (int k = 0; k < n_symbols && order[k] >= 0; ++k)
                    row.add(order[k], residue_counts[(size_t)r * n_symbols + order[k]]);
            }
            if (ok) {
                row.add(h, 2);
                row.add(o, 1);
                if (plus) ok = memchr(plus + 1, '+', (size_t)(end - plus - 1)) == nullptr && parse_formula(plus + 1, end, symbols, row);
            }
        }
        if (ok) {
            for (int s = 0; s < n_symbols; ++s)
                if (row.counts[s] == 0) row.seen[s] = -1;          // zero counts are dropped from a composition
            status[m] = 0;
        } else {
            for (int s = 0; s < n_symbols; ++s) {
                row.counts[s] = 0;
                row.seen[s] = -1;
            }
            status[m] = 1;
        }
    }
    return QMS_OK;
}

int qms_hill_formulas(const int32_t* counts, int64_t n, int32_t n_symbols, const char* symbols_blob, const int32_t* symbol_off,
                      char* out, int64_t out_cap, int64_t* out_off, int64_t* needed) {
    if (!counts || n < 0 || n_symbols <= 0 || !symbols_blob || !symbol_off || !out_off || !needed) return QMS_E_ARG;
    std::vector<std::string> names(n_symbols);
    for (int32_t s = 0; s < n_symbols; ++s) names[s].assign(symbols_blob + symbol_off[s], symbols_blob + symbol_off[s + 1]);
    // Hill order as the reference writes it: C, H, then the remaining symbols in plain string order
    std::vector<int> order;
    for (const char* first : {"C", "H"})
        for (int32_t s = 0; s < n_symbols; ++s)
            if (names[s] == first) order.push_back(s);
    std::vector<int> rest;
    for (int32_t s = 0; s < n_symbols; ++s)
        if (names[s] != "C" && names[s] != "H") rest.push_back(s);
    std::sort(rest.begin(), rest.end(), [&](int a, int b) { return names[a] < names[b]; });
    order.insert(order.end(), rest.begin(), rest.end());
    int64_t used = 0;
    char number[16];
    out_off[0] = 0;
    for (int64_t m = 0; m < n; ++m) {
        for (int s : order) {
            const int32_t c = counts[m * n_symbols + s];
            if (c == 0) continue;
            // "(<c>)" written back to front (a printf call per count was most of the time of a 500 000-formula library)
            char* tail = number + sizeof(number);
            *--tail = ')';
            uint32_t magnitude = c < 0 ? 0u - (uint32_t)c : (uint32_t)c;
            do {
                *--tail = (char)('0' + magnitude % 10u);
                magnitude /= 10u;
            } while (magnitude);
            if (c < 0) *--tail = '-';
            *--tail = '(';
            const int digits = (int)(number + sizeof(number) - tail);
            const int64_t len = (int64_t)names[s].size() + digits;
            if (out && used + len <= out_cap) {
                memcpy(out + used, names[s].data(), names[s].size());
                memcpy(out + used + names[s].size(), tail, (size_t)digits);
            }
            used += len;
        }
        out_off[m + 1] = used;
    }
    *needed = used;
    return (out && used <= out_cap) ? QMS_OK : QMS_E_CAPACITY;
}

}  // extern "C"
