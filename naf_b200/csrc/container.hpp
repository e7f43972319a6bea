// container.hpp — host side of the .naf container: header, VLE numbers, section table.
// Replaces unnaf/src/input.c:31 read_header, unnaf/src/utils.c:117 read_number and, on the encode
// side, ennaf/src/ennaf.c:538-589 (header + sections) and ennaf/src/encoders.c:175
// write_variable_length_encoded_number.  Error strings are the reference's die() messages.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace nafc {

struct Section { uint64_t orig = 0, comp = 0, off = 0; bool present = false; };   // off: file offset of the compressed bytes

struct Header {
    int version = 1, seq_type = 0;
    bool has_title = false, has_ids = false, has_names = false, has_lengths = false, has_mask = false, has_data = false, has_quality = false;
    uint8_t sep = ' ';
    uint64_t line_length = 0, n_sequences = 0;
    uint64_t title_off = 0, title_len = 0;
    Section sec[6];                 // ids, names(comments), lengths, mask, data, quality
};

enum { SEC_IDS = 0, SEC_NAMES = 1, SEC_LEN = 2, SEC_MASK = 3, SEC_DATA = 4, SEC_QUAL = 5 };

inline bool get_vle(const uint8_t *p, size_t n, size_t &pos, uint64_t &v, std::string &err)
{
    uint64_t a = 0;
    if (pos >= n) { err = "incomplete or truncated input\n"; return false; }
    uint8_t c = p[pos++];
    if (c == 128) { err = "invalid input: error parsing variable length encoded number\n"; return false; }
    while (c & 128) {
        if (a & (127ull << 57)) { err = "invalid input: overflow reading a variable length encoded number\n"; return false; }
        a = (a << 7) | (c & 127);
        if (pos >= n) { err = "incomplete or truncated input\n"; return false; }
        c = p[pos++];
    }
    if (a & (127ull << 57)) { err = "invalid input: overflow reading a variable length encoded number\n"; return false; }
    v = (a << 7) | c;
    return true;
}

inline void put_vle(std::vector<uint8_t> &out, uint64_t v)
{
    uint8_t tmp[10]; int n = 0;
    tmp[n++] = (uint8_t)(v & 127); v >>= 7;
    while (v) { tmp[n++] = (uint8_t)(128 | (v & 127)); v >>= 7; }
    while (n) out.push_back(tmp[--n]);
}

// level 0: fixed part only (magic .. separator); level 1: + line length, N, title, section table
inline bool read_header(const uint8_t *p, size_t n, Header &h, bool sections, std::string &err)
{
    static const char *trunc = "incomplete or truncated input\n";
    if (n == 0) { err = "empty input"; return false; }
    if (n < 3) { err = trunc; return false; }
    if (p[0] != 0x01 || p[1] != 0xF9 || p[2] != 0xEC) { err = "not a NAF format\n"; return false; }
    size_t pos = 3;
    if (pos >= n) { err = trunc; return false; }
    h.version = p[pos++];
    if (h.version < 1 || h.version > 2) { err = "unknown version (" + std::to_string(h.version) + ") of NAF format\n"; return false; }
    h.seq_type = 0;
    if (h.version > 1) {
        if (pos >= n) { err = trunc; return false; }
        int t = p[pos++];
        if (t < 1 || t > 3) { err = "unknown sequence type (" + std::to_string(t) + ") found in NAF file\n"; return false; }
        h.seq_type = t;
    }
    if (pos + 2 > n) { err = trunc; return false; }
    int flags = p[pos++];
    h.has_title = (flags >> 6) & 1; h.has_ids = (flags >> 5) & 1; h.has_names = (flags >> 4) & 1; h.has_lengths = (flags >> 3) & 1;
    h.has_mask = (flags >> 2) & 1; h.has_data = (flags >> 1) & 1; h.has_quality = flags & 1;
    h.sep = p[pos++];
    if (h.sep < 0x20 || h.sep > 0x7E) { err = "unsupported name separator character\n"; return false; }
    if (!sections) return true;
    if (!get_vle(p, n, pos, h.line_length, err)) return false;
    if (!get_vle(p, n, pos, h.n_sequences, err)) return false;
    if (h.has_title) {
        if (!get_vle(p, n, pos, h.title_len, err)) return false;
        if (h.title_len > n - pos) { err = trunc; return false; }
        h.title_off = pos; pos += h.title_len;
    }
    const bool present[6] = { h.has_ids, h.has_names, h.has_lengths, h.has_mask, h.has_data, h.has_quality };
    for (int k = 0; k < 6; k++) {
        if (!present[k]) continue;
        if (h.n_sequences == 0 && pos >= n) break;
        if (!get_vle(p, n, pos, h.sec[k].orig, err)) return false;
        if (!get_vle(p, n, pos, h.sec[k].comp, err)) return false;
        if (h.sec[k].comp > n - pos) { err = trunc; return false; }
        h.sec[k].off = pos; h.sec[k].present = true; pos += h.sec[k].comp;
    }
    return true;
}

}  // namespace nafc
