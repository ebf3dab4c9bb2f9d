#include "fasta.hpp"

#include <cctype>
#include <cstdio>
#include <cstring>

namespace gpm {

bool read_fasta(const std::string& path, std::vector<FastaRecord>& out, std::string& fatal)
{
    out.clear();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return true;                                   // ifstream on a missing file: no records
    std::string data;
    if (fseek(f, 0, SEEK_END) == 0) { const long sz = ftell(f); if (sz > 0) data.reserve((size_t)sz); rewind(f); }
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, got);
    fclose(f);
    // Byte classes of the reader loop (fastareader.cpp:196-225): 1 letter (kept, upper-cased), 2 white space (skipped),
    // 0 anything else (fatal).  Sequence lines are handled a line at a time: one pass for the class, one for the copy.
    struct Tables {
        unsigned char cls[256], up[256];
        Tables() { for (int c = 0; c < 256; ++c) { cls[c] = std::isalpha(c) ? 1 : std::isspace(c) ? 2 : 0; up[c] = (unsigned char)std::toupper(c); } }
    };
    static const Tables T;                                 // initialised once, thread-safe
    const unsigned char* cls = T.cls;
    const unsigned char* up = T.up;
    bool have = false;
    size_t i = 0;
    const size_t n = data.size();
    const char* d = data.data();
    while (i < n) {
        if (d[i] == '>') {                                 // fastareader.cpp:196-210
            ++i;
            out.emplace_back();
            have = true;
            const char* nl = (const char*)memchr(d + i, '\n', n - i);
            const size_t e = nl ? (size_t)(nl - d) : n;
            std::string& name = out.back().name;
            name.reserve(e - i);
            for (size_t q = i; q < e; ++q) if (d[q] != '\r') name += d[q];
            i = nl ? e + 1 : n;
            continue;
        }
        // a run of sequence data up to the next '>' (a '>' anywhere starts a header, as in the reference's char loop)
        const char* gt = (const char*)memchr(d + i, '>', n - i);
        const size_t e = gt ? (size_t)(gt - d) : n;
        size_t letters = 0;
        for (size_t q = i; q < e; ++q) {
            const unsigned char k = cls[(unsigned char)d[q]];
            if (k == 1) { if (!have) { fatal = "header missing"; return false; } ++letters; }   // :219-222
            else if (k == 0) { fatal = std::string("Bad char in sequence ") + d[q]; return false; }   // :213-218
        }
        if (letters) {
            std::string& seq = out.back().seq;
            const size_t at = seq.size();
            seq.resize(at + letters);
            char* w = &seq[at];
            for (size_t q = i; q < e; ++q) { const unsigned char c = (unsigned char)d[q]; if (cls[c] == 1) *w++ = (char)up[c]; }   // :225
        }
        i = e;
    }
    return true;
}

void append_fasta(std::string& out, const std::string& name, const std::string& seq, int width)
{
    out += '>';
    out += name;
    // a '\n' before every `width` letters (`i % lineLength == 0`, fastareader.cpp:70), one after the last; a zero/negative
    // -l would divide by zero in the reference, here it simply never breaks the line.
    const size_t n = seq.size();
    if (width <= 0) { out += seq; out += '\n'; return; }
    out.reserve(out.size() + n + n / (size_t)width + 2);
    for (size_t i = 0; i < n; i += (size_t)width) {
        out += '\n';
        out.append(seq, i, (size_t)width);
    }
    out += '\n';
}

} // namespace gpm
