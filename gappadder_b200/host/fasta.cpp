#include "fasta.hpp"

#include <cctype>
#include <cstdio>

namespace gpm {

bool read_fasta(const std::string& path, std::vector<FastaRecord>& out, std::string& fatal)
{
    out.clear();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return true;                                   // ifstream on a missing file: no records
    std::string data;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, got);
    fclose(f);
    bool have = false;
    size_t i = 0;
    const size_t n = data.size();
    while (i < n) {
        const unsigned char c = (unsigned char)data[i++];
        if (c == '>') {                                    // fastareader.cpp:196-210
            out.emplace_back();
            have = true;
            while (i < n && data[i] != '\n') {
                if (data[i] != '\r') out.back().name += data[i];
                ++i;
            }
            if (i < n) ++i;                                // the '\n'
            continue;
        }
        if (std::isspace(c)) continue;                     // :212
        if (!std::isalpha(c)) {                            // :213-218
            fatal = std::string("Bad char in sequence ") + (char)c;
            return false;
        }
        if (!have) { fatal = "header missing"; return false; }   // :219-222
        out.back().seq += (char)std::toupper(c);           // :225 (ReadFromFile always upper-cases)
    }
    return true;
}

void append_fasta(std::string& out, const std::string& name, const std::string& seq, int width)
{
    out += '>';
    out += name;
    const int n = (int)seq.size();
    for (int i = 0; i < n; ++i) {
        // `i % lineLength == 0` (fastareader.cpp:70); a zero/negative -l would divide by zero in the
        // reference, here it simply never breaks the line.
        if (width > 0 && i % width == 0) out += '\n';
        out += seq[i];
    }
    out += '\n';
}

} // namespace gpm
