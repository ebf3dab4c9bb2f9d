// fasta.hpp -- FASTA input/output with the reference's exact behaviour
// (ContigsCompactor-v0.2.0/ContigsMerger/fastareader.cpp:185-229 reader, :65-74 printer,
// fastaMultiSeqs.cpp:46-73 file loop).
#pragma once
#include <string>
#include <vector>

namespace gpm {

struct FastaRecord {
    std::string name;   // whole header line after '>' ('\r' dropped)
    std::string seq;    // upper-cased letters
};

// Reads every record of `path`.  Returns false and sets `fatal` (the text the reference prints after
// "FATAL ERROR: " before exit(1)) on a non-letter in a sequence or sequence data before any header.
// A file that cannot be opened yields zero records, as in the reference.
bool read_fasta(const std::string& path, std::vector<FastaRecord>& out, std::string& fatal);

// ">" name, then the sequence `width` letters per line, each line started by '\n', closed by '\n'
// (FastaSequence::printFasta).
void append_fasta(std::string& out, const std::string& name, const std::string& seq, int width);

} // namespace gpm
