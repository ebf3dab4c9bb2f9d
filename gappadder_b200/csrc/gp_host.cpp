// gp_host.cpp -- host-side half of the C ABI (include/gappadder_b200.h): sequence packing, the
// candidate filter and the exact integer/double epilogue of Evaluate.  No CUDA in this file.
// file:line citations are relative to /root/reference/ContigsCompactor-v0.2.0/ContigsMerger/.
#include "gappadder_b200.h"

#include <algorithm>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

extern "C" {

int gp_abi_version(void) { return GP_ABI_VERSION; }

static inline size_t words_for(uint32_t len)
{
    // eight 4-bit codes per word, rounded up to 4 words (16 bytes) so every sequence can be read
    // with aligned 128-bit loads; always at least one block so empty sequences own an address.
    size_t w = ((size_t)len + 7) / 8;
    w = (w + 3) & ~(size_t)3;
    return w ? w : 4;
}

size_t gp_packed_size(const uint32_t *seq_len, uint32_t n_seq)
{
    size_t words = 0;
    for (uint32_t i = 0; i < n_seq; ++i) words += words_for(seq_len[i]);
    return words * sizeof(uint32_t);
}

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

#if defined(__x86_64__)
// 16 bases -> two packed words, for the bytes A C G T N (codes 0..4: their low nibbles 1 3 7 4 E are distinct, so
// one PSHUFB looks the code up and a second one checks that the byte really is that letter).  Returns false when
// some byte is another character (the caller takes the table-driven path); *has_n is set when an N went through.
__attribute__((target("ssse3"))) inline bool pack16_ssse3(const unsigned char *q, uint32_t *dst, bool *has_n)
{
    const __m128i code_of = _mm_setr_epi8(-1, 0, -1, 1, 3, -1, -1, 2, -1, -1, -1, -1, -1, -1, 4, -1);
    const __m128i byte_of = _mm_setr_epi8(0, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 'N', 0);
    const __m128i v = _mm_loadu_si128((const __m128i *)q);
    const __m128i nib = _mm_and_si128(v, _mm_set1_epi8(0x0f));
    const __m128i codes = _mm_shuffle_epi8(code_of, nib);
    const __m128i expect = _mm_shuffle_epi8(byte_of, nib);
    if (_mm_movemask_epi8(_mm_cmpeq_epi8(expect, v)) != 0xffff) return false;
    if (_mm_movemask_epi8(_mm_cmpeq_epi8(codes, _mm_set1_epi8(4)))) *has_n = true;
    const __m128i pairs = _mm_maddubs_epi16(codes, _mm_set1_epi16(0x1001));        // c0 + 16*c1 per 16-bit lane
    const __m128i bytes = _mm_packus_epi16(pairs, pairs);
    _mm_storel_epi64((__m128i *)dst, bytes);
    return true;
}
#endif

// Packs sequences [s0, s1) with a fixed byte -> code table.  Returns the first sequence index that
// holds a byte the table does not know (code 0xff), or s1 when all went through.  max_code is updated.
uint32_t pack_range(const char *const *seqs, const uint32_t *seq_len, const size_t *word_off, uint32_t s0, uint32_t s1,
                    const uint8_t *code, uint32_t *packed, int *max_code)
{
    uint32_t mx = (uint32_t)(*max_code + 1);             // highest code seen + 1
#if defined(__x86_64__)
    const bool simd = __builtin_cpu_supports("ssse3");
#endif
    for (uint32_t s = s0; s < s1; ++s) {
        const unsigned char *p = (const unsigned char *)seqs[s];
        const uint32_t len = seq_len[s];
        const size_t nw = words_for(len);
        uint32_t *dst = packed + word_off[s];
        const uint32_t full = len / 8;
        uint32_t top = 0;                                // max code + 1 of this sequence; >= 0x100 with an unknown byte
        uint32_t w0 = 0;
#if defined(__x86_64__)
        // plain A/C/G/T/N stretches 16 bases at a time (the table below must still be the initial one for them)
        if (simd && code[(unsigned char)'A'] == 0 && code[(unsigned char)'C'] == 1 && code[(unsigned char)'G'] == 2 &&
            code[(unsigned char)'T'] == 3 && code[(unsigned char)'N'] == 4) {
            bool has_n = false;
            while (w0 + 2 <= full && pack16_ssse3(p + (size_t)w0 * 8, dst + w0, &has_n)) w0 += 2;
            if (w0) top = has_n ? 5 : 4;
        }
#endif
        for (uint32_t w = w0; w < full; ++w) {
            const unsigned char *q = p + (size_t)w * 8;
            uint32_t word = 0;
            for (int k = 0; k < 8; ++k) {
                const uint32_t c = code[q[k]];
                top = std::max(top, c + 1);
                word |= (c & 15u) << (4 * k);
            }
            dst[w] = word;
        }
        uint32_t word = 0;
        for (uint32_t i = full * 8, k = 0; i < len; ++i, ++k) {
            const uint32_t c = code[p[i]];
            top = std::max(top, c + 1);
            word |= (c & 15u) << (4 * k);
        }
        for (size_t w = full; w < nw; ++w) { dst[w] = word; word = 0; }
        if (top > 16) { *max_code = (int)mx - 1; return s; }   // unknown byte (0xff): caller assigns a code and redoes s
        mx = std::max(mx, top);
    }
    *max_code = (int)mx - 1;
    return s1;
}

} // namespace

int gp_pack_sequences(const char *const *seqs, const uint32_t *seq_len, uint32_t n_seq,
                      uint32_t *packed, uint32_t *seq_word_off, uint32_t *n_symbols)
{
    if ((!seqs || !seq_len || !packed || !seq_word_off) && n_seq) return GP_ERR_INVALID;
    std::vector<size_t> off(n_seq);
    size_t total = 0, bases = 0;
    for (uint32_t s = 0; s < n_seq; ++s) {
        if (!seqs[s] && seq_len[s]) return GP_ERR_INVALID;
        off[s] = total;
        seq_word_off[s] = (uint32_t)total;
        total += words_for(seq_len[s]);
        bases += seq_len[s];
        if (total > 0xffffffffull) return GP_ERR_RANGE;
    }
    // byte -> code.  Equality of codes == equality of bytes (ContigsCompactor.cpp:1641).
    uint8_t code[256];
    memset(code, 0xff, sizeof code);
    code[(unsigned char)'A'] = 0; code[(unsigned char)'C'] = 1;
    code[(unsigned char)'G'] = 2; code[(unsigned char)'T'] = 3; code[(unsigned char)'N'] = 4;
    int next = 5, max_code = -1;

    // Large batches of plain A/C/G/T/N input: several host threads, each on a contiguous range of
    // sequences.  A byte outside the table is rare (GAPPadder's contigs never have one); it sends the
    // whole batch through the sequential loop below, which assigns codes in order of first appearance.
    const unsigned hw = std::thread::hardware_concurrency();
    const uint32_t T = bases < (4u << 20) ? 1u : std::min<uint32_t>(hw ? hw : 1u, 16u);
    bool done = false;
    if (T > 1) {
        std::vector<std::thread> th;
        std::vector<uint32_t> stop(T);
        std::vector<int> mx(T, -1);
        std::vector<uint32_t> cut(T + 1, n_seq);
        cut[0] = 0;
        for (uint32_t t = 1, s = 0; t < T; ++t) {      // equal shares of the packed words
            const size_t want = total * t / T;
            while (s < n_seq && off[s] < want) ++s;
            cut[t] = s;
        }
        for (uint32_t t = 0; t < T; ++t)
            th.emplace_back([&, t] { stop[t] = pack_range(seqs, seq_len, off.data(), cut[t], cut[t + 1], code, packed, &mx[t]); });
        for (auto &x : th) x.join();
        done = true;
        for (uint32_t t = 0; t < T; ++t) { done = done && stop[t] == cut[t + 1]; max_code = std::max(max_code, mx[t]); }
        if (!done) max_code = -1;
    }
    if (!done) {
        uint32_t s = 0;
        while (s < n_seq) {
            s = pack_range(seqs, seq_len, off.data(), s, n_seq, code, packed, &max_code);
            if (s == n_seq) break;
            // sequence s holds unknown bytes: give each a code (order of first appearance) and pack it again
            const unsigned char *p = (const unsigned char *)seqs[s];
            for (uint32_t i = 0; i < seq_len[s]; ++i)
                if (code[p[i]] == 0xff) {
                    if (next >= 16) return GP_ERR_ALPHABET;
                    code[p[i]] = (uint8_t)next++;
                }
        }
    }
    // highest code in use + 1: 4 for plain ACGT input, 5 with N, more with other letters
    if (n_symbols) *n_symbols = (uint32_t)(max_code + 1);
    return GP_OK;
}

/* ContigsCompactor::IsScoreSignificant, ContigsCompactor.cpp:1876-1976 (doubles, as there). */
int gp_is_score_significant(const gp_thresholds *t, int32_t scoreMax, int32_t szSeq1, int32_t szSeq2,
                            int32_t rowStart, int32_t colStart, int32_t nclip)
{
    int szOverlap0 = std::min(szSeq1, szSeq2);                              // :1894
    int szOverlap1 = szOverlap0, szOverlap2 = szOverlap0;
    if (rowStart + nclip == szSeq1) szOverlap1 = colStart;                  // :1896
    if (colStart + nclip == szSeq2) szOverlap2 = rowStart;                  // :1900
    int szOverlap = std::min(szOverlap0, std::min(szOverlap1, szOverlap2)); // :1904
    if (szOverlap < szSeq1 * t->frac_min_overlap && szOverlap < szSeq2 * t->frac_min_overlap)
        return 0;                                                           // :1911
    const int MIN_ASM_EXT_LEN = 5;                                          // :1916
    if (rowStart + nclip == szSeq1 && colStart + MIN_ASM_EXT_LEN - 1 >= szSeq2) return 0;  // :1919
    if (colStart + nclip == szSeq2 && rowStart + MIN_ASM_EXT_LEN - 1 >= szSeq1) return 0;  // :1926
    double scoreMinThres = szOverlap * (1 - t->fraction_loss_score);        // :1958
    if (scoreMax < scoreMinThres) return 0;                                 // :1960
    if (szOverlap < t->min_overlap_len_with_scaffold) return 0;             // :1972
    if (szOverlap < t->min_overlap_len) return 1;                           // :1973
    return 2;
}

/* ContigsCompactorAction::IsContainment, ContigsCompactor.cpp:155-159. */
int gp_is_containment(int32_t len1, int32_t len2, const gp_result *r)
{
    const bool bcontained = (r->flags & GP_FLAG_CONTAINED) != 0;
    return bcontained && ((r->row_end + r->nclip == len1 && len1 < r->col_end) ||
                          (r->col_end + r->nclip == len2 && len2 < r->row_end));
}

/* Which branch of SetMergedStringConcat (ContigsCompactor.cpp:108-153) applies:
 * 0: merged = s2 (:116-121)  1: merged = s1 (:122-127)
 * 2: s1[0:len1-nclip] + s2[col_end:] (:131-139)   3: s2[0:len2-nclip] + s1[row_end:] (:141-149) */
static int merged_case(int32_t len1, int32_t len2, const gp_result *r)
{
    const bool bcontained = (r->flags & GP_FLAG_CONTAINED) != 0;
    if (bcontained && r->row_end + r->nclip == len1 && len1 < len2) return 0;
    if (bcontained && r->col_end + r->nclip == len2 && len2 < len1) return 1;
    if (r->row_end + r->nclip == len1) return 2;
    return 3;
}

int32_t gp_merged_length(int32_t len1, int32_t len2, const gp_result *r)
{
    switch (merged_case(len1, len2, r)) {
    case 0: return len2;
    case 1: return len1;
    case 2: return (len1 - r->nclip) + (len2 - r->col_end);
    default: return (len2 - r->nclip) + (len1 - r->row_end);
    }
}

int32_t gp_merged_concat(const char *s1, int32_t len1, const char *s2, int32_t len2,
                         const gp_result *r, char *out)
{
    int32_t n = 0;
    switch (merged_case(len1, len2, r)) {
    case 0: memcpy(out, s2, (size_t)len2); n = len2; break;
    case 1: memcpy(out, s1, (size_t)len1); n = len1; break;
    case 2:
        memcpy(out, s1, (size_t)(len1 - r->nclip)); n = len1 - r->nclip;
        memcpy(out + n, s2 + r->col_end, (size_t)(len2 - r->col_end)); n += len2 - r->col_end;
        break;
    default:
        memcpy(out, s2, (size_t)(len2 - r->nclip)); n = len2 - r->nclip;
        memcpy(out + n, s1 + r->row_end, (size_t)(len1 - r->row_end)); n += len1 - r->row_end;
        break;
    }
    out[n] = 0;
    return n;
}

/* ContigsCompactorAction::GetOverlapSize, ContigsCompactor.h:51. */
int32_t gp_overlap_size(int32_t len1, int32_t len2, const gp_result *r)
{
    return len1 + len2 - r->nclip - gp_merged_length(len1, len2, r);
}

/* ---- dedup stage: the reference's rules, restated (see the header) -------------------------------------------------- */

/* Refiner::gnrtUniqueFa, TERefiner/refiner.cpp:1045-1140: sort (name, index), drop every record whose name equals the
 * previous one's -- i.e. of equally named records the one with the lowest index stays. */
int gp_dedup_unique_names(const char *const *names, uint32_t n_contigs, uint8_t *keep)
{
    if ((!names || !keep) && n_contigs) return GP_ERR_INVALID;
    std::vector<uint32_t> idx(n_contigs);
    for (uint32_t i = 0; i < n_contigs; ++i) { idx[i] = i; keep[i] = 1; }
    std::sort(idx.begin(), idx.end(), [&](uint32_t a, uint32_t b) {                       // cmp_vfa, :472-478
        const int c = strcmp(names[a], names[b]);
        return c != 0 ? c < 0 : a < b;
    });
    for (uint32_t k = 1; k < n_contigs; ++k)
        if (strcmp(names[idx[k]], names[idx[k - 1]]) == 0) keep[idx[k]] = 0;              // :1075-1082
    return GP_OK;
}

/* Alignment::isFullyMapped, TERefiner/Alignment.cpp:397-426 (rlength = the query contig's length). */
static bool dedup_fully_mapped(const gp_dedup_record &a, uint32_t rlength, double cutoff)
{
    if (a.single_m && a.m_len <= rlength) return true;                                   // :400-403
    const int cnt = (int)a.m_len, total_len = (int)a.m_len + (int)a.other_len;           // :408-417
    const double percent = (double)cnt / (double)total_len;                              // :418
    return percent > cutoff;                                                             // :420 (READ_FULL_MAPPED_CUTOFF = -c, main.cpp:165)
}

/* Refiner::removeDupRepeatsOfOneContigSet, TERefiner/refiner.cpp:660-801. */
int gp_dedup_decide(const gp_dedup_record *recs, uint64_t n_recs, const char *const *names, const uint32_t *contig_len,
                    uint32_t n_contigs, double cutoff_ratio, int remove_contained, uint8_t *removed)
{
    if ((!names || !contig_len || !removed) && n_contigs) return GP_ERR_INVALID;
    if (!recs && n_recs) return GP_ERR_INVALID;
    memset(removed, 0, n_contigs);
    for (uint64_t k = 0; k < n_recs; ++k) {
        const gp_dedup_record &a = recs[k];
        if (a.q >= n_contigs || a.r >= n_contigs) return GP_ERR_INVALID;
        const char *qname = names[a.q], *rname = names[a.r];
        if (!remove_contained) {                                                          // :717-766, remove duplicate
            const bool is_fully_map = dedup_fully_mapped(a, contig_len[a.q], cutoff_ratio);
            if (is_fully_map && strcmp(qname, rname) > 0) {                               // qname > rname, std::string order
                const int iq = (int)contig_len[a.q], ir = (int)contig_len[a.r];
                if (iq == ir) removed[a.q] = 1;                                            // :726-737
                else {
                    const int idiff = iq > ir ? iq - ir : ir - iq, imin = iq > ir ? ir : iq;   // :741-752
                    if (((double)idiff / (double)imin) <= (1.0 - cutoff_ratio)) removed[a.q] = 1;   // :754
                }
            }
        } else {                                                                          // :768-786, remove contained
            if (strcmp(qname, rname) == 0) continue;                                      // :770
            if (a.single_m && a.m_len == contig_len[a.q]) removed[a.q] = 1;               // isPerfectMapped, Alignment.cpp:428-437
        }
    }
    return GP_OK;
}

/* Builder-defined record synthesis (see the header): BWA parity unpinned. */
int gp_dedup_records(uint32_t q, uint32_t r, int32_t len_q, int32_t len_r, const gp_result *res, double max_frac_score_loss,
                     gp_dedup_record *out)
{
    if (!res || !out || len_q < 1 || len_r < 1) return 0;
    const int32_t ov = gp_overlap_size(len_q, len_r, res);
    if (ov < 1) return 0;
    if ((double)res->score < (double)ov * (1.0 - max_frac_score_loss)) return 0;
    const uint32_t mq = (uint32_t)std::min(ov, len_q), mr = (uint32_t)std::min(ov, len_r);
    out[0] = gp_dedup_record{q, r, mq == (uint32_t)len_q ? 1u : 0u, mq, (uint32_t)len_q - mq};
    out[1] = gp_dedup_record{r, q, mr == (uint32_t)len_r ? 1u : 0u, mr, (uint32_t)len_r - mr};
    return 2;
}

/* GetComplement, GenSeqsUtils.cpp:24-61; FastaSequence::RevsereComplement, fastareader.cpp. */
void gp_revcomp(const char *s, uint32_t len, char *out)
{
    struct Table {                                         // the rule as a byte table: N/n stay, A<->T, C<->G (either case -> upper), all else N
        char t[256];
        Table()
        {
            for (int b = 0; b < 256; ++b) t[b] = 'N';
            t[(unsigned char)'n'] = 'n';
            t[(unsigned char)'A'] = t[(unsigned char)'a'] = 'T';
            t[(unsigned char)'T'] = t[(unsigned char)'t'] = 'A';
            t[(unsigned char)'G'] = t[(unsigned char)'g'] = 'C';
            t[(unsigned char)'C'] = t[(unsigned char)'c'] = 'G';
        }
    };
    static const Table T;
    const unsigned char *p = (const unsigned char *)s + len;
    for (uint32_t i = 0; i < len; ++i) out[i] = T.t[*--p];
    out[len] = 0;
}

/* 2-bit k-mer letter, SetKmerTypeForNtAt, KmerUtils.cpp:22-58: C=1, G=2, T=3, anything else 0. */
static inline uint64_t kmer_letter(char nt)
{
    switch (nt) {
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': return 3;
    default: return 0;
    }
}

/* Candidate list of the pairwise phase (MultiThreadQuickChecker::threadQuickCheck,
 * ContigsCompactor.cpp:1068-1100, with QuickCheckerContigsMatch :1982-2095): pair (i,j), i <= j, is a
 * candidate iff a k-mer of the first or last 30 bases of node j occurs anywhere in node i.  The
 * reference keeps every k-mer of node i in a std::map and probes it with node j's window k-mers,
 * N(N+1)/2 times; here the test is inverted: the window k-mers of ALL nodes go into one open-addressing
 * table (k-mer -> list of owner nodes j), every node i is scanned once against it and sets hit(i,j) for
 * the owners j >= i of each k-mer it holds.  Same predicate, O(total bases) instead of
 * O(N^2 * 42 * log L).  The reference keeps each k-mer left-aligned in a uint64 (KmerUtils.cpp:60-115);
 * equal windows give equal values, so comparing right-aligned window codes is the same test. */
int64_t gp_candidate_pairs(const char *const *nodes, const uint32_t *node_len, uint32_t n_nodes,
                           int32_t k, gp_pair *pairs, uint64_t cap)
{
    const uint32_t lenContigLen = 30;                                   // :2024
    if (k <= 0 || k > 30 || (!nodes && n_nodes)) return GP_ERR_INVALID;
    if (n_nodes == 0) return 0;
    // Nodes shorter than 30 bases make the reference read outside the string (undefined); here the
    // two windows are clipped to the sequence.  A node shorter than k has no k-mer at all.
    const uint64_t mask = (1ull << (2 * k)) - 1;                        // k <= 30
    // (k-mer, owner) of every window k-mer (:2026-2029), sorted so equal k-mers are adjacent
    struct Probe { uint64_t kmer; uint32_t owner; };
    std::vector<Probe> probes;
    probes.reserve((size_t)n_nodes * 2 * (lenContigLen - (uint32_t)k + 1));
    for (uint32_t j = 0; j < n_nodes; ++j) {
        const uint32_t wlen = node_len[j] < lenContigLen ? node_len[j] : lenContigLen;
        for (int side = 0; side < 2; ++side) {
            const char *w = side == 0 ? nodes[j] : nodes[j] + node_len[j] - wlen;
            uint64_t v = 0;
            for (uint32_t a = 0; a < wlen; ++a) {
                v = ((v << 2) | kmer_letter(w[a])) & mask;
                if (a + 1 >= (uint32_t)k) probes.push_back(Probe{v, j});
            }
        }
    }
    std::sort(probes.begin(), probes.end(), [](const Probe &a, const Probe &b) { return a.kmer != b.kmer ? a.kmer < b.kmer : a.owner < b.owner; });
    // open-addressing table: slot -> [first, last) range of `probes` with that k-mer
    size_t n_unique = 0;
    for (size_t q = 0; q < probes.size(); ++q) if (q == 0 || probes[q].kmer != probes[q - 1].kmer) ++n_unique;
    size_t tcap = 16;
    while (tcap < 2 * n_unique + 2) tcap <<= 1;
    struct Slot { uint64_t kmer; uint32_t first, last; };
    std::vector<Slot> table(tcap, Slot{0, 0, 0});                       // first == last: empty
    auto slot_of = [&](uint64_t v) { return (size_t)((v * 0x9E3779B97F4A7C15ull) >> 20) & (tcap - 1); };
    for (size_t q = 0; q < probes.size();) {
        size_t e = q + 1;
        while (e < probes.size() && probes[e].kmer == probes[q].kmer) ++e;
        size_t h = slot_of(probes[q].kmer);
        while (table[h].first != table[h].last) h = (h + 1) & (tcap - 1);
        table[h] = Slot{probes[q].kmer, (uint32_t)q, (uint32_t)e};
        q = e;
    }
    // hit(i, j): one byte per pair, row i reused for every node
    std::vector<uint8_t> hit(n_nodes);
    int64_t np = 0;
    for (uint32_t i = 0; i < n_nodes; ++i) {
        std::fill(hit.begin() + i, hit.end(), 0);
        const char *s = nodes[i];
        const uint32_t len = node_len[i];
        uint64_t v = 0;
        for (uint32_t a = 0; a < len; ++a) {
            v = ((v << 2) | kmer_letter(s[a])) & mask;
            if (a + 1 < (uint32_t)k) continue;
            size_t h = slot_of(v);
            while (table[h].first != table[h].last) {
                if (table[h].kmer == v) {
                    for (uint32_t q = table[h].first; q < table[h].last; ++q) hit[probes[q].owner] = 1;   // owners < i are never read
                    break;
                }
                h = (h + 1) & (tcap - 1);
            }
        }
        for (uint32_t j = i; j < n_nodes; ++j)                          // i <= j incl. j == i (:1008), row-major like -t 1
            if (hit[j]) {
                if ((uint64_t)np < cap && pairs) { pairs[np].row_seq = i; pairs[np].col_seq = j; }
                ++np;
            }
    }
    return np;
}

uint64_t gp_estimate_gap_cells(const uint32_t *contig_len, uint32_t n_contigs)
{
    // nodes = contigs and their reverse complements: every contig pair appears as 4 node pairs, a
    // contig with itself as 3:  4*sum_{a<b} la*lb + 3*sum la^2 = 2*(sum la)^2 + sum la^2
    uint64_t sum = 0, sq = 0;
    for (uint32_t i = 0; i < n_contigs; ++i) { sum += contig_len[i]; sq += (uint64_t)contig_len[i] * contig_len[i]; }
    return 2 * sum * sum + sq;
}

int gp_partition_gaps(const uint64_t *cost, uint64_t n_gaps, int32_t n_parts, int32_t *part)
{
    if ((!cost || !part) && n_gaps) return GP_ERR_INVALID;
    if (n_parts < 1) return GP_ERR_INVALID;
    std::vector<uint64_t> idx(n_gaps);
    std::iota(idx.begin(), idx.end(), (uint64_t)0);
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return cost[a] > cost[b]; });
    std::vector<uint64_t> load((size_t)n_parts, 0);
    for (uint64_t g : idx) {
        int best = 0;
        for (int p = 1; p < n_parts; ++p) if (load[p] < load[best]) best = p;
        part[g] = best;
        load[best] += cost[g];
    }
    return GP_OK;
}

} // extern "C"
