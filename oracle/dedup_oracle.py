"""oracle/dedup_oracle.py -- CPU restatement of the dedup stage (test infrastructure only; see oracle/README.md).

Reference: MergeContigs.py:15-70 `remove_duplicate_contained(fcontig, foutput, cutoff, rm_contained)`:
  TERefiner_1 -U   unique names                 TERefiner/refiner.cpp:1045-1140 (gnrtUniqueFa), :472-478 (cmp_vfa)
  bwa mem -a self-alignment                      NOT in /root/reference (module `bwa`, unpinned): PARITY UNPINNED
  TERefiner_1 -P -c cutoff [-g]                  refiner.cpp:660-801, Alignment.cpp:397-437, rmCotigs :401-470
The rules (unique_names, decide) follow the reference statement by statement and are pinned to the prebuilt TERefiner_1
by tests/golden/dedup/dedup_rules.json.  The alignment records are builder-defined (records_of below; the same
definition as gp_dedup_records in include/gappadder_b200.h): Evaluate -- the merger's own overlap DP, itself pinned to
the reference -- of every ordered contig pair whose query ends share a k-mer with the reference contig.
The caller passes the primitive oracles (tests/_oracle.py: gpo_evaluate, gpo_candidate_pairs, gpo_revcomp, all in
oracle/overlap_oracle.c), so this file has no dependencies of its own."""


def read_raw(text: bytes):
    """Records as rmCotigs reads them: (header line, body with '\\n' after every line, name, DP letters)."""
    recs = []
    for line in text.split(b"\n")[:-1] if text.endswith(b"\n") else text.split(b"\n"):
        if line[:1] == b">":
            name = line[1:].split(b" ")[0].split(b"\t")[0].split(b"\r")[0]
            recs.append([line, b"", name, bytearray()])
        elif recs:
            recs[-1][1] += line + b"\n"
            for ch in line.upper():
                if ch in b"\r \t":
                    continue
                recs[-1][3].append(ch if ch in b"ACGT" else ord("N"))
    return [(h, b, n, bytes(s)) for h, b, n, s in recs]


def unique_names(names):
    """gnrtUniqueFa (refiner.cpp:1045-1140): sort (name, index); a record whose name equals the previous one's goes."""
    order = sorted(range(len(names)), key=lambda i: (names[i], i))           # cmp_vfa, :472-478 (std::string order = bytes order)
    keep = [True] * len(names)
    for a, b in zip(order, order[1:]):
        if names[a] == names[b]:
            keep[b] = False                                                    # :1075-1082
    return keep


def fully_mapped(cigar, rlength, cutoff):
    """Alignment::isFullyMapped, Alignment.cpp:397-426.  cigar: [(op, len)]."""
    if len(cigar) == 1 and cigar[0][1] <= rlength and cigar[0][0] == "M":
        return True
    cnt = sum(l for o, l in cigar if o == "M")
    total = cnt + sum(l for o, l in cigar if o in "SHI")
    return total > 0 and cnt / total > cutoff


def perfect_mapped(cigar, rlength):
    """Alignment::isPerfectMapped, Alignment.cpp:428-437."""
    return len(cigar) == 1 and cigar[0][1] == rlength and cigar[0][0] == "M"


def decide(records, names, lens, cutoff, remove_contained):
    """removeDupRepeatsOfOneContigSet, refiner.cpp:660-801.  records: [(q, r, cigar)] -> removed[i]."""
    removed = [False] * len(names)
    for q, r, cigar in records:
        if not remove_contained:                                               # :717-766
            if fully_mapped(cigar, lens[q], cutoff) and names[q] > names[r]:
                iq, ir = lens[q], lens[r]
                if iq == ir:
                    removed[q] = True
                else:
                    idiff, imin = abs(iq - ir), min(iq, ir)
                    if idiff / imin <= 1.0 - cutoff:                           # :754
                        removed[q] = True
        else:                                                                  # :768-786
            if names[q] == names[r]:
                continue
            if perfect_mapped(cigar, lens[q]):
                removed[q] = True
    return removed


def overlap_size(m, n, res):
    """ContigsCompactorAction::GetOverlapSize (ContigsCompactor.h:51) over SetMergedStringConcat's cases (:108-153)."""
    if res.bcontained and res.row_end + res.nclip == m and m < n:
        merged = n
    elif res.bcontained and res.col_end + res.nclip == n and n < m:
        merged = m
    elif res.row_end + res.nclip == m:
        merged = (m - res.nclip) + (n - res.col_end)
    else:
        merged = (n - res.nclip) + (m - res.row_end)
    return m + n - res.nclip - merged


def records_of(q, r, len_q, len_r, res, frac_loss):
    """Builder-defined (gp_dedup_records): the two records one Evaluate(q, r or its reverse complement) stands for."""
    ov = overlap_size(len_q, len_r, res)
    if ov < 1 or res.score < ov * (1.0 - frac_loss):
        return []
    out = []
    for a, b, la in ((q, r, len_q), (r, q, len_r)):
        m = min(ov, la)
        out.append((a, b, [("M", m)] if m == la else [("M", m), ("S", la - m)]))
    return out


def dedup(text: bytes, cutoff, remove_contained, evaluate, candidate_pairs, revcomp, k=10, frac_loss=0.4):
    """The whole stage on one FASTA file's bytes -> (output bytes, removed names).  evaluate(s1, s2) -> result with
    score,row_end,col_end,nclip,bcontained; candidate_pairs(nodes, k) -> [(i, j)] (the merger's filter); revcomp(s)."""
    recs = read_raw(text)
    keep = unique_names([r[2] for r in recs])
    kept = [i for i, kp in enumerate(keep) if kp]
    names = [recs[i][2] for i in kept]
    seqs = [recs[i][3] for i in kept]
    lens = [len(s) for s in seqs]
    records = [(c, c, [("M", lens[c])]) for c in range(len(kept)) if lens[c]]        # bwa mem -a: every contig on itself
    for q in range(len(kept)):
        for r in range(len(kept)):
            if q == r or not lens[q] or not lens[r]:
                continue
            for ref in (seqs[r], revcomp(seqs[r])):
                if (0, 1) not in candidate_pairs([ref, seqs[q]], k):                 # the ends of q occur in ref
                    continue
                records += records_of(q, r, lens[q], lens[r], evaluate(seqs[q], ref), frac_loss)
    removed = decide(records, names, lens, cutoff, remove_contained)
    if len(kept) == len(recs) and not any(removed):
        return text, []                                                                # rmCotigs copies the file (:406-418)
    out = b"".join(recs[i][0] + b"\n" + recs[i][1] for c, i in enumerate(kept) if not removed[c])
    return out, [names[c] for c in range(len(kept)) if removed[c]]
