"""CPU: the builder-written semi-global definition (oracle/overlap_oracle.c gpo_semiglobal -- BWA parity UNPINNED, see
its header) against an independent brute-force statement of the same definition in pure Python: for every end column
the global alignment score of the flank against every contig[j0:j_end) by a plain Needleman-Wunsch."""
import random

from _oracle import oracle_semiglobal


def _nw(a, b, mismatch, indel):
    prev = [j * indel for j in range(len(b) + 1)]
    for i in range(1, len(a) + 1):
        cur = [i * indel] + [0] * len(b)
        for j in range(1, len(b) + 1):
            cur[j] = max(prev[j - 1] + (1 if a[i - 1] == b[j - 1] else mismatch), prev[j] + indel, cur[j - 1] + indel)
        prev = cur
    return prev[len(b)]


def _brute(flank, contig, mismatch, indel):
    best = None
    for end in range(len(contig) + 1):
        for start in range(end + 1):
            s = _nw(flank, contig[start:end], mismatch, indel)
            # maximal score; then the smallest end; then the largest start
            key = (s, -end, start)
            if best is None or key > best:
                best = key
    return best[0], best[2], -best[1]


def test_semiglobal_definition_small_cases():
    rng = random.Random(5)
    for case in range(120):
        alpha = rng.choice(["AC", "ACGT", "ACGT", "ACGTN"])
        m, n = rng.randint(0, 9), rng.randint(0, 14)
        contig = "".join(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.6 and n >= 3:
            st = rng.randrange(0, n - 1)
            flank = list(contig[st:st + max(1, min(m, n - st))])
            for p in range(len(flank)):
                if rng.random() < 0.15:
                    flank[p] = rng.choice(alpha)
            flank = "".join(flank)
        else:
            flank = "".join(rng.choice(alpha) for _ in range(m))
        for mismatch, indel in ((-2, -2), (-1, -3), (-3, -1), (0, -1)):
            r = oracle_semiglobal(flank.encode(), contig.encode(), mismatch, indel)
            assert (r.score, r.col_start, r.col_end) == _brute(flank, contig, mismatch, indel), (flank, contig, mismatch, indel)


def test_semiglobal_known_answers():
    r = oracle_semiglobal(b"ACGTACGT", b"TTTTACGTACGTTTT")
    assert (r.score, r.col_start, r.col_end) == (8, 4, 12)
    r = oracle_semiglobal(b"", b"ACGT")
    assert (r.score, r.col_start, r.col_end) == (0, 0, 0)
    r = oracle_semiglobal(b"ACG", b"")
    assert (r.score, r.col_start, r.col_end) == (-6, 0, 0)
