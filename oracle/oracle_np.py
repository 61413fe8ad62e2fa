"""
NumPy restatement of the hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Independent of gk_oracle.c: instead of running a comparator sort it rank-encodes every k-mer
window into 4-bit symbols ('$'/end-of-record = 0, then the 15 IUPAC letters in ASCII order,
which is the order kmers.py:381-388 compares in), packs 16 symbols per uint64 and uses
np.lexsort with the start index as the last key -- the reference's break_ties=True order
(kmers.py:1710-1711).  Symbols after a terminator are zero, which reproduces "the shorter k-mer
sorts first" (kmers.py:360-378).
"""
import numpy as np

_ALPHABET = b"$ABCDGHKMNRSTVWY"  # ascending ASCII == ascending rank, sequence_collection.py:441-459
_RANK = np.zeros(256, dtype=np.uint8)
for _r, _b in enumerate(_ALPHABET):
    _RANK[_b] = _r


def window_words(sba: np.ndarray, starts: np.ndarray, max_len: int) -> np.ndarray:
    """uint64[n, ceil(max_len/16)] terminator-aware 4-bit packing of each window."""
    sba = np.asarray(sba, dtype=np.uint8)
    starts = np.asarray(starts, dtype=np.int64)
    n = len(starts)
    n_words = (max_len + 15) // 16
    words = np.zeros((n, n_words), dtype=np.uint64)
    alive = np.ones(n, dtype=bool)
    padded = np.concatenate([sba, np.full(max_len + 1, 36, dtype=np.uint8)])
    for j in range(max_len):
        sym = _RANK[padded[starts + j]]
        alive &= sym != 0
        sym = np.where(alive, sym, 0).astype(np.uint64)
        words[:, j // 16] |= sym << np.uint64(4 * (15 - (j % 16)))
    return words


def effective_max_len(seg_starts, sba_len, max_len):
    if max_len is not None:
        return int(max_len)
    st = np.asarray(seg_starts, dtype=np.int64)
    ends = np.concatenate([st[1:] - 2, [sba_len - 1]])
    return int((ends - st + 1).max())


def sort_indices(sba, starts, seg_starts, max_len):
    """Canonical (k-mer, start) order of the given start indices."""
    m = effective_max_len(seg_starts, len(sba), max_len)
    starts = np.asarray(starts, dtype=np.uint64)
    words = window_words(sba, starts, m)
    keys = [starts] + [words[:, w] for w in range(words.shape[1] - 1, -1, -1)]
    order = np.lexsort(keys)
    return starts[order]


def group_sizes(sba, sorted_starts, seg_starts, kmer_len):
    """Run lengths of equal kmer_len-windows over an already sorted start array."""
    m = effective_max_len(seg_starts, len(sba), kmer_len)
    words = window_words(sba, sorted_starts, m)
    if len(words) == 0:
        return np.zeros(0, dtype=np.int64)
    head = np.ones(len(words), dtype=bool)
    head[1:] = (words[1:] != words[:-1]).any(axis=1)
    pos = np.flatnonzero(head)
    return np.diff(np.concatenate([pos, [len(words)]])).astype(np.int64)


def group_hist(sizes, min_group=1, max_group=None, max_bin=1000000):
    """kmers.py:514-518 applied to a vector of group sizes."""
    sizes = np.asarray(sizes, dtype=np.int64)
    keep = sizes >= min_group
    if max_group is not None:
        keep &= sizes <= max_group
    sizes = sizes[keep]
    hist = np.bincount(np.minimum(sizes, max_bin), minlength=max_bin + 1).astype(np.int64)
    return hist, int(sizes.sum())
