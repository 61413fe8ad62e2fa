"""
SequenceCollection: host-side mirror of the reference container
(/root/reference/src/genome_kmers/sequence_collection.py), i.e. SURVEY.md 8a rows A1/A2 -- the input
contract of the GPU hot path.  Same constructor, members, layout and error behaviour:

    forward_sba              uint8  records joined by '$'                 (ref :212-230, :663-699)
    _forward_sba_seg_starts  uint32 start of every segment                (ref :702-726)
    revcomp_sba / _revcomp_sba_seg_starts / revcomp_record_names          (ref :42-73, :905-928)

The arrays are built with vectorised NumPy instead of per-record Python loops, and two extra
constructors (`from_arrays`, `from_sba`) accept uint8 data directly, because materialising a
multi-gigabase Python `str` costs more than sorting its k-mers on a B200.
"""
import re
import shelve
from collections import Counter
from pathlib import Path
from typing import Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

_SEP = 36  # ord("$")
_STRANDS = ("forward", "reverse_complement", "both")

# IUPAC complement pairs (ref :410-427); '$' separates records and maps to itself
_COMPLEMENT_PAIRS = "AT CG GC TA RY YR SS WW KM MK BV DH HD VB NN $$".split()


def _complement_table() -> np.ndarray:
    table = np.zeros(256, dtype=np.uint8)
    for src, dst in _COMPLEMENT_PAIRS:
        table[ord(src)] = ord(dst)
    return table


_COMPLEMENT = _complement_table()
_ALLOWED = np.zeros(256, dtype=bool)
_ALLOWED[[ord(pair[0]) for pair in _COMPLEMENT_PAIRS]] = True


def _segment_ends(seg_starts: np.ndarray, sba_len: int) -> np.ndarray:
    """Inclusive end index of every segment (ref :181-185)."""
    ends = np.empty(len(seg_starts), dtype=np.int64)
    ends[:-1] = seg_starts[1:].astype(np.int64) - 2
    ends[-1] = sba_len - 1
    return ends


def reverse_complement_sba(sba: np.ndarray, complement_mapping_arr: np.ndarray = _COMPLEMENT,
                           inplace: bool = False) -> np.ndarray:
    """rc[len-1-i] = complement[sba[i]] (ref :42-73)."""
    rc = complement_mapping_arr[sba[::-1]]
    if inplace:
        sba[:] = rc
        return sba
    return rc


_UINT8_TO_U1 = np.array([chr(i) for i in range(256)], dtype="U1")
_U1_TO_UINT8 = {chr(i): i for i in range(256)}


class SequenceCollection:
    """Holds the records of a FASTA file as one '$'-joined byte array per strand."""

    def __init__(self, fasta_file_path: Union[Path, None] = None,
                 sequence_list: Union[List[Tuple[str, str]], None] = None,
                 strands_to_load: str = "forward") -> None:
        self.forward_sba = None
        self._forward_sba_seg_starts = None
        self.forward_record_names = None
        self.revcomp_sba = None
        self._revcomp_sba_seg_starts = None
        self.revcomp_record_names = None
        self._strands_loaded = None
        self._fasta_file_path = None
        self._complement_mapping_arr = _COMPLEMENT
        self._allowed_uint8 = {int(v) for v in np.flatnonzero(_ALLOWED)}
        self._allowed_bases = {chr(v) for v in self._allowed_uint8}
        # byte <-> one-character string tables the reference keeps on the object (ref :463-474)
        self._uint8_to_u1_mapping = _UINT8_TO_U1
        self._u1_to_uint8_mapping = _U1_TO_UINT8

        if fasta_file_path is None and sequence_list is None:
            return
        if fasta_file_path is not None and sequence_list is not None:
            raise ValueError("Only one of fasta_file_path and sequence_list can be specified")
        if strands_to_load not in _STRANDS:
            raise ValueError(f"strands_to_load unrecognized ({strands_to_load})")

        if fasta_file_path is not None:
            self._fasta_file_path = fasta_file_path
            names, sba, starts = _read_fasta(fasta_file_path)
            _check_alphabet(sba)
            self._verify_record_names_are_unique(names)
            self._finish(sba, starts, names, strands_to_load)
            return
        names = [name for name, _ in sequence_list]
        chunks = [np.frombuffer(seq.encode("utf-8"), dtype=np.uint8) for _, seq in sequence_list]
        self._set_from_records(names, chunks, strands_to_load, None)

    # ------------------------------------------------------------------ extra constructors
    @classmethod
    def from_arrays(cls, records: Sequence[Tuple[str, Union[np.ndarray, bytes]]],
                    strands_to_load: str = "forward") -> "SequenceCollection":
        """Like sequence_list=, but each sequence is a uint8 array / bytes of ASCII bases."""
        if strands_to_load not in _STRANDS:
            raise ValueError(f"strands_to_load unrecognized ({strands_to_load})")
        self = cls()
        names = [name for name, _ in records]
        chunks = [np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray))
                  else np.asarray(seq, dtype=np.uint8) for _, seq in records]
        self._set_from_records(names, chunks, strands_to_load, None)
        return self

    @classmethod
    def from_sba(cls, forward_sba: np.ndarray, seg_starts: np.ndarray, record_names: List[str],
                 strands_to_load: str = "forward", validate: bool = True) -> "SequenceCollection":
        """Adopt a ready forward byte array (records already joined by '$') without copying."""
        if strands_to_load not in _STRANDS:
            raise ValueError(f"strands_to_load unrecognized ({strands_to_load})")
        self = cls()
        sba = np.ascontiguousarray(forward_sba, dtype=np.uint8)
        starts = np.ascontiguousarray(seg_starts, dtype=np.uint32)
        if len(starts) != len(record_names) or len(starts) == 0:
            raise ValueError("seg_starts and record_names must have one entry per record")
        if validate:
            _check_alphabet(sba)
            if (np.diff(starts.astype(np.int64)) < 2).any() or len(sba) <= int(starts[-1]):
                raise ValueError("Each sequence in the collection must have length > 0.")
        cls._verify_record_names_are_unique(record_names)
        self._finish(sba, starts, list(record_names), strands_to_load)
        return self

    # ------------------------------------------------------------------ construction helpers
    def _set_from_records(self, names, chunks, strands_to_load, source):
        for name, chunk in zip(names, chunks):
            if len(chunk) == 0:
                if source is not None:
                    raise ValueError(
                        f"At least one empty sequence was found in the input file ({source})")
                raise ValueError(
                    "Each sequence in the collection must have length > 0.  "
                    f"Record '{name}' has a sequence lengt of 0")
        if not chunks:
            raise ValueError("the collection must contain at least one sequence")
        lengths = np.fromiter((len(c) for c in chunks), dtype=np.int64, count=len(chunks))
        starts64 = np.concatenate([[0], np.cumsum(lengths[:-1] + 1)])
        sba = np.full(int(lengths.sum()) + len(chunks) - 1, _SEP, dtype=np.uint8)
        for start, chunk in zip(starts64, chunks):
            sba[start:start + len(chunk)] = chunk
        _check_alphabet(sba)
        self._verify_record_names_are_unique(names)
        self._finish(sba, starts64.astype(np.uint32), list(names), strands_to_load)

    def _finish(self, sba, starts, names, strands_to_load):
        if strands_to_load in ("forward", "both"):
            self.forward_sba = sba
            self._forward_sba_seg_starts = starts
            self.forward_record_names = names
        if strands_to_load in ("reverse_complement", "both"):
            self.revcomp_sba = reverse_complement_sba(sba)
            self._revcomp_sba_seg_starts = self._get_opposite_strand_sba_start_indices(starts, len(sba))
            self.revcomp_record_names = names[::-1]
        self._strands_loaded = strands_to_load

    @staticmethod
    def _verify_record_names_are_unique(record_names):
        counts = Counter(record_names)
        if len(counts) != len(record_names):
            repeated = sum(1 for c in counts.values() if c > 1)
            raise ValueError(f"sequence_list contains {repeated} repeated record_names")

    @staticmethod
    def _get_fasta_record_name(line: str) -> str:
        if not line.startswith(">"):
            raise ValueError("line does not start with '>'")
        return line[1:].strip().split()[0]

    # ------------------------------------------------------------------ basic protocol
    def __len__(self) -> int:
        if self._strands_loaded in ("forward", "both"):
            return len(self._forward_sba_seg_starts)
        if self._strands_loaded == "reverse_complement":
            return len(self._revcomp_sba_seg_starts)
        raise AssertionError(f"strands_loaded ({self._strands_loaded}) not recognized")

    def __str__(self) -> str:
        strand = "reverse_complement" if self._strands_loaded == "reverse_complement" else "forward"
        sba = self.revcomp_sba if strand == "reverse_complement" else self.forward_sba
        lines = []
        for name, start, end in self.iter_records(strand):
            lines.append(f">{name}")
            lines.append(sba[start:end + 1].tobytes().decode())
        return "\n".join(lines)

    def sequence_length(self, record_num=None, record_name=None):
        if record_name is not None and record_num is not None:
            raise ValueError(
                f"record_num ({record_num}) and record_name ({record_name}) cannot both be specified")
        raise NotImplementedError()

    def strands_loaded(self) -> str:
        return self._strands_loaded

    def _strand_arrays(self, sba_strand: str):
        if sba_strand == "forward":
            return self.forward_sba, self._forward_sba_seg_starts, self.forward_record_names
        if sba_strand == "reverse_complement":
            return self.revcomp_sba, self._revcomp_sba_seg_starts, self.revcomp_record_names
        raise ValueError(f"sba_strand ({sba_strand}) not recognized")

    def _get_sba_strand_to_use(self, sba_strand: Optional[str]) -> str:
        """Strand selection rules of the reference (ref :1013-1033)."""
        if sba_strand is not None:
            if sba_strand not in ("forward", "reverse_complement"):
                raise ValueError(f"sba_strand ({sba_strand}) not recognized")
            other = "reverse_complement" if sba_strand == "forward" else "forward"
            if self._strands_loaded == other:
                raise ValueError(
                    f"sba_strand ({sba_strand}) does not match _strands_loaded ({self._strands_loaded})")
        if self._strands_loaded == "both" and sba_strand is None:
            raise ValueError("sba_strand must be specified when both strands are loaded")
        return self._strands_loaded if self._strands_loaded != "both" else sba_strand

    def iter_records(self, sba_strand: str = None) -> Iterator[Tuple[str, int, int]]:
        """(record_name, sba_start, sba_end) in record order on either strand (ref :356-391)."""
        strand = self._get_sba_strand_to_use(sba_strand)
        sba, starts, names = self._strand_arrays(strand)
        ends = _segment_ends(starts, len(sba))
        order = range(len(starts)) if strand == "forward" else range(len(starts) - 1, -1, -1)
        for seg in order:
            yield names[seg], int(starts[seg]), int(ends[seg])

    # ------------------------------------------------------------------ strand handling
    def reverse_complement(self) -> None:
        """Swap the loaded strand in place (ref :821-870)."""
        if self._strands_loaded == "both":
            raise ValueError(f"self._strands_loaded ({self._strands_loaded}) cannot be 'both'")
        if self._strands_loaded == "forward":
            sba, starts, names = self.forward_sba, self._forward_sba_seg_starts, self.forward_record_names
        else:
            sba, starts, names = self.revcomp_sba, self._revcomp_sba_seg_starts, self.revcomp_record_names
        reverse_complement_sba(sba, inplace=True)
        starts = self._get_opposite_strand_sba_start_indices(starts, len(sba))
        names.reverse()
        if self._strands_loaded == "forward":
            self.revcomp_sba, self._revcomp_sba_seg_starts, self.revcomp_record_names = sba, starts, names
            self.forward_sba = self._forward_sba_seg_starts = self.forward_record_names = None
            self._strands_loaded = "reverse_complement"
        else:
            self.forward_sba, self._forward_sba_seg_starts, self.forward_record_names = sba, starts, names
            self.revcomp_sba = self._revcomp_sba_seg_starts = self.revcomp_record_names = None
            self._strands_loaded = "forward"

    @staticmethod
    def _get_complement_mapping_array() -> np.ndarray:
        return _COMPLEMENT.copy()

    @staticmethod
    def _get_opposite_strand_sba_index(sba_idx: int, sba_len: int) -> int:
        if sba_idx < 0 or sba_idx >= sba_len:
            raise ValueError(f"sba_idx ({sba_idx}) is out of bounds")
        return sba_len - 1 - sba_idx

    @staticmethod
    def _get_opposite_strand_sba_indices(sba_indices: np.ndarray, sba_len: int) -> np.ndarray:
        if (sba_indices < 0).any() or (sba_indices >= sba_len).any():
            raise ValueError("There is at least one sba index that is out of bounds")
        return sba_len - 1 - sba_indices

    @staticmethod
    def _get_opposite_strand_sba_start_indices(sba_starts: np.ndarray, sba_len: int) -> np.ndarray:
        """New segment starts are the mirrored old segment ends, reversed (ref :905-928)."""
        ends = _segment_ends(sba_starts, sba_len)
        return (sba_len - 1 - ends[::-1]).astype(sba_starts.dtype)

    # ------------------------------------------------------------------ index -> record lookups
    def get_segment_num_from_sba_index(self, sba_idx: int, sba_strand: str = None) -> int:
        strand = self._get_sba_strand_to_use(sba_strand)
        sba, starts, _ = self._strand_arrays(strand)
        if sba_idx < 0 or sba_idx >= len(sba):
            raise IndexError(f"sba_idx ({sba_idx}) is out of bounds")
        return int(np.searchsorted(starts, sba_idx, side="right")) - 1

    def get_sba_start_end_indices_for_segment(self, segment_num: int,
                                              sba_strand: str = None) -> Tuple[int, int]:
        strand = self._get_sba_strand_to_use(sba_strand)
        sba, starts, _ = self._strand_arrays(strand)
        if segment_num < 0 or segment_num >= len(starts):
            raise ValueError(f"segment_num ({segment_num}) is out of bounds")
        end = len(sba) - 1 if segment_num == len(starts) - 1 else int(starts[segment_num + 1]) - 2
        return int(starts[segment_num]), end

    def get_record_name_from_sba_index(self, sba_idx: int, sba_strand: str = None) -> str:
        strand = self._get_sba_strand_to_use(sba_strand)
        _, starts, names = self._strand_arrays(strand)
        return names[int(np.searchsorted(starts, sba_idx, side="right")) - 1]

    def get_record_loc_from_sba_index(self, sba_idx: int, sba_strand: str = None,
                                      one_based: bool = False) -> Tuple[str, str, int]:
        """(strand symbol, record name, forward sequence index) of a byte array index (ref :930-978)."""
        strand = self._get_sba_strand_to_use(sba_strand)
        _, starts, names = self._strand_arrays(strand)
        seg = int(np.searchsorted(starts, sba_idx, side="right")) - 1
        seg_start, seg_end = self.get_sba_start_end_indices_for_segment(seg, strand)
        seq_idx = get_forward_seq_idx(sba_idx, strand, seg_start, seg_end, one_based=one_based)
        return ("+" if strand == "forward" else "-", names[seg], seq_idx)

    def generate_get_record_info_from_sba_index_func(self, one_based: bool = False):
        """Closure sba_idx -> (segment number, segment start, segment end, strand symbol, record name,
        forward sequence index) for the loaded strand (ref :1113-1187).  Errors as in the reference:
        ValueError when the index lies outside every segment (e.g. on a '$')."""
        strand = self._get_sba_strand_to_use(self.strands_loaded())
        sba, starts, names = self._strand_arrays(strand)
        names, len_sba = tuple(names), len(sba)
        symbol = "+" if strand == "forward" else "-"

        def get_record_info_from_sba_index(sba_idx: int):
            seg = get_segment_num_from_sba_index(sba_idx, strand, starts)
            seg_start, seg_end = get_sba_start_end_indices_for_segment(seg, strand, starts, len_sba)
            seq_idx = get_forward_seq_idx(sba_idx, strand, seg_start, seg_end, one_based=one_based)
            return seg, seg_start, seg_end, symbol, names[seg], seq_idx

        return get_record_info_from_sba_index

    def locate_sba_indices(self, sba_indices: np.ndarray, sba_strand: str = None,
                           one_based: bool = False):
        """Vectorised get_record_loc_from_sba_index: (segment numbers, forward sequence indices)."""
        strand = self._get_sba_strand_to_use(sba_strand)
        sba, starts, _ = self._strand_arrays(strand)
        idx = np.asarray(sba_indices, dtype=np.int64)
        seg = np.searchsorted(starts, idx, side="right") - 1
        if strand == "forward":
            seq_idx = idx - starts.astype(np.int64)[seg]
        else:
            seq_idx = _segment_ends(starts, len(sba))[seg] - idx
        return seg, seq_idx + (1 if one_based else 0)

    # ------------------------------------------------------------------ equality / persistence
    def __ne__(self, other):
        return not self.__eq__(other)

    def __eq__(self, other):
        for attr in ("forward_sba", "_forward_sba_seg_starts", "revcomp_sba", "_revcomp_sba_seg_starts"):
            a, b = getattr(self, attr), getattr(other, attr)
            if (a is None) != (b is None) or (a is not None and not np.array_equal(a, b)):
                return False
        for attr in ("forward_record_names", "revcomp_record_names", "_strands_loaded"):
            if getattr(self, attr) != getattr(other, attr):
                return False
        return True

    _PERSISTED = ("forward_sba", "_forward_sba_seg_starts", "forward_record_names", "revcomp_sba",
                  "_revcomp_sba_seg_starts", "revcomp_record_names", "_strands_loaded")

    def save(self, save_file_path: Path, mode: str = "w", format: str = "hdf5") -> None:
        """Same on-disk layout as the reference (sequence_collection.py:1331-1365 HDF5, :1407-1426 shelve),
        so that either implementation loads what the other one saved."""
        if format == "shelve":
            with shelve.open(str(save_file_path)) as db:
                for attr in self._PERSISTED:
                    db["seq_coll." + attr] = getattr(self, attr)
                db["seq_coll._fasta_file_path"] = self._fasta_file_path
        elif format == "hdf5":
            h5py = _require_h5py()
            with h5py.File(save_file_path, mode) as file:
                grp = file.create_group("seq_coll")
                empty_u8, empty_u32 = np.array([], dtype=np.uint8), np.array([], dtype=np.uint32)
                grp["forward_sba"] = empty_u8 if self.forward_sba is None else self.forward_sba
                grp["_forward_sba_seg_starts"] = (
                    empty_u32 if self._forward_sba_seg_starts is None else self._forward_sba_seg_starts)
                grp["forward_record_names"] = self.forward_record_names or []
                grp["revcomp_sba"] = empty_u8 if self.revcomp_sba is None else self.revcomp_sba
                grp["_revcomp_sba_seg_starts"] = (
                    empty_u32 if self._revcomp_sba_seg_starts is None else self._revcomp_sba_seg_starts)
                grp["revcomp_record_names"] = self.revcomp_record_names or []
                grp["_strands_loaded"] = self._strands_loaded or ""
                grp["_fasta_file_path"] = str(self._fasta_file_path or "")
        else:
            raise ValueError(f"format ({format}) not recognized")

    def load(self, load_file_path: Path, format: str = "hdf5") -> None:
        """Reads the reference's layout (sequence_collection.py:1367-1405 HDF5, :1428-1446 shelve)."""
        if format == "shelve":
            with shelve.open(str(load_file_path)) as db:
                for attr in self._PERSISTED:
                    setattr(self, attr, db["seq_coll." + attr])
                self._fasta_file_path = db["seq_coll._fasta_file_path"]
        elif format == "hdf5":
            h5py = _require_h5py()
            with h5py.File(load_file_path, "r") as file:
                grp = file["seq_coll"]

                def arr(name):
                    value = grp[name][:]
                    return None if value.shape == (0,) else value

                def names(name):
                    value = [v.decode("utf-8") for v in grp[name][:]]
                    return value or None

                self.forward_sba = arr("forward_sba")
                self._forward_sba_seg_starts = arr("_forward_sba_seg_starts")
                self.forward_record_names = names("forward_record_names")
                self.revcomp_sba = arr("revcomp_sba")
                self._revcomp_sba_seg_starts = arr("_revcomp_sba_seg_starts")
                self.revcomp_record_names = names("revcomp_record_names")
                self._strands_loaded = grp["_strands_loaded"][()].decode("utf-8") or None
                fasta = grp["_fasta_file_path"][()].decode("utf-8")
                self._fasta_file_path = Path(fasta) if fasta else None
        else:
            raise ValueError(f"format ({format}) not recognized")


def bisect_right(a, x) -> int:
    """Number of entries of the sorted sequence `a` that are <= x (module-level helper of the reference,
    ref :16-40; NumPy's searchsorted does the work here)."""
    return int(np.searchsorted(np.asarray(a), x, side="right"))


def get_segment_num_from_sba_index(sba_idx: int, sba_strand: str, sba_seg_starts: np.ndarray) -> int:
    """Segment (record) that holds a byte array index (ref :77-98); `sba_strand` is unused there too."""
    return bisect_right(sba_seg_starts, sba_idx) - 1


def get_sba_start_end_indices_for_segment(segment_num: int, sba_strand: str, sba_seg_starts: np.ndarray,
                                          len_sba: int) -> Tuple[int, int]:
    """First and last byte array index of a segment (ref :156-187)."""
    if segment_num < 0 or segment_num >= len(sba_seg_starts):
        raise ValueError(f"segment_num ({segment_num}) is out of bounds")
    last = segment_num == len(sba_seg_starts) - 1
    end = len_sba - 1 if last else int(sba_seg_starts[segment_num + 1]) - 2
    return int(sba_seg_starts[segment_num]), end


def get_forward_seq_idx(sba_idx: int, sba_strand: str, seg_sba_start_idx: int, seg_sba_end_idx: int,
                        one_based: bool = False) -> int:
    """Forward-strand sequence index of a byte array index (ref :100-152)."""
    if sba_idx < seg_sba_start_idx:
        raise ValueError(f"sba_idx ({sba_idx}) must be >= seg_sba_start_idx ({seg_sba_start_idx})")
    if sba_idx > seg_sba_end_idx:
        raise ValueError(f"sba_idx ({sba_idx}) must be <= seg_end_start_idx ({seg_sba_end_idx})")
    if seg_sba_start_idx > seg_sba_end_idx:
        raise ValueError(
            f"seg_sba_start_idx ({seg_sba_start_idx}) must be <= seg_sba_end_idx ({seg_sba_end_idx})")
    if seg_sba_start_idx < 0:
        raise ValueError(f"seg_sba_start_idx ({seg_sba_start_idx}) must be > 0")
    if sba_strand == "forward":
        seq_idx = sba_idx - seg_sba_start_idx
    elif sba_strand == "reverse_complement":
        seq_idx = seg_sba_end_idx - sba_idx
    else:
        raise ValueError(f"sba_strand ({sba_strand}) not recognized")
    return seq_idx + (1 if one_based else 0)


_ALLOWED_BYTES = bytes(int(v) for v in np.flatnonzero(_ALLOWED))


def _check_alphabet(sba: np.ndarray) -> None:
    """Every byte must be one of the reference's allowed symbols (sequence_collection.py:441-459, :693-697).
    Deleting the allowed bytes with bytes.translate leaves exactly the offenders (one C pass per chunk)."""
    bad = set()
    for lo in range(0, len(sba), 1 << 26):
        bad.update(sba[lo:lo + (1 << 26)].tobytes().translate(None, _ALLOWED_BYTES))
    if bad:
        raise ValueError(f"Sequence contains non-allowed characters! ({bad})")


_FASTA_BLOCK = 64 << 20   # bytes of file text handled per vectorised step
_STRIP_BYTES = np.zeros(256, dtype=bool)
_STRIP_BYTES[[9, 10, 11, 12, 13, 28, 29, 30, 31, 32]] = True   # what str.strip() removes (ASCII range)
_UPPER = np.arange(256, dtype=np.uint8)
_UPPER[ord("a"):ord("z") + 1] -= 32


def _fasta_block(data: np.ndarray):
    """One block of whole lines -> (header names, kept sequence bytes, offsets into them where a record starts).

    The reference walks the file line by line (sequence_collection.py:517-576): a line that starts with '>' opens
    a record, every other line contributes line.strip().upper().  Here the same is done with array operations
    over the block: line table from the newline positions, header lines masked out, leading/trailing whitespace
    of each line peeled off (whitespace inside a line stays and fails the alphabet check, as in the reference).
    """
    n = len(data)
    nl = np.flatnonzero(data == 10)
    line_start = np.concatenate([[0], nl + 1])
    line_end = np.concatenate([nl, [n]])              # exclusive
    if line_start[-1] >= n:                           # the block ends with a newline
        line_start, line_end = line_start[:-1], line_end[:-1]
    is_hdr = data[line_start] == 62                   # '>'
    lo, hi = line_start.copy(), line_end.copy()       # the stripped line is data[lo:hi]
    strip = _STRIP_BYTES
    active = np.flatnonzero(~is_hdr & (hi > lo))
    while len(active):                                # trailing whitespace ('\r' of CRLF files, blanks)
        active = active[strip[data[hi[active] - 1]]]
        hi[active] -= 1
        active = active[hi[active] > lo[active]]
    active = np.flatnonzero(~is_hdr & (hi > lo))
    while len(active):                                # leading whitespace
        active = active[strip[data[lo[active]]]]
        lo[active] += 1
        active = active[hi[active] > lo[active]]
    seq_lines = np.flatnonzero(~is_hdr & (hi > lo))
    lengths = (hi - lo)[seq_lines]
    # keep mask from a difference array: +1 at every kept line's first byte, -1 one past its last
    delta = np.zeros(n + 1, dtype=np.int8)
    delta[lo[seq_lines]] = 1
    delta[hi[seq_lines]] -= 1                         # (hi of one line is never lo of another: a '\n' lies between)
    keep = np.cumsum(delta[:-1], dtype=np.int8).view(bool)
    kept = _UPPER[data[keep]]
    # a record starts at the number of sequence bytes that precede its header line
    line_off = np.concatenate([[0], np.cumsum(lengths)])
    hdr_lines = np.flatnonzero(is_hdr)
    cuts = line_off[np.searchsorted(seq_lines, hdr_lines)]
    names = [SequenceCollection._get_fasta_record_name(
        data[line_start[i]:line_end[i]].tobytes().decode("utf-8")) for i in hdr_lines]
    return names, kept, cuts


_UPPER_TABLE = bytes(_UPPER)
_OTHER_STRIP_BYTES = [bytes([v]) for v in (9, 11, 12, 28, 29, 30, 31, 32)]   # str.strip() minus '\n', '\r'


def _upper_inplace(seq: np.ndarray) -> None:
    lower = (seq >= 97) & (seq <= 122)
    if lower.any():                                    # soft-masked genomes: a..z -> A..Z, arithmetic passes
        seq -= lower.view(np.uint8) << 5


def _fasta_sequence(segment: bytes) -> np.ndarray:
    """Sequence lines between two headers -> their stripped, upper-cased concatenation (uint8)."""
    if any(ws in segment for ws in _OTHER_STRIP_BYTES):
        # blanks or tabs somewhere: only those at the ends of a line may go (line.strip()), so take the
        # line-table path
        return _fasta_block(np.frombuffer(segment, dtype=np.uint8))[1]
    width = segment.find(b"\n")
    if width > 0 and b"\r" not in segment:
        # the usual layout -- every line `width` bases long, the last one shorter: a strided copy
        rows = len(segment) // (width + 1)
        tail = len(segment) - rows * (width + 1)
        tail_nl = 1 if (tail > 0 and segment.endswith(b"\n")) else 0
        raw = np.frombuffer(segment, dtype=np.uint8)
        if (segment.count(b"\n") == rows + tail_nl
                and (raw[width:rows * (width + 1):width + 1] == 10).all()):
            out = np.empty(rows * width + tail - tail_nl, dtype=np.uint8)
            out[:rows * width].reshape(rows, width)[...] = raw[:rows * (width + 1)].reshape(rows, width + 1)[:, :width]
            out[rows * width:] = raw[rows * (width + 1):len(raw) - tail_nl]
            _upper_inplace(out)
            return out
    # anything else (ragged lines, blank lines, CRLF): one C pass that drops the line ends and upper-cases
    return np.frombuffer(segment.translate(_UPPER_TABLE, b"\n\r"), dtype=np.uint8)


def _read_fasta(path) -> Tuple[List[str], np.ndarray, np.ndarray]:
    """FASTA -> (record names, forward byte array with '$' between records, uint32 segment starts).

    Same result as the reference's two-pass line reader (sequence_collection.py:476-576: Bowtie-style record
    names :497-515, line.strip().upper() :554), without a Python loop per line: the file is read in blocks of
    whole lines, header lines are located with bytes.find, and the text between two headers is cleaned by one
    strided copy (equal-width lines) or one bytes.translate call (NumPy line table when a line carries blanks
    that strip() would trim).
    """
    names: List[str] = []
    pieces: List[np.ndarray] = []      # sequence bytes, in file order
    starts_in_seq: List[int] = []      # for every record: sequence bytes that precede it in the whole file
    total = 0

    def take(segment: bytes):
        nonlocal total
        if segment:
            seq = _fasta_sequence(segment)
            if len(seq):
                pieces.append(seq)
                total += len(seq)

    with open(path, "rb") as handle:
        carry = b""
        while True:
            block = handle.read(_FASTA_BLOCK)
            if not block:
                text, carry = carry, b""
            else:
                cut = block.rfind(b"\n")
                if cut < 0:                            # no line end yet: keep reading
                    carry += block
                    continue
                text, carry = carry + block[:cut + 1], block[cut + 1:]
            if text:
                if b"\r" in text and text.count(b"\r") != text.count(b"\r\n"):
                    # universal newlines, like the reference's text-mode open(): a lone '\r' ends a line
                    # (a text never ends between '\r' and '\n')
                    text = re.sub(rb"\r(?!\n)", b"\n", text)
                pos = 0
                hdr = 0 if text.startswith(b">") else text.find(b"\n>") + 1   # 0 from find() = -1: no header
                if hdr == 0 and not text.startswith(b">"):
                    hdr = -1
                while hdr >= 0:
                    take(text[pos:hdr])
                    end = text.find(b"\n", hdr)
                    end = len(text) if end < 0 else end
                    names.append(SequenceCollection._get_fasta_record_name(
                        text[hdr:end].decode("utf-8")))
                    starts_in_seq.append(total)
                    pos = min(end + 1, len(text))
                    if text.startswith(b">", pos):
                        hdr = pos
                    else:
                        nxt = text.find(b"\n>", pos)
                        hdr = nxt + 1 if nxt >= 0 else -1
                take(text[pos:])
            if not block:
                break
    if not names:
        raise ValueError("the collection must contain at least one sequence")
    if starts_in_seq[0] != 0:
        # sequence text before the first header: the reference's writer runs past its array (:555-562)
        raise AssertionError("After parsing the fasta file, we expect sba to be full")
    bounds = np.asarray(starts_in_seq + [total], dtype=np.int64)
    lengths = np.diff(bounds)
    if (lengths == 0).any():
        raise ValueError(f"At least one empty sequence was found in the input file ({path})")
    n_rec = len(names)
    if total + n_rec - 1 > np.iinfo(np.uint32).max:
        raise NotImplementedError("collections of 2^32 or more positions need 64-bit segment starts")
    starts = bounds[:-1] + np.arange(n_rec)           # one '$' per earlier record
    # pieces never straddle a record (a header closes the piece before it): write them behind each other and
    # step over one '$' slot whenever a record is complete
    sba = np.full(total + n_rec - 1, _SEP, dtype=np.uint8)
    w, done, rec = 0, 0, 0                            # write position, sequence bytes written, current record
    for piece in pieces:
        while done >= bounds[rec + 1]:
            rec += 1
            w += 1
        sba[w:w + len(piece)] = piece
        w += len(piece)
        done += len(piece)
    return names, sba, starts.astype(np.uint32)


def _require_h5py():
    try:
        import h5py  # noqa: WPS433  (optional dependency, as in the reference's pyproject)
    except ImportError as exc:  # pragma: no cover - depends on the image
        raise ImportError("format='hdf5' needs h5py, which is not installed; use format='shelve'") from exc
    return h5py
