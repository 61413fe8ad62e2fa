"""
CPU-only tests: the C-ABI library loads and exports every symbol include/gkb200.h declares, and the
host-side mirror of the reference API validates arguments with the reference's exact messages
(tests/test_kmers.py:262-329, :467-469 of the reference).  No compute calls (no GPU here).
"""
import os
import re

import numpy as np
import pytest

from conftest import (expected_get_kmers_tuples, filter_from_spec, get_kmers_entries, get_kmers_entry_id,
                      golden_case)
from genome_kmers import _native
from genome_kmers.kmers import (Kmers, KmerFilter, compare_sba_kmers_lexicographically,
                                crispr_ngg_pam_filter, gen_kmer_gc_content_filter_func,
                                gen_kmer_homopolymer_filter_func, gen_kmer_length_filter_func,
                                gen_no_ambiguous_bases_filter, kmer_filter_keep_all)
from genome_kmers.sequence_collection import SequenceCollection

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SL1 = [("chr1", "ATCGAATTAG")]
SL2 = [("chr1", "ATCGAATTAG"), ("chr2", "GGATCTTGCATT"), ("chr3", "GTGATTGACCCCT")]


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gkb200.h")).read()
    declared = set(re.findall(r"\b(gk_[a-z0-9_]+)\s*\(", header))
    declared -= {"gk_status", "gk_filter_id"}
    lib = _native.lib()
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, missing
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert lib.gk_version() == 100
    assert lib.gk_status_string(3) == b"unsupported on the GPU path"


def test_kmer_count_host_side():
    import ctypes

    lib = _native.lib()
    starts = np.array([0, 11, 24], dtype=np.uint64)
    n = ctypes.c_uint64(0)
    assert lib.gk_kmer_count(_native.host_ptr(starts), 3, 37, 3, ctypes.byref(n)) == 0
    assert n.value == 29
    assert lib.gk_kmer_count(_native.host_ptr(starts), 3, 37, 11, ctypes.byref(n)) == _native.GK_ERR_ARG
    assert b"exceeds the length of record" in lib.gk_last_error()


def test_sequence_collection_golden_layout():
    """tests/test_sequence_collection.py:26-50 of the reference."""
    sc = SequenceCollection(sequence_list=SL2, strands_to_load="both")
    assert sc.forward_sba.tobytes() == b"ATCGAATTAG$GGATCTTGCATT$GTGATTGACCCCT"
    assert sc._forward_sba_seg_starts.tolist() == [0, 11, 24]
    assert sc._forward_sba_seg_starts.dtype == np.uint32 and sc.forward_sba.dtype == np.uint8
    assert sc.revcomp_sba.tobytes() == b"AGGGGTCAATCAC$AATGCAAGATCC$CTAATTCGAT"
    assert sc._revcomp_sba_seg_starts.tolist() == [0, 14, 27]
    assert sc.forward_record_names == ["chr1", "chr2", "chr3"]
    assert sc.revcomp_record_names == ["chr3", "chr2", "chr1"]
    assert len(sc) == 3 and sc.strands_loaded() == "both"
    assert list(sc.iter_records("forward")) == [("chr1", 0, 9), ("chr2", 11, 22), ("chr3", 24, 36)]
    assert list(sc.iter_records("reverse_complement")) == [("chr1", 27, 36), ("chr2", 14, 25), ("chr3", 0, 12)]
    with pytest.raises(ValueError, match="sba_strand must be specified"):
        list(sc.iter_records())


def test_sequence_collection_reverse_complement_round_trip():
    sc = SequenceCollection(sequence_list=SL2)
    ref = SequenceCollection(sequence_list=SL2, strands_to_load="reverse_complement")
    sc.reverse_complement()
    assert sc == ref and sc.forward_sba is None
    sc.reverse_complement()
    assert sc == SequenceCollection(sequence_list=SL2)
    assert str(sc) == ">chr1\nATCGAATTAG\n>chr2\nGGATCTTGCATT\n>chr3\nGTGATTGACCCCT"
    assert sc.get_record_loc_from_sba_index(12) == ("+", "chr2", 1)
    assert ref.get_record_loc_from_sba_index(0) == ("-", "chr3", 12)
    assert sc.get_segment_num_from_sba_index(24) == 2
    with pytest.raises(IndexError):
        sc.get_segment_num_from_sba_index(37)


def test_sequence_collection_errors():
    with pytest.raises(ValueError, match="Only one of fasta_file_path and sequence_list"):
        SequenceCollection(fasta_file_path="x.fa", sequence_list=SL1)
    with pytest.raises(ValueError, match="strands_to_load unrecognized"):
        SequenceCollection(sequence_list=SL1, strands_to_load="nope")
    with pytest.raises(ValueError, match="non-allowed characters"):
        SequenceCollection(sequence_list=[("a", "ACGTX")])
    with pytest.raises(ValueError, match="must have length > 0"):
        SequenceCollection(sequence_list=[("a", "ACGT"), ("b", "")])
    with pytest.raises(ValueError, match="sequence_list contains 1 repeated record_names"):
        SequenceCollection(sequence_list=[("a", "ACGT"), ("a", "AC")])


def test_sequence_collection_array_constructors_and_fasta(tmp_path):
    a = SequenceCollection(sequence_list=SL2, strands_to_load="both")
    b = SequenceCollection.from_arrays([(n, s.encode()) for n, s in SL2], strands_to_load="both")
    c = SequenceCollection.from_sba(a.forward_sba, a._forward_sba_seg_starts, ["chr1", "chr2", "chr3"],
                                    strands_to_load="both")
    assert a == b == c
    fasta = tmp_path / "t.fa"
    fasta.write_text(">chr1 some description\nATCGA\nattag\n>chr2\nGGATCTTGCATT\n>chr3\tx\nGTGATTGACCCCT\n")
    assert SequenceCollection(fasta_file_path=fasta) == SequenceCollection(sequence_list=SL2)
    a.save(tmp_path / "sc.db", format="shelve")
    d = SequenceCollection()
    d.load(tmp_path / "sc.db", format="shelve")
    assert d == a


def test_kmers_argument_errors_match_reference_messages():
    sc = SequenceCollection(sequence_list=SL2)
    with pytest.raises(ValueError, match=r"min_kmer_len \(0\) must be greater than zero"):
        Kmers(sc, min_kmer_len=0)
    with pytest.raises(ValueError, match=r"max_kmer_len \(0\) must be greater than zero"):
        Kmers(sc, min_kmer_len=1, max_kmer_len=0)
    with pytest.raises(ValueError, match=r"max_kmer_len \(2\) is less than min_kmer_len \(3\)"):
        Kmers(sc, min_kmer_len=3, max_kmer_len=2)
    with pytest.raises(ValueError, match=r"min_kmer_len \(11\) must be <= the shortest sequence length \(10\)"):
        Kmers(sc, min_kmer_len=11)
    with pytest.raises(ValueError, match=r"source_strand \(reverse_complement\) does not match sequence_collection loaded strand \(forward\)"):
        Kmers(sc, source_strand="reverse_complement")
    with pytest.raises(ValueError, match=r"source_strand \(sideways\) not recognized"):
        Kmers(sc, source_strand="sideways")
    with pytest.raises(NotImplementedError, match="track_strands_separately"):
        Kmers(sc, track_strands_separately=True)
    with pytest.raises(NotImplementedError, match="double_pass"):
        Kmers(sc, method="double_pass")
    with pytest.raises(ValueError, match="method 'x' not recognized"):
        Kmers(sc, method="x")


def test_kmers_host_state_without_gpu():
    empty = Kmers()
    assert empty.kmer_sba_start_indices is None and not empty._is_initialized
    sc = SequenceCollection(sequence_list=SL2)
    km = Kmers(sc, min_kmer_len=3, max_kmer_len=3)
    assert len(km) == 29 and km._is_initialized and not km._is_sorted and not km._is_set
    assert km.min_kmer_len == 3 and km.max_kmer_len == 3 and km.kmer_source_strand == "forward"
    both = Kmers(SequenceCollection(sequence_list=SL2, strands_to_load="both"), 3, 3, source_strand="both")
    assert len(both) == 58
    with pytest.raises(ValueError, match="Did you mean to run sort"):
        km.get_kmer_count(3, min_group_size=2)
    with pytest.raises(AssertionError, match="must be sorted"):
        km.get_kmer_group_counts(3)
    with pytest.raises(ValueError, match=r"kmer_len \(0\) must be > 0"):
        km.get_kmer_count(0)


def test_filters_host_evaluation_matches_reference_examples():
    sba = np.frombuffer(b"ATCGAATTAGNNNRYACGTTGCAWSACGT$GGGGGGCCAAAATTTTACGTACGTAGG", dtype=np.uint8)
    assert kmer_filter_keep_all(sba, "forward", 0)
    no_amb = gen_no_ambiguous_bases_filter(5)
    assert no_amb(sba, "forward", 0) and not no_amb(sba, "forward", 6)
    with pytest.raises(ValueError, match="end of segment was reached"):
        no_amb(sba, "forward", 26)
    assert gen_kmer_length_filter_func(4)(sba, "forward", 25) and not gen_kmer_length_filter_func(5)(sba, "forward", 25)
    homo = gen_kmer_homopolymer_filter_func(3, 6)
    assert homo(sba, "forward", 0) and not homo(sba, "forward", 30)
    gc = gen_kmer_gc_content_filter_func(0.4, 0.6, 5)
    assert gc(sba, "forward", 0) and not gc(sba, "forward", 30)
    assert crispr_ngg_pam_filter(sba, "forward", 33) == bool(sba[54] == 71 and sba[55] == 71)
    with pytest.raises(ValueError, match="must be >= 1"):
        gen_kmer_homopolymer_filter_func(0, 5)
    with pytest.raises(ValueError, match="must be <= max_allowed_gc_frac"):
        gen_kmer_gc_content_filter_func(0.7, 0.6, 5)
    assert isinstance(no_amb, KmerFilter) and no_amb.native().p0 == 5


def test_scalar_comparator_matches_reference_docstring_example():
    """kmers.py:341-351: the comparison table in the reference's docstring."""
    sba = np.frombuffer(b"ATGGGCTGCAAGCTCGA$AATTTAGCGGCCTAGGCTTA", dtype=np.uint8)
    a, b = 7, 11
    assert [compare_sba_kmers_lexicographically(sba, sba, a, b, m)[0] for m in (1, 2)] == [0, 0]
    assert compare_sba_kmers_lexicographically(sba, sba, a, b, 3)[0] == -1
    assert compare_sba_kmers_lexicographically(sba, sba, a, b, None)[0] == -1
    assert compare_sba_kmers_lexicographically(sba, sba, 15, 36, None) == (-1, 0)
    assert compare_sba_kmers_lexicographically(sba, sba, 16, 37, None) == (0, 0)   # both terminate
    assert compare_sba_kmers_lexicographically(sba, sba, 16, 18, None) == (-1, 0)  # 'A$' < 'AAT..'


# ---------------------------------------------------------------------------------------------------
# get_kmers host logic (tuple construction, record lookup, group limits) against tuples the REAL
# reference yielded (tests/golden/golden_get_kmers.json).  The group table normally comes from the GPU
# (gk_index_groups / gk_index_groups_filtered); here a TEST DOUBLE derives it on the host with the
# scalar comparator and the host evaluation of the filters, so the rest of get_kmers runs without a GPU.
# tests/test_gpu_parity.py::test_golden_get_kmers runs the same entries through the device path.
# ---------------------------------------------------------------------------------------------------
class _HostGroupTableKmers(Kmers):
    def _host_groups(self, kmer_len, flt):
        sba = self._indexed_bytes()
        idx = self.kmer_sba_start_indices
        strand = "forward"
        kept = [p for p, s in enumerate(idx) if flt(sba, strand, int(s))]
        offsets, sizes = [], []
        for j, p in enumerate(kept):
            same = (self._is_sorted and j > 0 and compare_sba_kmers_lexicographically(
                sba, sba, int(idx[kept[j - 1]]), int(idx[p]), kmer_len)[0] == 0)
            if same:
                sizes[-1] += 1
            else:
                offsets.append(j)
                sizes.append(1)
        return (np.array(kept, dtype=np.uint64), np.array(offsets, dtype=np.uint64),
                np.array(sizes, dtype=np.uint64))

    def get_kmer_groups(self, kmer_len):
        _, offsets, sizes = self._host_groups(kmer_len, kmer_filter_keep_all)
        return offsets, sizes

    def _filtered_groups(self, kmer_len, flt):
        return self._host_groups(kmer_len, flt)


@pytest.mark.parametrize("entry", get_kmers_entries(), ids=get_kmers_entry_id)
def test_get_kmers_host_logic_against_reference_tuples(entry):
    case, qu = golden_case(entry["case"]), entry["query"]
    sc = SequenceCollection(sequence_list=[tuple(r) for r in case["seq_list"]], strands_to_load=case["strands"])
    km = _HostGroupTableKmers(sc, min_kmer_len=case["min_len"], max_kmer_len=case["max_len"],
                              source_strand=case["strands"])
    km.kmer_sba_start_indices = (case["sorted"] if qu["sorted"] else case["init"]).astype(np.uint32)
    km._is_sorted = bool(qu["sorted"])
    got = list(km.get_kmers(qu["kmer_len"], one_based_seq_index=qu["one_based"],
                            kmer_filter_func=filter_from_spec(qu["filter"]), kmer_info_to_yield=qu["info"],
                            min_group_size=qu["min_group"], max_group_size=qu["max_group"],
                            yield_first_n=qu["first_n"]))
    assert got == expected_get_kmers_tuples(entry, case)


# ---------------------------------------------------------------------------------------------------
# Row N3: persistence.  Round trip through this repo's classes, and -- in the build container, where the
# reference is mounted -- cross-loading with the REAL reference in a subprocess (both packages are named
# genome_kmers, so they cannot share a process).  HDF5 needs h5py, which the image does not have.
# ---------------------------------------------------------------------------------------------------
REFERENCE_SRC = "/root/reference/src"


def _sorted_kmers_on_host(strands="forward"):
    case = golden_case("sl2_both_k3" if strands == "both" else "sl2_k3")
    sc = SequenceCollection(sequence_list=SL2, strands_to_load=strands)
    km = Kmers(sc, min_kmer_len=3, max_kmer_len=3, source_strand=strands)
    km.kmer_sba_start_indices = case["sorted"].astype(np.uint32)
    km._is_sorted = True
    return sc, km, case


@pytest.mark.parametrize("strands", ["forward", "both"])
def test_kmers_shelve_round_trip(tmp_path, strands):
    sc, km, case = _sorted_kmers_on_host(strands)
    km.save(tmp_path / "km.db", include_sequence_collection=True, format="shelve")
    back = Kmers()
    back.load(tmp_path / "km.db", format="shelve")
    assert back == km and back._is_sorted and back.seq_coll == sc
    assert back.kmer_sba_start_indices.dtype == np.uint32
    assert np.array_equal(back.kmer_sba_start_indices, case["sorted"])
    other = Kmers()
    other.load(tmp_path / "km.db", seq_coll=sc, format="shelve")
    assert other.seq_coll is sc and other == km
    with pytest.raises(ValueError, match="format \\(nope\\) not recognized"):
        km.save(tmp_path / "x", format="nope")


_REF_PRELUDE = """
import sys, types, json
import numpy as np
sys.modules.setdefault("h5py", types.ModuleType("h5py"))
sys.path.insert(0, %r)
from genome_kmers.kmers import Kmers
from genome_kmers.sequence_collection import SequenceCollection
""" % REFERENCE_SRC


def _run_reference(script):
    import subprocess
    import sys

    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    out = subprocess.run([sys.executable, "-c", _REF_PRELUDE + script], capture_output=True, text=True, env=env,
                         cwd="/tmp", timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="the reference is only mounted in the build container")
def test_shelve_files_cross_load_with_the_reference(tmp_path):
    import json

    # ours -> reference
    sc, km, case = _sorted_kmers_on_host("forward")
    ours = str(tmp_path / "ours.db")
    km.save(ours, include_sequence_collection=True, format="shelve")
    got = json.loads(_run_reference("""
km = Kmers()
km.load(%r, format="shelve")
sc = km.seq_coll
print(json.dumps(dict(
    min=int(km.min_kmer_len), max=int(km.max_kmer_len), strand=km.kmer_source_strand,
    flags=[bool(km._is_initialized), bool(km._is_set), bool(km._is_sorted)],
    idx=km.kmer_sba_start_indices.tolist(), idx_dtype=str(km.kmer_sba_start_indices.dtype),
    sba=sc.forward_sba.tobytes().decode(), starts=sc._forward_sba_seg_starts.tolist(),
    names=sc.forward_record_names, loaded=sc.strands_loaded(),
    first=km.get_kmer_str(0, 3), text=str(sc))))
""" % ours))
    assert got["min"] == 3 and got["max"] == 3 and got["strand"] == "forward"
    assert got["flags"] == [True, False, True] and got["idx_dtype"] == "uint32"
    assert got["idx"] == case["sorted"].tolist()
    assert got["sba"] == sc.forward_sba.tobytes().decode() and got["starts"] == [0, 11, 24]
    assert got["names"] == ["chr1", "chr2", "chr3"] and got["loaded"] == "forward"
    assert got["first"] == km.get_kmer_str(0, 3) and got["text"] == str(sc)

    # reference -> ours
    theirs = str(tmp_path / "theirs.db")
    _run_reference("""
sc = SequenceCollection(sequence_list=%r, strands_to_load="forward")
km = Kmers(sc, min_kmer_len=3, max_kmer_len=3)
km.kmer_sba_start_indices = np.array(%r, dtype=np.uint32)
km._is_sorted = True
km.save(%r, include_sequence_collection=True, format="shelve")
""" % (SL2, case["sorted"].tolist(), theirs))
    back = Kmers()
    back.load(theirs, format="shelve")
    assert back == km and back.seq_coll == sc
    assert back.get_kmer_str(0, 3) == km.get_kmer_str(0, 3)


# ---------------------------------------------------------------------------------------------------
# Row N4: FASTA ingest.  The block-wise reader must give what the reference's line loop gives
# (sequence_collection.py:517-576: a line starting with '>' opens a record, any other line contributes
# line.strip().upper()), restated here line by line as the checker.
# ---------------------------------------------------------------------------------------------------
def _fasta_line_by_line(path):
    names, parts = [], []
    with open(path, "r") as handle:
        for line in handle:
            if line.startswith(">"):
                names.append(line[1:].strip().split()[0])
                parts.append([])
            else:
                parts[-1].append(line.strip().upper())
    return names, ["".join(p) for p in parts]


FASTA_TEXTS = {
    "plain": ">chr1 some description\nATCGA\nattag\n>chr2\nGGATCTTGCATT\n>chr3\tx\nGTGATTGACCCCT\n",
    "no_trailing_newline": ">a\nACGT\nAC\n>b\nGG",
    "crlf": ">a desc\r\nACGT\r\nacgt\r\n>b\r\nNNRY\r\n",
    "blank_lines_and_blanks_at_line_ends": ">a\n\nACGT  \n  \t\n\tACG\n\n>b  x y\n  TT\n\n",
    "lone_cr_line_ends": ">a\rACGT\rAC\r>b\rGG\r",
    "equal_width_lines": ">a\n" + "ACGTNRYKM\n" * 50 + "ACG\n>b\n" + "acgtacgtac\n" * 7,
    "two_short_lines_fill_one_width": ">a\nACGTACGTAC\nACGT\nACGTA\nACGTACGTAC\nAC\n",
    "many_small_records": "".join(f">r{i}\nACGTAC\nGT\n" for i in range(300)),
    "vertical_tab_and_form_feed": ">a\nACGT\x0b\n\x0cAC\n",
}


@pytest.mark.parametrize("block", [7, 64, 64 << 20])
@pytest.mark.parametrize("name", sorted(FASTA_TEXTS))
def test_fasta_reader_matches_line_by_line_semantics(tmp_path, monkeypatch, name, block):
    from genome_kmers import sequence_collection as scm

    path = tmp_path / "t.fa"
    with open(path, "w", newline="") as f:
        f.write(FASTA_TEXTS[name])
    want_names, want_seqs = _fasta_line_by_line(path)
    monkeypatch.setattr(scm, "_FASTA_BLOCK", block)
    sc = SequenceCollection(fasta_file_path=path, strands_to_load="both")
    assert sc.forward_record_names == want_names
    assert sc.forward_sba.tobytes().decode().split("$") == want_seqs
    assert sc._forward_sba_seg_starts.dtype == np.uint32
    assert sc == SequenceCollection(sequence_list=list(zip(want_names, want_seqs)), strands_to_load="both")


def test_fasta_reader_errors(tmp_path):
    def load(text):
        path = tmp_path / "e.fa"
        path.write_text(text)
        return SequenceCollection(fasta_file_path=path)

    with pytest.raises(ValueError, match="At least one empty sequence was found in the input file"):
        load(">a\n>b\nAC\n")
    with pytest.raises(ValueError, match="At least one empty sequence was found in the input file"):
        load(">a\nAC\n>b\n")
    with pytest.raises(ValueError, match="non-allowed characters"):
        load(">a\nAC GT\n")                      # strip() trims the ends of a line only
    with pytest.raises(ValueError, match="non-allowed characters"):
        load(">a\nAC>GT\n")
    with pytest.raises(ValueError, match="sequence_list contains 1 repeated record_names"):
        load(">a x\nAC\n>a y\nGT\n")
    with pytest.raises(AssertionError, match="we expect sba to be full"):
        load("ACGT\n>a\nAC\n")                   # text before the first header (ref :555-569)
    # a header that is not valid UTF-8: the reference reads the file in text mode and raises
    bad = tmp_path / "bad.fa"
    bad.write_bytes(b">chr\xff1\nACGT\n")
    with pytest.raises(UnicodeDecodeError):
        SequenceCollection(fasta_file_path=bad)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_SRC), reason="the reference is only mounted in the build container")
def test_fasta_reader_matches_the_reference_loader(tmp_path):
    import json

    paths = {}
    for name, text in FASTA_TEXTS.items():
        paths[name] = str(tmp_path / f"{name}.fa")
        with open(paths[name], "w", newline="") as f:
            f.write(text)
    got = json.loads(_run_reference("""
out = {}
for name, path in %r.items():
    sc = SequenceCollection(fasta_file_path=path, strands_to_load="both")
    out[name] = [sc.forward_sba.tobytes().decode(), sc._forward_sba_seg_starts.tolist(), sc.forward_record_names,
                 sc.revcomp_sba.tobytes().decode(), sc._revcomp_sba_seg_starts.tolist(), sc.revcomp_record_names]
print(json.dumps(out))
""" % paths))
    for name, path in paths.items():
        sc = SequenceCollection(fasta_file_path=path, strands_to_load="both")
        ours = [sc.forward_sba.tobytes().decode(), sc._forward_sba_seg_starts.tolist(), sc.forward_record_names,
                sc.revcomp_sba.tobytes().decode(), sc._revcomp_sba_seg_starts.tolist(), sc.revcomp_record_names]
        assert ours == got[name], name


def test_integration_md_stub_binds_the_library():
    """The ctypes stub INTEGRATION.md shows a maintainer is real code: it loads libgkb200.so, binds
    gk_sort_count_host with the documented argument list and defines sort_and_count (no compute call here)."""
    import ctypes

    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes, numpy as np\n.*?)```", text, flags=re.S).group(1)
    lib_path = os.path.join(ROOT, "genome-kmers_b200", "lib", "libgkb200.so")
    assert 'ctypes.CDLL("libgkb200.so")' in block
    scope = {}
    exec(block.replace('ctypes.CDLL("libgkb200.so")', f"ctypes.CDLL({lib_path!r})"), scope)
    assert callable(scope["sort_and_count"])
    bound = scope["_lib"].gk_sort_count_host
    assert len(bound.argtypes) == len(_native.SIGNATURES["gk_sort_count_host"][1]) == 13
    assert bound.restype is ctypes.c_int
