"""
Drop-in check in the build container: the REFERENCE'S OWN test module for SequenceCollection
(/root/reference/tests/test_sequence_collection.py, 76 tests: byte arrays, segment starts, reverse complement,
record lookups, FASTA loading, error cases) runs unmodified against THIS repo's genome_kmers package.
The reference is not copied: pytest is pointed at the file where it lies, with this repo's package first on the
import path.  Skipped where /root/reference does not exist (the GPU box).  The two HDF5 save/load tests need
h5py, which the image lacks (they cannot run against the reference itself here either).
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests/test_sequence_collection.py"


@pytest.mark.skipif(not os.path.isfile(REF_TESTS), reason="the reference is only mounted in the build container")
def test_reference_sequence_collection_tests_pass_against_this_package(tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "genome-kmers_b200"), os.path.join(ROOT, "tests")])
    # an empty ini file as the configuration: the reference's pyproject would put ITS src/ on the path
    ini = tmp_path / "pytest.ini"
    ini.write_text("[pytest]\n")
    out = subprocess.run(
        [sys.executable, "-m", "pytest", REF_TESTS, "-q", "--no-header", "-p", "no:cacheprovider",
         "-p", "ref_suite_plugin", "-c", str(ini), "--rootdir", str(tmp_path), "-rf"],
        capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=900)
    text = out.stdout + out.stderr
    # the module under test must be this repo's, not the reference's
    probe = subprocess.run([sys.executable, "-c", "import genome_kmers.sequence_collection as m; print(m.__file__)"],
                           capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert os.path.join(ROOT, "genome-kmers_b200") in probe.stdout, probe.stdout + probe.stderr
    passed = int(re.search(r"(\d+) passed", text).group(1)) if re.search(r"(\d+) passed", text) else 0
    failed = re.findall(r"^FAILED \S+::(\S+)", text, flags=re.M)
    try:
        import h5py  # noqa: F401
        allowed = set()
    except ImportError:
        allowed = {"test_save_load_01", "test_save_load_02"}      # HDF5 round trips
    unexpected = [f for f in failed if f.split("::")[-1] not in allowed]
    assert not unexpected, text[-3000:]
    assert passed >= 74, text[-3000:]
