"""pytest plugin for tests/test_reference_suite.py: a minimal stand-in for pytest-mock's `mocker` fixture (the
image does not have pytest-mock), with just what the reference's SequenceCollection tests use -- mock_open
(read_data=...) and patch("builtins.open", ...).  A file "opened" in binary mode yields the same text as bytes."""
import io
from unittest import mock

import pytest


class _Mocker:
    def __init__(self):
        self._patches = []

    def mock_open(self, read_data=""):
        def opener(path, mode="r", *args, **kwargs):
            return io.BytesIO(read_data.encode()) if "b" in mode else io.StringIO(read_data)

        return opener

    def patch(self, target, new):
        patcher = mock.patch(target, new)
        patcher.start()
        self._patches.append(patcher)

    def stop(self):
        for patcher in self._patches:
            patcher.stop()


@pytest.fixture
def mocker():
    m = _Mocker()
    yield m
    m.stop()
