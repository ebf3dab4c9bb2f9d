"""GPU: the resident service with the real binary -- concurrent thin clients, one gap each (GAPPadder's call shape,
/root/reference/assemble_gaps.py:296-318 + MergeContigs.py:85), against the reference's golden bytes."""
import os

import pytest

from test_server_client import run_server_test

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_server_batches_concurrent_clients_on_the_gpu():
    run_server_test(os.path.join(ROOT, "build", "ContigsMerger_b200"), n_rounds=3)
