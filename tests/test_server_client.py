"""CPU: the resident service (gappadder_b200/host/server.cpp).  `ContigsMerger_b200 --serve SOCKET` holds the context;
the same binary with GAPPADDER_B200_SOCKET set is a thin client with the reference's command line.  Concurrent clients --
GAPPadder's own call shape, Pool(nthreads).map(run_merge) at /root/reference/assemble_gaps.py:296-318 -- are batched into one
merge_gaps call and each gets exactly the single-gap bytes.  The DP is served by the oracle shim here; tests/test_gpu_server.py
runs the real binary."""
import os
import subprocess
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import pytest

from test_contigsmerger_host import hosttest_binary, GOLD, FLAGS, CASES  # noqa: F401


def _client(binary, sock, case, td):
    d = os.path.join(td, case)
    os.makedirs(d, exist_ok=True)
    p = subprocess.run([binary] + FLAGS + ["-o", "x.info", os.path.join(GOLD, case + ".fa")], cwd=d, capture_output=True,
                       env=dict(os.environ, GAPPADDER_B200_SOCKET=sock))
    info = os.path.join(d, "x.info")
    gml = os.path.join(d, "tmp.gml")
    return (p.returncode, p.stdout, open(info, "rb").read() if os.path.exists(info) else b"", open(gml, "rb").read() if os.path.exists(gml) else b"")


def _golden(case):
    return (int(open(os.path.join(GOLD, case + ".rc")).read()), open(os.path.join(GOLD, case + ".stdout"), "rb").read(),
            open(os.path.join(GOLD, case + ".info"), "rb").read(), open(os.path.join(GOLD, case + ".gml"), "rb").read())


def run_server_test(binary, n_rounds=2):
    with tempfile.TemporaryDirectory() as td:
        sock = os.path.join(td, "gp.sock")
        srv = subprocess.Popen([binary, "--serve", sock, "--window-ms", "20"], stderr=subprocess.PIPE)
        try:
            for _ in range(200):
                if os.path.exists(sock):
                    break
                time.sleep(0.05)
            assert os.path.exists(sock), "server did not come up"
            for _ in range(n_rounds):
                with ThreadPoolExecutor(max_workers=len(CASES)) as ex:           # all cases at once: one batch on the server
                    got = list(ex.map(lambda c: _client(binary, sock, c, td), CASES))
                for c, g in zip(CASES, got):
                    assert g == _golden(c), c
            # a flag the reference rejects, through the server
            p = subprocess.run([binary, "-Q", "x"], capture_output=True, env=dict(os.environ, GAPPADDER_B200_SOCKET=sock))
            assert p.returncode == 1 and p.stdout == b"Wrong input.\n"
            subprocess.run([binary, "--shutdown"], env=dict(os.environ, GAPPADDER_B200_SOCKET=sock), timeout=30)
            srv.wait(timeout=30)
            err = srv.stderr.read().decode()
            assert "served %d gaps" % (n_rounds * len(CASES)) in err, err
            batches = int(err.split("gaps in ")[1].split()[0])
            assert batches < n_rounds * len(CASES)                                 # concurrent requests shared launches
        finally:
            if srv.poll() is None:
                srv.kill()


def test_server_batches_concurrent_clients(hosttest_binary):  # noqa: F811
    run_server_test(hosttest_binary)


def test_client_without_server_runs_in_process(hosttest_binary):  # noqa: F811
    with tempfile.TemporaryDirectory() as td:
        assert _client(hosttest_binary, os.path.join(td, "nobody.sock"), "tiny1", td) == _golden("tiny1")
