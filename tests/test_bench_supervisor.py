"""bench.py's job supervisor (N > 1) on CPU: two torchrun ranks, the measurement child replaced by its
gloo test double (`ALDI_BENCH_FAKE=1`).  Held: the children rendezvous among themselves although torchrun's agent-store
flag is set in their environment; a clean run prints ONE line with attempts = 1; a rank that dies on the first attempt
makes ALL ranks restart and the line then says so (attempts = 2, the error text in `restarts`)."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _run(extra_env):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, ALDI_BENCH_FAKE="1", **extra_env)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=280, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    return r, lines


@pytest.mark.timeout(300)
def test_clean_run_prints_one_line():
    r, lines = _run({})
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] == 3.0 and d["attempts"] == 1 and d["restarts"] == []


@pytest.mark.timeout(300)
def test_dead_rank_restarts_the_whole_job_and_reports_it():
    r, lines = _run({"ALDI_BENCH_INJECT_FAIL": "1:0"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["value"] == 3.0 and d["attempts"] == 2
    assert len(d["restarts"]) == 1 and d["restarts"][0]["attempt"] == 1
    assert any("injected failure on rank 1" in w for w in d["restarts"][0]["failed"]), d["restarts"]
