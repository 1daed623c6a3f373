"""N>1 path on CPU: world_size-2 gloo run of the sharding + reduction logic bench.py uses under torchrun.
The path has no data-path collective (independent workers); the only collectives are the
timing max / token sum / TTFT gather of the benchmark."""
import json
import os
import socket
import subprocess
import sys

import pytest

import bench
from oracle.dispatch_oracle import OracleC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_user_sharding_is_the_reference_backend_pick(n):
    shards = bench.shard_users(n)
    assert sorted(u for s in shards for u in s) == list(range(bench.USERS))
    assert all(len(s) == bench.USERS // n for s in shards)
    # same assignment from the C oracle of dispatcher.rs:247-254 with capacity 64
    o = OracleC(n, capacity=bench.USERS)
    for u in range(bench.USERS):
        o.enqueue("user%02d" % u)
    ref = [[] for _ in range(n)]
    while True:
        d = o.next()
        if d is None:
            break
        ref[d[2]].append(int(d[0][4:]))
    assert shards == ref


def test_world_size_2_gloo():
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_rank_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=180) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, e[-2000:]
    res = json.loads(outs[0][0].strip().split("\n")[-1])
    assert sorted(u for s in res["shards"] for u in s) == list(range(bench.USERS))
    assert res["shards"] == bench.shard_users(2)
    assert res["max_time"] == 2.0                 # max over ranks, not the sum or rank 0's own time
    assert res["tokens"] == bench.USERS * 3       # whole-job aggregate
    assert res["log_len"] == bench.USERS // 2
