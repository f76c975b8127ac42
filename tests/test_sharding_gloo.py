"""The N > 1 path on CPU: world_size-2 `gloo` job of the CLI's host logic (shard by cost, detect,
gather on rank 0, rows in input order) must produce the same TSV as a single process."""
import os
import socket
import subprocess
import sys

from strique_b200 import sharding
from .conftest import ROOT

WORKER = os.path.join(ROOT, 'tests', '_gloo_worker.py')


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(world, out_file):
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, WORKER, out_file], env=env))
    for p in procs:
        assert p.wait(timeout=300) == 0


def test_lpt_partition_is_balanced_and_deterministic():
    costs = [(i * 7919) % 1000 + 1 for i in range(200)]
    shards = sharding.lpt_partition(costs, 8)
    assert sorted(i for s in shards for i in s) == list(range(200))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(costs)
    assert shards == sharding.lpt_partition(costs, 8)
    assert sharding.lpt_partition([], 4) == [[], [], [], []]
    assert sharding.lpt_partition([5, 1], 1) == [[0, 1]]


def test_two_rank_gloo_job_matches_single_process(tmp_path):
    one, two = str(tmp_path / 'one.tsv'), str(tmp_path / 'two.tsv')
    _run(1, one)
    _run(2, two)
    a, b = open(one).read(), open(two).read()
    assert a == b
    lines = a.strip().split('\n')
    assert lines[0].split('\t') == ['ID', 'target', 'strand', 'count', 'score_prefix', 'score_suffix', 'log_p', 'offset',
                                    'ticks', 'mod']
    ids = [l.split('\t')[0] for l in lines[1:]]
    assert ids == sorted(ids)                       # input order survives the sharded run
    assert not any(i.endswith('missing') for i in ids) and 'offtarget' not in ids
    assert len(ids) == 60 - len([i for i in range(60) if i % 17 == 5])
