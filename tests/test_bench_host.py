"""Host-side pieces of bench.py that need no GPU: the rank-to-core binding and the shared `config` of both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_BIND = """
import importlib.util, json, os, sys
spec = importlib.util.spec_from_file_location("bench", os.path.join(sys.argv[1], "bench.py"))
b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
before = sorted(os.sched_getaffinity(0))
info = b.bind_rank_to_cores(int(sys.argv[2]), int(sys.argv[3]))
print(json.dumps({"before": before, "after": sorted(os.sched_getaffinity(0)), "info": info}))
"""


def _bind(rank, world):
    out = subprocess.run([sys.executable, "-c", _BIND, ROOT, str(rank), str(world)], check=True, capture_output=True, text=True).stdout
    return json.loads(out.strip().splitlines()[-1])


@pytest.mark.skipif(not hasattr(os, "sched_setaffinity"), reason="no sched_setaffinity on this platform")
def test_ranks_get_disjoint_core_slices():
    n = len(os.sched_getaffinity(0))
    if n < 4:
        pytest.skip("needs at least 4 CPUs")
    r0, r1 = _bind(0, 2), _bind(1, 2)
    assert r0["info"] is not None and r1["info"] is not None
    a, b = set(r0["after"]), set(r1["after"])
    assert a and b and not (a & b) and (a | b) <= set(r0["before"])
    assert r0["info"]["threads_per_rank"] == len(a) and r1["info"]["threads_per_rank"] == len(b)
    # a single rank is left alone
    solo = _bind(0, 1)
    assert solo["info"] is None and solo["after"] == solo["before"]


def test_both_arms_describe_the_same_config():
    """The driver pairs the two arms' JSON lines by `config`; both are built by one function from the same arguments."""
    sys.path.insert(0, ROOT)
    import bench

    for cfg in ("C2", "C3", "C5"):
        args = bench.parse_args(["--config", cfg])
        ref = bench.parse_args(["--config", cfg, "--impl", "reference"])
        for world in (1, 8):
            a = bench.config_dict(args, world, 1000, 640, 480)
            b = bench.config_dict(ref, world, 1000, 640, 480)
            assert a == b and a["workload"].startswith(cfg)


def test_rank_cpus_on_a_two_socket_hyperthreaded_topology():
    """128 CPUs as Linux numbers them on a 2 x 32-core host with SMT: node 0 = 0-31,64-95, node 1 = 32-63,96-127, CPU c and
    c +/- 64 share a core.  Eight ranks: whole cores, 8 each, ranks 0-3 on node 0, 4-7 on node 1, nothing shared."""
    sys.path.insert(0, ROOT)
    import bench

    ordered = list(range(0, 32)) + list(range(64, 96)) + list(range(32, 64)) + list(range(96, 128))
    sib = lambda c: sorted({c % 64, c % 64 + 64})
    seen = set()
    for r in range(8):
        cpus, per = bench.rank_cpus(ordered, sib, r, 8)
        assert per == 8 and len(cpus) == 16 and not (seen & set(cpus))
        seen |= set(cpus)
        node = 0 if r < 4 else 1
        assert all((c % 64 < 32) == (node == 0) for c in cpus)
        assert all((c + 64) % 128 in cpus for c in cpus)  # both hardware threads of every core
    assert seen == set(range(128))
    assert bench.rank_cpus(ordered, sib, 0, 2)[0] == sorted(list(range(0, 32)) + list(range(64, 96)))
    # more ranks than cores: no binding
    assert bench.rank_cpus([0, 1], lambda c: [0, 1], 0, 2) == ([], 0)
