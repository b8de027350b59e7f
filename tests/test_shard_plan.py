"""The split of one proof over `world` ranks (csrc/prover.cu shard_plan, reported by zkr_shard_ranges): the ranks'
ranges tile each point set exactly, the hExps slices tile the domain over the H group only, and the uniform split
(ZKR_SHARD_TASKS=0) is sharding.point_range.  No GPU: the plan is host arithmetic."""
import ctypes as C
import os
import subprocess
import sys

import pytest

from simple_zk_rollups_b200 import _lib, sharding as sh

SHAPES = [(1 << 20, 73, 1 << 20), (858_473, 73, 1 << 20), (4_000_000, 73, 1 << 22), (1000, 4, 512), (40, 3, 64), (3, 1, 2)]


def ranges(world, rank, n, l, m):
    o = (C.c_uint64 * 6)()
    assert _lib.lib().zkr_shard_ranges(world, rank, n, l, m, o) == 0
    return list(o)


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("n,l,m", SHAPES)
def test_ranges_tile_every_point_set(world, n, l, m):
    rs = [ranges(world, r, n, l, m) for r in range(world)]
    for k, total in ((0, n + 2), (2, n - l)):
        assert rs[0][k] == 0 and rs[-1][k + 1] == total
        for a, b in zip(rs, rs[1:]):
            assert a[k + 1] == b[k] and a[k] <= a[k + 1]
    g_h = sum(1 for r in rs if r[5] > r[4])
    assert g_h == (world if world <= 1 else (1 if world == 2 else world // 2)) or m < world
    h = [r[4:6] for r in rs if r[5] > r[4]]
    assert h[0][0] == 0 and h[-1][1] == m
    for a, b in zip(h, h[1:]):
        assert a[1] == b[0]
    for r in rs[g_h:]:
        assert r[4] == r[5]


def test_h_group_ranks_get_less_witness_work_at_rollup_size():
    n, l, m = 1 << 20, 73, 1 << 20
    for world in (2, 4, 8):
        rs = [ranges(world, r, n, l, m) for r in range(world)]
        first, last = rs[0][1] - rs[0][0], rs[-1][1] - rs[-1][0]
        assert first < last
        # modelled work (G1-point units, as shard_plan): within 25 % of each other across ranks
        def work(r):
            return 4.4 * (r[1] - r[0]) + (r[3] - r[2]) + ((0.85 * m + (r[5] - r[4])) if r[5] > r[4] else 0.0)
        w = [work(r) for r in rs]
        assert max(w) <= 1.25 * min(w), (world, w)


def test_uniform_split_matches_point_range():
    code = ("import ctypes as C; from simple_zk_rollups_b200 import _lib; o = (C.c_uint64 * 6)();\n"
            "import json; out = []\n"
            "for r in range(4):\n"
            "    assert _lib.lib().zkr_shard_ranges(4, r, 1001, 4, 512, o) == 0; out.append(list(o))\n"
            "print(json.dumps(out))")
    env = dict(os.environ, ZKR_SHARD_TASKS="0")
    import json
    got = json.loads(subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True,
                                    cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))).stdout)
    for r in range(4):
        assert tuple(got[r][0:2]) == sh.point_range(1003, r, 4)
        assert tuple(got[r][2:4]) == sh.point_range(997, r, 4)
        assert tuple(got[r][4:6]) == sh.point_range(512, r, 4)


def test_bad_arguments():
    o = (C.c_uint64 * 6)()
    L = _lib.lib()
    assert L.zkr_shard_ranges(0, 0, 10, 1, 8, o) != 0
    assert L.zkr_shard_ranges(2, 2, 10, 1, 8, o) != 0
    assert L.zkr_shard_ranges(2, 0, 1, 1, 8, o) != 0
