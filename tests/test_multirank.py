"""World-size-2 gloo test of the N > 1 host logic: ranks shard a batch of independent proofs with no
data-path collective and agree on the max-over-ranks time (bench.py's reduction), on CPU."""
import os
import socket
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, time
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from simple_zk_rollups_b200 import synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_proofs = 5
    mine = list(range(rank, n_proofs, world))              # round-robin, as zkr_prove_batch does
    # each rank derives its own witnesses; nothing is exchanged on the data path
    ws = [synth.generate(30, 2, seed=100 + i)[1] for i in mine]
    assert all(w[0] == 1 for w in ws)
    ms = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)              # time = max over ranks
    assert ms.item() == 10.0 + world - 1
    cnt = torch.tensor([len(mine)])
    dist.all_reduce(cnt)
    assert cnt.item() == n_proofs                            # every proof is owned by exactly one rank
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write("rank " + str(rank) + " ok\\n"); sys.stdout.flush()
""") % ROOT


def test_two_rank_gloo_sharding(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 ok" in out.stdout and "rank 1 ok" in out.stdout


SHARD_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np
    import torch, torch.distributed as dist
    from oracle import bn254 as bn
    from oracle import groth16 as g
    from simple_zk_rollups_b200 import sharding as sh
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    R = bn.R

    # 1. handle exchange protocol of Comm.connect_torch: all_gather_object keeps rank order
    mine = bytes([rank]) * sh.IPC_HANDLE_BYTES
    handles = [None] * world
    dist.all_gather_object(handles, mine)
    assert [h[0] for h in handles] == list(range(world)) and all(len(h) == 64 for h in handles)

    # 2. the sharded four-step NTT, restated with oracle transforms on the slabs this rank owns:
    #    COLS/natural -> (column NTTs, twiddle) -> all-to-all -> (row NTTs) -> ROWS/bit-reversed
    log_n = 8
    n = 1 << log_n
    k0 = sh.rows_log(log_n, world); s0 = log_n - k0
    Rr, Cc = 1 << k0, 1 << s0
    cl, rl = Cc // world, Rr // world
    import random
    rng = random.Random(5)
    x = [rng.randrange(R) for _ in range(n)]
    pack = lambda v: np.frombuffer(b"".join(int(a).to_bytes(32, "little") for a in v), dtype=np.uint8).reshape(-1, 32)
    unpack = lambda a: [int.from_bytes(a[i].tobytes(), "little") for i in range(a.shape[0])]
    slab = unpack(sh.cols_slab(pack(x), log_n, rank, world))         # [i * cl + jl]
    w = g.root_of_unity(log_n)
    out_blocks = [[None] * (rl * cl) for _ in range(world)]           # block for rank q: its rows x my columns
    for jl in range(cl):
        j = rank * cl + jl
        col = g.ntt([slab[i * cl + jl] for i in range(Rr)])           # frequency kr, natural order
        for kr in range(Rr):
            rho = g.bit_reverse(kr, k0)                               # in-place DIF leaves kr at row brev(kr)
            v = col[kr] * pow(w, j * kr, R) %% R
            out_blocks[rho // rl][(rho %% rl) * cl + jl] = v
    gathered = [None] * world
    dist.all_gather_object(gathered, out_blocks)                      # gathered[p][q] = block from p for q
    rows = [[None] * Cc for _ in range(rl)]
    for p in range(world):
        blk = gathered[p][rank]
        for il in range(rl):
            for jl in range(cl):
                rows[il][p * cl + jl] = blk[il * cl + jl]
    mine_out = []
    for il in range(rl):
        f = g.ntt(rows[il])
        mine_out += [f[g.bit_reverse(pos, s0)] for pos in range(Cc)]
    full = g.ntt(x)
    want = [full[g.bit_reverse(P, log_n)] for P in range(n)]
    assert mine_out == unpack(sh.rows_slab(pack(want), log_n, rank, world))

    # 3. MSM sharded by point range: partial sums over point_range slices add up to the full MSM
    npts = 11
    cur = bn.G1
    pts = [cur.mul(bn.G1_GEN, rng.randrange(1, 1 << 30)) for _ in range(npts)]
    ks = [rng.randrange(R) for _ in range(npts)]
    lo, hi = sh.point_range(npts, rank, world)
    part = cur.to_affine(g.msm_naive(cur, pts[lo:hi], ks[lo:hi]))
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi, part))
    assert [p[0] for p in parts][0] == 0 and parts[-1][1] == npts
    assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
    acc = None
    for _, _, pt in parts:
        acc = cur.add(acc, pt)
    assert acc == cur.to_affine(g.msm_naive(cur, pts, ks))
    dist.barrier()
    dist.destroy_process_group()
    sys.stdout.write("rank " + str(rank) + " shard ok\\n"); sys.stdout.flush()
""") % ROOT


def _torchrun(tmp_path, body, nproc):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(body)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)]
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))


def test_two_rank_gloo_sharded_ntt_and_msm(tmp_path):
    """world_size 2 on CPU: the slab layouts, the exchange pattern of the fused four-step NTT and the
    MSM point-range partition, restated with oracle arithmetic, reproduce the single-rank results."""
    out = _torchrun(tmp_path, SHARD_WORKER, 2)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank 0 shard ok" in out.stdout and "rank 1 shard ok" in out.stdout
