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
