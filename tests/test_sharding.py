"""Shard invariance of the multi-GPU layer on CPU-emulated ranks: world_size 2 and 3 over gloo.
The per-shard solver in these tests is the oracle (there is no GPU here); what is under test is
optimal_control_problem_b200.sharding: partition, padding, gather order."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from optimal_control_problem_b200.sharding import gather_shards, shard_range, sharded_solve  # noqa: E402


def test_shard_range_covers_the_batch_exactly():
    for batch in (0, 1, 7, 8, 4096, 65536, 65537):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(batch, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == batch
            pos = 0
            for start, count in spans:
                assert start == pos or count == 0
                pos += count
            assert max(c for _, c in spans) == -(-batch // world) or batch == 0


def _worker(rank, world, port, batch, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import _oracle
    ora = _oracle.OracleProblem("cartpole", horizon=6, alpha=1.0, step_num=2)
    frames, refs = ora.sample_inputs(batch, 77)

    def solve_shard(f, r):
        if f.shape[0] == 0:
            return torch.zeros(0, ora.N, dtype=torch.float64), torch.zeros(0, 12, dtype=torch.float64)
        x0 = np.tile(f.numpy(), (1, ora.horizon))
        x, fo, st = ora.solve_batch(f.numpy(), r.numpy(), x0=x0)
        return torch.from_numpy(x), torch.from_numpy(st)

    x, st = sharded_solve(solve_shard, torch.from_numpy(frames), torch.from_numpy(refs))
    np.savez(Path(out_dir) / f"rank{rank}.npz", x=x.numpy(), st=st.numpy())
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,batch", [(2, 6), (3, 7), (2, 1)])
def test_sharded_solve_equals_single_rank(tmp_path, world, batch):
    import _oracle
    mp.spawn(_worker, args=(world, _free_port(), batch, str(tmp_path)), nprocs=world, join=True)
    ora = _oracle.OracleProblem("cartpole", horizon=6, alpha=1.0, step_num=2)
    frames, refs = ora.sample_inputs(batch, 77)
    x_ref, _, st_ref = ora.solve_batch(frames, refs, x0=np.tile(frames, (1, ora.horizon)))
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["x"], x_ref)          # bit-identical: same code, same inputs, any partition
        assert np.array_equal(got["st"], st_ref)


def test_gather_without_process_group_is_identity():
    t = torch.arange(12, dtype=torch.float64).reshape(4, 3)
    assert torch.equal(gather_shards(t, 4), t)
