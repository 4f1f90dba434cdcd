"""N > 1 host logic on CPU (gloo, world_size 2): the rank decomposition + single all-reduce of the
accumulators, exercised with the oracle's rank partials (the reference's own round-robin
decomposition, valence.F90:1089,1162-1163, summed as xm_equalize_scalar does)."""
import os
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, path, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.oracle import Oracle
    from valence_b200 import distributed as vd
    o = Oracle(path)
    e, w, nuc = o.guess_partial(rank, world)
    o.close()
    acc = torch.tensor([e, w], dtype=torch.float64)
    vd.allreduce_sum(acc)
    energy = float(acc[0] / acc[1]) + nuc
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
        fh.write(repr(energy))
    dist.destroy_process_group()


def test_two_rank_allreduce_reproduces_serial_energy(write_input, tmp_path):
    from oracle.oracle import Oracle
    path, gold = write_input("examples__h2o")
    o = Oracle(path)
    serial = o.guess_energy()["energy"]
    o.close()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, path, port, str(tmp_path)), nprocs=2, join=True)
    vals = [float(open(tmp_path / f"rank{r}.txt").read()) for r in range(2)]
    assert vals[0] == vals[1]
    assert abs(vals[0] - serial) < 1e-11
    assert abs(vals[0] - gold["guess_energy"]) < 1e-9


@pytest.mark.parametrize("ntiles,nranks", [(0, 2), (1, 2), (7, 2), (8, 8), (1000, 3)])
def test_block_cyclic_shards_partition_the_tile_list(ntiles, nranks):
    from valence_b200 import distributed as vd
    seen = sorted(k for r in range(nranks) for k in vd.shard(ntiles, r, nranks))
    assert seen == list(range(ntiles))
    sizes = [len(vd.shard(ntiles, r, nranks)) for r in range(nranks)]
    assert max(sizes) - min(sizes) <= 1
