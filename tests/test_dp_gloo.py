"""World-size-2 data-parallel logic on CPU (gloo): dialogue sharding + N_rank/N_global loss scaling +
one all-reduce(sum) of a flat gradient bucket reproduces the single-process gradient.
The model arithmetic here is the oracle (test infrastructure); the GPU path uses the same host logic
(mm-dfn_b200/dp.py) with NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, lengths, out_dir):
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import mmdfn_oracle as O
    from helpers import model_shapes
    from mmdfn_b200.dp import shard_dialogues, USED_PREFIXES
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    t, a, v, q, u, lab = O.synthetic_batch(lengths, 100, 24, 16, 2, 6, seed=1)      # global batch, padded to global T
    P = {k: w.clone().requires_grad_(True) for k, w in O.formula_weights(model_shapes(100, 24, 16, 2, 6, 1)).items()}
    lo, hi = shard_dialogues(len(lengths), rank, world)
    offs = np.cumsum([0] + list(lengths))
    sl = slice(lo, hi)
    lp = O.forward_gdf(P, t[:, sl], q[:, sl], lengths[lo:hi], a[:, sl], v[:, sl], nlayers=1, speaker_weights=(3.0, 0.0, 1.0))
    n_local, n_global = int(sum(lengths[lo:hi])), int(sum(lengths))
    loss = O.focal_loss(lp, lab[offs[lo]:offs[hi]], 1.0) * (n_local / n_global)
    loss.backward()
    names = [k for k in P if k.startswith(USED_PREFIXES)]
    flat = torch.cat([P[k].grad.reshape(-1) for k in names])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    tot = loss.detach().clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        torch.save({"flat": flat, "loss": tot, "names": names}, os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_dialogues_partition():
    from mmdfn_b200.dp import shard_dialogues
    for n in (1, 2, 5, 32, 33):
        for w in (1, 2, 4, 8):
            cuts = [shard_dialogues(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gradient_equals_single_process(tmp_path):
    import mmdfn_oracle as O
    from helpers import model_shapes
    from mmdfn_b200.dp import USED_PREFIXES
    lengths = [7, 3, 9, 4, 6]            # uneven shards: 3 + 2 dialogues, 19 vs 10 utterances
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, lengths, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "dp.pt"))
    t, a, v, q, u, lab = O.synthetic_batch(lengths, 100, 24, 16, 2, 6, seed=1)
    P = {k: w.clone().requires_grad_(True) for k, w in O.formula_weights(model_shapes(100, 24, 16, 2, 6, 1)).items()}
    lp = O.forward_gdf(P, t, q, lengths, a, v, nlayers=1, speaker_weights=(3.0, 0.0, 1.0))
    loss = O.focal_loss(lp, lab, 1.0)
    loss.backward()
    ref = torch.cat([P[k].grad.reshape(-1) for k in got["names"]])
    assert abs(float(got["loss"]) - float(loss)) < 1e-6
    assert float((got["flat"] - ref).norm() / ref.norm()) < 1e-5
    assert [k for k in P if k.startswith(USED_PREFIXES)] == got["names"]
    assert all(P[k].grad is None for k in P if not k.startswith(USED_PREFIXES))     # Adam skips exactly these
