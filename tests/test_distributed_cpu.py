"""world_size-2 gloo test of the multi-GPU host logic: one flat weight broadcast, shard-invariant
seeds, gather.  Runs on CPU tensors (the collective path is backend-agnostic torch.distributed)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import json
    import k_diffusion as K
    torch.manual_seed(100 + rank)                                   # ranks start with DIFFERENT weights
    cfg = K.config.load_config(json.loads((ROOT / "tests/golden/cfg1_mnist_shapes.json").read_text())["config"])
    model = K.config.make_model(cfg)
    if rank == 0:
        K.synth.synth_init_(model, seed=3)
    nbytes = K.parallel.broadcast_weights(model, src=0)
    ref = K.synth.synth_state_dict({k: v.shape for k, v in model.state_dict().items()}, 3, model.state_dict())
    same = all(torch.equal(v, ref[k]) for k, v in model.state_dict().items())
    lo, hi = K.parallel.shard_range(10, rank, world)
    seeds = K.parallel.sample_seeds(5, lo, hi)
    shard = torch.tensor(seeds[:5], dtype=torch.int64)              # equal shard sizes: 5 + 5
    allv = K.parallel.gather_samples(shard)
    # the sampling driver (reference evaluation.py:80-90) over a 2-process group: every process "samples" a tensor
    # that encodes (process, call, row); the gathered result must be [p0 batch0, p1 batch0, p0 batch1, ...] cut to n
    pg = K.parallel.ProcessGroup.__new__(K.parallel.ProcessGroup)
    pg.num_processes, pg.process_index, pg.local_process_index, pg.device = world, rank, rank, torch.device("cpu")
    calls = [0]

    def sample_fn(cur):
        calls[0] += 1
        return (1000 * rank + 100 * calls[0] + torch.arange(cur + 2)).float()[:, None]       # longer than asked: the driver trims

    feats = K.evaluation.compute_features(pg, sample_fn, lambda t: t * 2, 11, 4)
    q.put((rank, same, nbytes, allv.tolist(), feats.flatten().tolist()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_broadcast_and_shards():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    sys.path.insert(0, str(ROOT / "k-diffusion_b200"))
    import k_diffusion as K
    want = []           # n = 11, P = 2 -> 6 per process in batches of 4 and 2 (cur = min(n - i, 4) as in the reference)
    for call, cur in ((1, 4), (2, 4)):
        for r in range(2):
            want += [2.0 * (1000 * r + 100 * call + k) for k in range(cur)]
    for rank, same, nbytes, allv, feats in res:
        assert feats == want[:11], (feats, want[:11])
        assert same, f"rank {rank} weights differ from rank 0 after broadcast"
        assert nbytes > 4_000_000
        assert allv == K.parallel.sample_seeds(5, 0, 10)            # gather reproduces the unsharded seed list
