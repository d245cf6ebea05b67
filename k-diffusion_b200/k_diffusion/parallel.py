"""Batch-axis data parallelism for sampling: one process per GPU, no per-step collective.

The reference samples on several GPUs by running independent replicas on batch shards
(evaluation.py:80-90, train.py:352-364).  Here the only collective is ONE flat NCCL broadcast of
the weights at start-up; per-sample seeds make every image independent of the shard layout.
"""
import torch
import torch.distributed as dist

from . import _native

_SEED_MIX = 0x9E3779B97F4A7C15
_INIT_STREAM = 0x494E4954          # "INIT": Philox stream reserved for the initial latent


def shard_range(n, rank, world_size):
    """Contiguous shard [start, stop) of n items for `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def sample_seeds(global_seed, start, stop):
    """Per-sample 63-bit seeds that depend only on (global_seed, global sample index)."""
    out = []
    for i in range(start, stop):
        z = (int(global_seed) * 0xD1342543DE82EF95 + (i + 1) * _SEED_MIX) & (2 ** 64 - 1)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)       # splitmix64 finaliser
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        out.append((z ^ (z >> 31)) & (2 ** 63 - 1))
    return out


def init_noise(seeds, shape, sigma_max, device):
    """x = N(0, 1) * sigma_max for the given per-sample seeds ([len(seeds), *shape]); shard-invariant."""
    like = torch.empty(len(seeds), *shape, device=device, dtype=torch.float32)
    s = torch.tensor(seeds, dtype=torch.int64, device=device)
    z = _native.noise_normal(like, s, _INIT_STREAM)
    return _native.lincomb([z], [float(sigma_max)], out=z)


def broadcast_weights(module, src=0, group=None):
    """One flat broadcast of every parameter and buffer of `module` from rank `src`."""
    tensors = [t for t in list(module.parameters()) + list(module.buffers()) if t is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    total = 0
    for dtype, group_tensors in by_dtype.items():
        flat = torch.cat([t.detach().reshape(-1) for t in group_tensors])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        with torch.no_grad():
            for t in group_tensors:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n
        total += flat.numel() * flat.element_size()
    return total


def gather_samples(x, group=None):
    """all_gather of equally sized shards along the batch axis (outside any timed region)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x
    out = [torch.empty_like(x) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, x.contiguous(), group=group)
    return torch.cat(out)


class ProcessGroup:
    """The slice of `accelerate.Accelerator` the sampling driver uses (reference sample.py:37-66, evaluation.py:80-90):
    `device`, `num_processes`, `process_index`, `is_main_process`, `is_local_main_process`, `gather`, `print`,
    `wait_for_everyone`.  One process per GPU; initialises torch.distributed (NCCL) from the torchrun environment when
    WORLD_SIZE > 1, otherwise it is a single-process stand-in."""

    def __init__(self, backend=None):
        import os
        world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_process_index = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = torch.device("cuda", self.local_process_index) if torch.cuda.is_available() else torch.device("cpu")
        if self.device.type == "cuda":
            torch.cuda.set_device(self.device)
        if world > 1 and not dist.is_initialized():
            backend = backend or ("nccl" if self.device.type == "cuda" else "gloo")
            kw = dict(device_id=self.device) if backend == "nccl" else {}
            dist.init_process_group(backend, **kw)
        self.num_processes = dist.get_world_size() if dist.is_initialized() else 1
        self.process_index = dist.get_rank() if dist.is_initialized() else 0

    @property
    def is_main_process(self):
        return self.process_index == 0

    @property
    def is_local_main_process(self):
        return self.local_process_index == 0

    def gather(self, x):
        return gather_samples(x)

    def print(self, *args, **kwargs):
        if self.is_main_process:
            print(*args, **kwargs)

    def wait_for_everyone(self):
        if dist.is_initialized():
            dist.barrier()
