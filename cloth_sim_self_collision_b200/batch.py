"""Scene batches across GPUs (BASELINE.json configs[4]): independent cloth scenes, contiguous blocks of scenes per
rank, NO collective on the data path.  torch.distributed is used only to gather per-scene results (checksums,
statistics, optionally states) on rank 0 and to take the max of the per-rank device times.

The functions here are pure host logic over `torch.distributed` (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np


def shard_scenes(nscene_total, world, rank):
    """Contiguous block of global scene indices owned by `rank`: the first `nscene_total % world` ranks own one
    more scene.  Returns range(begin, end)."""
    if world < 1 or not (0 <= rank < world) or nscene_total < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(nscene_total, world)
    begin = rank * base + min(rank, extra)
    return range(begin, begin + base + (1 if rank < extra else 0))


def scene_gravity(base_gravity, scene):
    """Per-scene gravity of config 5: gravity-y perturbed by 0.001 per scene so that trajectories differ
    (SURVEY.md section 8d); scene 0 is config 3 itself."""
    return (float(base_gravity[0]), float(base_gravity[1]) + 0.001 * scene, float(base_gravity[2]))


def gather_scene_values(dist, local_values, nscene_total, device="cpu"):
    """All ranks contribute one float64 per owned scene (in shard order); every rank receives the values of all
    scenes in global scene order.  Shards may differ in length by one, so the exchange pads to the longest."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    mine = shard_scenes(nscene_total, world, rank)
    if len(local_values) != len(mine):
        raise ValueError("rank %d owns %d scenes but passed %d values" % (rank, len(mine), len(local_values)))
    longest = max(len(shard_scenes(nscene_total, world, r)) for r in range(world))
    buf = torch.zeros(max(longest, 1), dtype=torch.float64, device=device)
    if len(mine):
        buf[:len(mine)] = torch.as_tensor(np.asarray(local_values, dtype=np.float64), device=device)
    parts = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = np.zeros(nscene_total, dtype=np.float64)
    for r in range(world):
        sh = shard_scenes(nscene_total, world, r)
        out[sh.start:sh.stop] = parts[r][:len(sh)].cpu().numpy()
    return out


def max_over_ranks(dist, values, device="cpu"):
    """Element-wise maximum of a short list of floats over all ranks (device times of the timed region)."""
    import torch
    t = torch.as_tensor(np.asarray(values, dtype=np.float64), device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def state_checksum(x):
    """The survey's trajectory checksum: sum_i x_i * (1 + i mod 7)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    return float(np.dot(x, 1.0 + (np.arange(x.size) % 7)))
