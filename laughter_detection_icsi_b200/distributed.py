"""Multi-GPU sharding of the inference path: one process per GPU, units = (meeting, channel) recordings -- the
reference's own unit of scale-out (one CLI process per channel: cluster_scripts/gen_eval_exp.py:99-114).
Recordings are independent, so ranks share no data in the loop; the only exchange is the final gather of the
per-channel segment lists to rank 0 (torch.distributed, NCCL on GPUs / gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_units(durations, world_size):
    """Greedy longest-first assignment of units (by duration) to ranks; returns per-rank lists of unit indices,
    each in ascending order.  Deterministic, so every rank computes the same map without communication."""
    loads = [0.0] * world_size
    shards = [[] for _ in range(world_size)]
    for i in sorted(range(len(durations)), key=lambda i: (-durations[i], i)):
        r = min(range(world_size), key=lambda r: (loads[r], r))
        shards[r].append(i)
        loads[r] += durations[i]
    return [sorted(s) for s in shards]


def gather_results(local_results, unit_ids, n_units, dst=0):
    """Gathers {unit index -> result} from all ranks on `dst`; returns the list ordered by unit index there, else None."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        merged = dict(zip(unit_ids, local_results))
        return [merged[i] for i in range(n_units)]
    payload = list(zip(unit_ids, local_results))
    out = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(payload, out, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = {}
    for part in out:
        merged.update(dict(part))
    missing = [i for i in range(n_units) if i not in merged]
    if missing:
        raise RuntimeError(f"units {missing} were not processed by any rank")
    return [merged[i] for i in range(n_units)]


def allreduce_gradients(params, world_size=None):
    """Data-parallel gradient averaging for training: the gradients of all parameters travel as ONE
    flat fp32 bucket (221 217 elements = 0.88 MB for resnet_base), summed over ranks with a single all-reduce (NCCL over
    NVLink on GPUs, gloo in the CPU tests) and divided by the world size, then written back to ``.grad``."""
    params = [p for p in params if p.grad is not None]
    if not params or not (dist.is_available() and dist.is_initialized()):
        return 0
    world_size = world_size or dist.get_world_size()
    if world_size == 1:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size)
    off = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()
