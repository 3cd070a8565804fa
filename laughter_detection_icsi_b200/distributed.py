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


def pack_segments(local_results, settings):
    """Per-unit instance dicts {(thr, min_len): (n, 2) float64 array} -> (counts int64 [units, settings], data float64 [total, 2]):
    two flat buffers instead of pickled Python objects, so the final gather moves raw bytes."""
    import numpy as np
    counts = np.zeros((len(local_results), len(settings)), dtype=np.int64)
    parts = []
    for u, res in enumerate(local_results):
        for k, key in enumerate(settings):
            a = np.asarray(res[key], dtype=np.float64).reshape(-1, 2)
            counts[u, k] = len(a)
            parts.append(a)
    data = np.concatenate(parts) if parts else np.zeros((0, 2), dtype=np.float64)
    return counts, data


def unpack_segments(counts, data, settings):
    out, off = [], 0
    for u in range(counts.shape[0]):
        d = {}
        for k, key in enumerate(settings):
            n = int(counts[u, k])
            d[key] = data[off:off + n]
            off += n
        out.append(d)
    return out


def gather_segments(local_results, unit_ids, n_units, settings, dst=0):
    """The final gather of multi-GPU inference: every rank's per-unit segment lists end up on `dst`, ordered by unit index
    (None elsewhere).  Buffers travel as tensors (NCCL: staged through the GPU over NVLink; gloo: host tensors): sizes
    first (one small all_gather), then one gather of the padded flat buffers."""
    import numpy as np
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        merged = dict(zip(unit_ids, local_results))
        return [merged[i] for i in range(n_units)]
    world, rank = dist.get_world_size(), dist.get_rank()
    on_gpu = dist.get_backend() == "nccl"
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    counts, data = pack_segments(local_results, settings)
    sizes = torch.tensor([len(unit_ids), data.shape[0]], dtype=torch.int64, device=dev)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [tuple(int(x) for x in t.tolist()) for t in all_sizes]
    max_units, max_rows = max(s[0] for s in all_sizes), max(s[1] for s in all_sizes)
    n_set = len(settings)
    # one int64 buffer per rank: unit ids | counts | segment data (float64 bit patterns)
    width = max_units * (1 + n_set) + 2 * max_rows
    buf = np.zeros(width, dtype=np.int64)
    buf[:len(unit_ids)] = unit_ids
    buf[max_units:max_units + counts.size] = counts.reshape(-1)
    buf[max_units * (1 + n_set):max_units * (1 + n_set) + data.size] = data.reshape(-1).view(np.int64)
    t = torch.from_numpy(buf)
    t = t.pin_memory().to(dev, non_blocking=True) if on_gpu else t
    out = [torch.empty(width, dtype=torch.int64, device=dev) for _ in range(world)] if rank == dst else None
    dist.gather(t, out, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for r, part in enumerate(out):
        n_u, n_rows = all_sizes[r]
        a = part.cpu().numpy()
        ids = a[:n_u]
        cnt = a[max_units:max_units + n_u * n_set].reshape(n_u, n_set)
        seg = a[max_units * (1 + n_set):max_units * (1 + n_set) + 2 * n_rows].view(np.float64).reshape(n_rows, 2)
        for uid, res in zip(ids.tolist(), unpack_segments(cnt, seg, settings)):
            merged[uid] = res
    missing = [i for i in range(n_units) if i not in merged]
    if missing:
        raise RuntimeError(f"units {missing} were not processed by any rank")
    return [merged[i] for i in range(n_units)]


def pack_group(unit_ids, local_results, settings):
    """One float64 buffer for a group of units: [n_units, n_rows | unit ids | counts (units x settings) | (start, end) pairs].
    The integers are far below 2^53, so they travel exactly as doubles; the pairs are written in place (no concatenation)."""
    import numpy as np
    n_u, n_set = len(unit_ids), len(settings)
    counts = np.zeros((n_u, n_set), dtype=np.int64)
    for u, res in enumerate(local_results):
        for k, key in enumerate(settings):
            counts[u, k] = len(res[key])
    n_rows = int(counts.sum())
    head = 2 + n_u + n_u * n_set
    buf = np.empty(head + 2 * n_rows, dtype=np.float64)
    buf[0], buf[1] = n_u, n_rows
    buf[2:2 + n_u] = unit_ids
    buf[2 + n_u:head] = counts.reshape(-1)
    off = head
    for res in local_results:
        for key in settings:
            a = np.asarray(res[key], dtype=np.float64).reshape(-1)
            buf[off:off + a.size] = a
            off += a.size
    return buf


def unpack_group(buf, settings):
    """Inverse of pack_group: {unit id -> {setting -> (n, 2) float64 view into buf}}."""
    import numpy as np
    n_u, n_rows = int(buf[0]), int(buf[1])
    n_set = len(settings)
    head = 2 + n_u + n_u * n_set
    ids = buf[2:2 + n_u].astype(np.int64)
    counts = buf[2 + n_u:head].astype(np.int64).reshape(n_u, n_set)
    data = buf[head:head + 2 * n_rows].reshape(n_rows, 2)
    return dict(zip(ids.tolist(), unpack_segments(counts, data, settings)))


class StreamingGather:
    """Final gather of sharded inference that overlaps with the computation: every rank hands each finished group of units
    (e.g. the channels of one meeting) to a background thread, which packs it into one float64 buffer and sends it to `dst`
    over a gloo (host-memory) process group while the GPU works on the next group; `dst` receives on a background thread of
    its own.  `finish()` then only waits for the last group in flight.  The schedule is deterministic (every rank knows
    `groups_per_rank`), so no extra handshake is needed: dst receives group g of ranks 1, 2, ... in turn, each as a two-double
    size header followed by the payload.  No collective runs inside the loop and the NCCL communicator is not touched."""

    def __init__(self, groups_per_rank, settings, n_units, dst=0):
        import queue
        import threading
        self.settings, self.n_units, self.dst = list(settings), n_units, dst
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.rank = dist.get_rank() if self.active else 0
        self.world = dist.get_world_size() if self.active else 1
        self.groups_per_rank = [int(g) for g in groups_per_rank]
        self.merged, self.error = {}, None
        self.group = None
        if not self.active:
            return
        # host-side transport next to the (possibly NCCL) default group; created by every rank, before the clock
        self.group = dist.new_group(backend="gloo") if dist.get_backend() != "gloo" else dist.group.WORLD
        if self.rank == dst:
            self.thread = threading.Thread(target=self._receive_all, daemon=True)
        else:
            self.queue = queue.Queue()
            self.thread = threading.Thread(target=self._send_all, daemon=True)
        self.thread.start()

    def submit(self, unit_ids, local_results):
        """A finished group: `local_results[i]` = {setting -> (n, 2) float64 array} of unit `unit_ids[i]`."""
        if not self.active or self.rank == self.dst:
            self.merged.update(dict(zip([int(u) for u in unit_ids], local_results)))
        else:
            self.queue.put((list(unit_ids), local_results))

    def _send_all(self):
        try:
            for _ in range(self.groups_per_rank[self.rank]):
                unit_ids, results = self.queue.get()
                buf = torch.from_numpy(pack_group(unit_ids, results, self.settings))
                dist.send(torch.tensor([float(buf.numel())], dtype=torch.float64), dst=self.dst, group=self.group)
                dist.send(buf, dst=self.dst, group=self.group)
        except Exception as e:   # noqa: BLE001 -- surfaced by finish()
            self.error = e

    def _receive_all(self):
        try:
            for g in range(max(self.groups_per_rank)):
                for r in range(self.world):
                    if r == self.dst or g >= self.groups_per_rank[r]:
                        continue
                    size = torch.zeros(1, dtype=torch.float64)
                    dist.recv(size, src=r, group=self.group)
                    buf = torch.empty(int(size.item()), dtype=torch.float64)
                    dist.recv(buf, src=r, group=self.group)
                    self.merged.update(unpack_group(buf.numpy(), self.settings))
        except Exception as e:   # noqa: BLE001
            self.error = e

    def finish(self):
        """Waits for the transfers still in flight; returns the per-unit results ordered by unit index on dst, else None."""
        if self.active:
            self.thread.join()
            if self.error is not None:
                raise self.error
            if self.rank != self.dst:
                return None
        missing = [i for i in range(self.n_units) if i not in self.merged]
        if missing:
            raise RuntimeError(f"units {missing[:8]} were not processed by any rank")
        return [self.merged[i] for i in range(self.n_units)]


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
