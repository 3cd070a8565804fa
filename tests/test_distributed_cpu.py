"""World-size-2 gloo test of the host-side sharding / gather logic used for multi-GPU inference."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from laughter_detection_icsi_b200 import distributed as ldd


def test_shard_units_balanced_and_complete():
    durations = [3600, 1800, 3500, 10, 1800, 3600, 900]
    shards = ldd.shard_units(durations, 3)
    assert sorted(i for s in shards for i in s) == list(range(len(durations)))
    loads = [sum(durations[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(durations)
    assert ldd.shard_units(durations, 3) == shards  # deterministic
    assert ldd.shard_units([1, 1], 4) == [[0], [1], [], []]


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    durations = [50.0, 10.0, 30.0, 20.0, 40.0]
    mine = ldd.shard_units(durations, world)[rank]
    # stand-in for the per-channel result of the GPU pipeline: {setting: [(start, end)]}
    local = [{(0.5, 0.2): [(float(u), float(u) + durations[u] / 100.0)], "rank": rank} for u in mine]
    merged = ldd.gather_results(local, mine, len(durations), dst=0)
    t = torch.tensor([float(len(mine))])
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        assert t.item() == len(durations)
        assert [list(m[(0.5, 0.2)][0])[0] for m in merged] == [0.0, 1.0, 2.0, 3.0, 4.0]
        torch.save({"ranks": [m["rank"] for m in merged]}, out_path)
    else:
        assert merged is None
    dist.destroy_process_group()


def test_two_rank_gather_over_gloo(tmp_path):
    out = str(tmp_path / "merged.pt")
    mp.spawn(_worker, args=(2, 29531 + os.getpid() % 500, out), nprocs=2, join=True)
    ranks = torch.load(out)["ranks"]
    assert sorted(set(ranks)) == [0, 1] and len(ranks) == 5


def test_single_process_gather_is_identity():
    assert ldd.gather_results(["a", "b"], [1, 0], 2) == ["b", "a"]


def _grad_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5.0) * (rank + 1)
    # params[2] has no gradient on any rank: it is left out of the bucket
    n = ldd.allreduce_gradients(params, world)
    assert n == 17
    if rank == 0:
        torch.save({"g0": params[0].grad, "g1": params[1].grad, "g2": params[2].grad}, out_path)
    dist.destroy_process_group()


def test_data_parallel_gradient_bucket_over_gloo(tmp_path):
    """The flat-bucket gradient all-reduce used for data-parallel training: mean over ranks, written back in place."""
    out = str(tmp_path / "grads.pt")
    mp.spawn(_grad_worker, args=(2, 29431 + os.getpid() % 500, out), nprocs=2, join=True)
    g = torch.load(out)
    assert torch.equal(g["g0"], torch.full((3, 4), 1.5)) and torch.equal(g["g1"], torch.arange(5.0) * 1.5) and g["g2"] is None
    # single process: nothing to do
    p = torch.nn.Parameter(torch.zeros(2)); p.grad = torch.ones(2)
    assert ldd.allreduce_gradients([p]) == 0 and torch.equal(p.grad, torch.ones(2))


def test_train_mirror_helpers():
    from laughter_detection_icsi_b200 import train as ld_train
    b = ld_train.synthetic_lad_batch(16, seed=3)
    assert b["inputs"].shape == (16, 100, 44) and b["inputs"].dtype == torch.float32 and b["is_laugh"].dtype == torch.int32
    acc, prec, rec = ld_train._calc_metrics(torch.tensor([0.9, 0.2, 0.8, 0.4]), torch.tensor([1.0, 0.0, 0.0, 1.0]))
    assert (acc, prec, rec) == (0.5, 0.5, 0.5)
    args = ld_train.build_parser().parse_args(["--config", "resnet_base", "--checkpoint_dir", "ck"])
    # string-typed numerics like the reference's argparse (train.py:68-117)
    assert args.num_epochs == 1 and args.dropout_rate == '0.5' and args.gradient_accumulation_steps == '1' and args.lhotse_dir == 'lhotse'


def _seg_worker(rank, world, port, out_path):
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    settings = [(0.5, 0.0), (0.5, 0.2), (0.9, 0.0)]
    durations = [5.0, 1.0, 3.0, 2.0, 4.0, 6.0, 0.5]
    mine = ldd.shard_units(durations, world)[rank]

    def result(u):   # unit u: u segments for the first setting, one for the second, none for the third
        return {settings[0]: np.array([[u + 0.125 * i, u + 0.125 * i + 0.1] for i in range(u)], dtype=np.float64).reshape(-1, 2),
                settings[1]: np.array([[float(u), u + 0.5]]), settings[2]: np.zeros((0, 2))}
    merged = ldd.gather_segments([result(u) for u in mine], mine, len(durations), settings, dst=0)
    if rank == 0:
        ok = all(np.array_equal(merged[u][k], result(u)[k]) for u in range(len(durations)) for k in settings)
        torch.save({"ok": ok, "n": len(merged)}, out_path)
    else:
        assert merged is None
    dist.destroy_process_group()


def test_two_rank_segment_gather_as_flat_buffers(tmp_path):
    """The final gather of corpus-shaped inference (bench.py --config corpus): packed int64/float64 buffers, not pickles."""
    out = str(tmp_path / "seg.pt")
    mp.spawn(_seg_worker, args=(2, 29331 + os.getpid() % 500, out), nprocs=2, join=True)
    r = torch.load(out)
    assert r["ok"] and r["n"] == 7
    import numpy as np
    settings = [(0.5, 0.2)]
    single = ldd.gather_segments([{settings[0]: np.array([[1.0, 2.0]])}], [0], 1, settings)
    assert np.array_equal(single[0][settings[0]], np.array([[1.0, 2.0]]))
    c, d = ldd.pack_segments([{settings[0]: np.zeros((0, 2))}, {settings[0]: np.array([[0.5, 0.75]])}], settings)
    assert c.tolist() == [[0], [1]] and d.tolist() == [[0.5, 0.75]]
    back = ldd.unpack_segments(c, d, settings)
    assert back[0][settings[0]].shape == (0, 2) and back[1][settings[0]].tolist() == [[0.5, 0.75]]


def _stream_worker(rank, world, port, out_path):
    import numpy as np
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    settings = [(0.5, 0.0), (0.5, 0.2), (0.9, 0.0)]
    durations = [3.0] * 11
    shards = ldd.shard_units(durations, world)
    group = 2
    groups_per_rank = [-(-len(s) // group) for s in shards]

    def result(u):
        return {settings[0]: np.array([[u + 0.125 * i, u + 0.125 * i + 0.1] for i in range(u + 1)], dtype=np.float64).reshape(-1, 2),
                settings[1]: np.array([[float(u), u + 0.5]]), settings[2]: np.zeros((0, 2))}
    sg = ldd.StreamingGather(groups_per_rank, settings, len(durations), dst=0)
    mine = shards[rank]
    for g0 in range(0, len(mine), group):
        ids = mine[g0:g0 + group]
        sg.submit(ids, [result(u) for u in ids])
    merged = sg.finish()
    if rank == 0:
        ok = all(np.array_equal(merged[u][k], result(u)[k]) for u in range(len(durations)) for k in settings)
        torch.save({"ok": ok, "n": len(merged)}, out_path)
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_streaming_gather_overlaps_transfers_with_work(tmp_path):
    """bench.py --config corpus: finished groups travel to rank 0 on background threads while the next group is computed."""
    out = str(tmp_path / "stream.pt")
    mp.spawn(_stream_worker, args=(3, 29131 + os.getpid() % 500, out), nprocs=3, join=True)
    r = torch.load(out)
    assert r["ok"] and r["n"] == 11
    import numpy as np
    settings = [(0.5, 0.2)]
    sg = ldd.StreamingGather([1], settings, 2)          # single process: identity
    sg.submit([1, 0], [{settings[0]: np.array([[1.0, 2.0]])}, {settings[0]: np.zeros((0, 2))}])
    got = sg.finish()
    assert got[1][settings[0]].tolist() == [[1.0, 2.0]] and got[0][settings[0]].shape == (0, 2)
    buf = ldd.pack_group([4, 9], [{settings[0]: np.array([[0.5, 0.75]])}, {settings[0]: np.zeros((0, 2))}], settings)
    back = ldd.unpack_group(buf, settings)
    assert back[4][settings[0]].tolist() == [[0.5, 0.75]] and back[9][settings[0]].shape == (0, 2)
