"""world_size-2 gloo test (CPU) of the multi-rank plumbing used by bench.py and the samplers:
group sharding without overlap, max-over-ranks timing, aggregate throughput."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moleculesde_b200.dist_util import max_over_ranks, shard_groups, sum_over_ranks
    a, b = shard_groups(1025, rank, world)
    owned = torch.zeros(1025)
    owned[a:b] = 1
    dist.all_reduce(owned)
    assert torch.all(owned == 1), "every sampling group is owned by exactly one rank"
    per_rank_ms = 100.0 + 50.0 * rank  # rank 1 is slower
    (ms,) = max_over_ranks([per_rank_ms], "cpu")
    (conf,) = sum_over_ranks([float(b - a) * 10], "cpu")
    if rank == 0:
        out.put((ms, conf))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ms, conf = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ms == 150.0       # max over ranks, not mean
    assert conf == 10250.0   # whole-job conformers = sum over ranks


def test_shard_groups_balanced():
    from moleculesde_b200.dist_util import shard_groups
    for n, w in ((1024, 8), (1025, 8), (3, 8), (0, 2)):
        ranges = [shard_groups(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(ranges[i][1] == ranges[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1


def _dp_worker(rank, world, port, out):
    """Data-parallel gradient exchange of the pretraining step: flat buffer layout, one all-reduce (sum), 1/world for Adam."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from moleculesde_b200.pretrain import ParamStore
    torch.manual_seed(0)  # identical replicas
    mods = {"a": torch.nn.Linear(5, 3), "b": torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.BatchNorm1d(2))}
    store = ParamStore(mods, torch.device("cpu"))
    # parameters are views of the flat buffer, state_dict keys unchanged
    assert mods["a"].weight.data_ptr() == store.flat.data_ptr()
    assert set(mods["b"].state_dict()) == {"0.weight", "0.bias", "1.weight", "1.bias", "1.running_mean", "1.running_var",
                                           "1.num_batches_tracked"}
    store.zero_grad()
    store.grad_view("a", "weight").fill_(float(rank + 1))      # rank-dependent gradients
    store.grad_view("b", "1.bias").fill_(10.0 * (rank + 1))
    scale = store.all_reduce()
    # bucketed exchange (PretrainStep.step: the SDE models' bucket overlaps the encoders' backward): a bucket = the contiguous slice of
    # the modules' gradients; the async form returns a work handle; together the buckets equal one whole-buffer all-reduce
    whole = store.grad.clone()
    store.grad_view("a", "weight").fill_(float(rank + 1))
    store.grad_view("a", "bias").zero_()
    for n in ("0.weight", "0.bias", "1.weight"):
        store.grad_view("b", n).zero_()
    store.grad_view("b", "1.bias").fill_(10.0 * (rank + 1))
    s2, work = store.all_reduce(("b",), async_op=True)
    assert work is not None and s2 == scale
    work.wait()
    assert torch.all(store.grad_view("a", "weight") == float(rank + 1)), "the other bucket is untouched"
    store.all_reduce(("a",))
    assert torch.equal(store.grad, whole)
    # plain Python values: a tensor sent through an mp.Queue is fetched from the PRODUCER process when the consumer unpickles it,
    # which races with this worker's exit (ConnectionRefusedError / EOFError in the parent)
    res = (scale, store.grad_view("a", "weight").flatten().tolist(), store.grad_view("b", "1.bias").flatten().tolist(), store.numel)
    if rank == 0:
        out.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_gradient_allreduce():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    scale, gw, gb, numel = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert scale == 0.5
    assert len(gw) == 15 and all(v == 3.0 for v in gw) and len(gb) == 2 and all(v == 30.0 for v in gb)   # sum over ranks; Adam multiplies by 1/world
    assert numel % 4 == 0 and numel >= 15 + 3 + 6 + 2 + 2 + 2
