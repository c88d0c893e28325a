"""The N>1 host logic on CPU: world_size-2 gloo ranks run the synchronous merge and the batch sharding that
bench.py uses; the 2-rank merge must reproduce the reference's two-way formula
(/root/reference/licos/federation_utils.py:47-53, oracle.federated_average)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import compressai_ref as R


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import licos_b200 as L
    from licos_b200.federated import FlatState, federated_average
    from licos_b200.sharding import shard_range

    torch.manual_seed(100 + rank)
    net = L.get_model("bmshj2018-factorized", False, 1, 1)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    state = FlatState(net)
    # views: the module still sees its own values after re-homing
    assert all(torch.equal(before[k], v) for k, v in net.state_dict().items())
    loss = [3.0, 1.0][rank]
    w = [0.25, 0.75]
    # a rank that reports a non-finite loss must not poison the others (it keeps a vanishing weight): on a copy
    torch.manual_seed(100 + rank)
    probe = L.get_model("bmshj2018-factorized", False, 1, 1)
    pstate = FlatState(probe)
    federated_average(pstate, [3.0, float("nan")][rank])
    nan_after = {k: v.clone() for k, v in probe.state_dict().items()}
    # explicit weights on one copy, the default inverse-loss rule (1/3 : 1 -> 0.25 : 0.75) on the model itself
    torch.manual_seed(100 + rank)
    explicit = L.get_model("bmshj2018-factorized", False, 1, 1)
    estate = FlatState(explicit)
    federated_average(estate, loss, weights=w)
    federated_average(state, torch.tensor(loss))
    after = {k: v.clone() for k, v in net.state_dict().items()}
    same = all(torch.allclose(v, explicit.state_dict()[k], rtol=1e-6, atol=1e-7) for k, v in after.items())
    moved = False
    try:  # the ordering guard: a module moved after FlatState was built is refused
        net.double()
        federated_average(state, loss)
    except RuntimeError:
        moved = True
    gathered = [None] * world
    dist.all_gather_object(gathered, before)
    # inverse-loss default weights: 1/3 : 1 -> 0.25 : 0.75
    inv = torch.zeros(world, dtype=torch.float64)
    inv[rank] = 1.0 / loss
    dist.all_reduce(inv)
    ok_w = abs(float(inv[rank] / inv.sum()) - w[rank]) < 1e-12
    lo, hi = shard_range(515, rank, world)
    dist.barrier()
    dist.destroy_process_group()
    npy = lambda d: {k: v.numpy() for k, v in d.items()}  # plain arrays: no fd sharing after exit
    q.put((rank, npy(after), [npy(g) for g in gathered], ok_w and same and moved, (lo, hi), npy(nan_after)))


def test_two_rank_merge_equals_reference_formula():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, after0, gathered, okw0, r0, nan0), (_, after1, _, okw1, r1, nan1) = res
    tt = lambda d: {k: torch.from_numpy(v) for k, v in d.items()}
    after0, after1, gathered, nan0, nan1 = tt(after0), tt(after1), [tt(g) for g in gathered], tt(nan0), tt(nan1)
    for k, v in gathered[0].items():  # NaN-loss rank 1 contributed (almost) nothing: everyone holds rank 0's model
        if v.dtype == torch.float32 and v.numel() > 0:
            assert torch.allclose(nan0[k], v, rtol=1e-5, atol=1e-7) and torch.equal(nan0[k], nan1[k]), k
    assert okw0 and okw1
    # reference rule with rank 0 as "local" (loss 3, best 1 -> weight 0.25) and rank 1 as "central"
    fp = {k: v for k, v in gathered[0].items() if v.dtype == torch.float32 and v.numel() > 0}
    expect = R.federated_average(fp, gathered[1], loss=3.0, best_loss=1.0)
    for k, v in expect.items():
        assert torch.allclose(after0[k], v, atol=1e-6), k
        assert torch.equal(after0[k], after1[k]), k  # both ranks adopt the same model
    assert r0 == (0, 258) and r1 == (258, 515)
