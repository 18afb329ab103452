"""Host-side logic of the sharded conference bus (wmix_b200/conference.py), on CPU:
placement plans, and the world_size-2 exchange wiring over gloo with a numpy stand-in for the CUDA
backend (the kernels themselves are covered by tests/test_gpu_parity.py and tests/test_multi_gpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from wmix_b200.conference import ConferencePlan, ShardedConference  # noqa: E402


@pytest.mark.parametrize("placement", ["striped", "local"])
@pytest.mark.parametrize("sizes,world", [([1024] * 64, 8), ([16] * 4096, 8), ([1, 2, 3, 58, 7, 0, 5], 2), ([5], 4), ([3, 9, 4], 1)])
def test_plan_partitions_every_participant_exactly_once(placement, sizes, world):
    plan = ConferencePlan(sizes, world, placement)
    seen = np.concatenate([plan.local_members(r) for r in range(world)])
    assert sorted(seen.tolist()) == list(range(plan.total))
    for r in range(world):
        cs = plan.local_conf_start(r)
        assert cs[0] == 0 and cs[-1] == plan.local_count(r) == len(plan.local_members(r)) and (np.diff(cs) >= 0).all()
        # local order is conference-major and members keep their conference
        m = plan.local_members(r)
        for c in range(plan.n_conf):
            seg = m[cs[c]:cs[c + 1]]
            assert ((seg >= plan.global_start[c]) & (seg < plan.global_start[c + 1])).all()
    if placement == "local":
        assert not plan.spans_ranks()
    if placement == "striped" and world > 1 and max(sizes) > 1:
        assert plan.spans_ranks()
    if placement == "striped":
        per_rank = [plan.local_count(r) for r in range(world)]
        assert max(per_rank) - min(per_rank) <= plan.n_conf      # balanced to within one member per conference


def test_plan_config5_shapes():
    """BASELINE config 5: 65 536 participants over 8 GPUs = 8192 per GPU, both groupings"""
    for sizes in ([1024] * 64, [16] * 4096):
        for placement in ("striped", "local"):
            plan = ConferencePlan(sizes, 8, placement)
            assert plan.total == 65536 and all(plan.local_count(r) == 8192 for r in range(8))


def test_plan_and_mode_errors():
    with pytest.raises(ValueError):
        ConferencePlan([], 2)
    with pytest.raises(ValueError):
        ConferencePlan([4, -1], 2)
    with pytest.raises(ValueError):
        ConferencePlan([4], 0)
    with pytest.raises(ValueError):
        ConferencePlan([4], 2, "ring")
    from tests._conf import NumpyBackend

    with pytest.raises(ValueError):
        ShardedConference(ConferencePlan([8, 8], 2, "striped"), 0, mode="local", backend=NumpyBackend())
    with pytest.raises(ValueError):
        ShardedConference(ConferencePlan([8, 8], 2), 0, mode="smoke-signals", backend=NumpyBackend())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, sizes, placement, mode, law, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from tests._conf import NumpyBackend, conference_oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = ConferencePlan(sizes, world, placement)
        frame = 80
        rng = np.random.default_rng(1234)                      # same legs on every rank
        T = 3
        legs = (rng.integers(0, 256, (T, plan.total, frame)).astype(np.uint8) if law >= 0
                else rng.integers(-32768, 32768, (T, plan.total, frame)).astype(np.int16))
        mine = plan.local_members(rank)
        conf = ShardedConference(plan, rank, law=law, freq=8000, mode=mode, backend=NumpyBackend())
        ok = True
        for t in range(T):
            d_in = torch.from_numpy(np.ascontiguousarray(legs[t][mine]))
            d_out = torch.empty_like(d_in)
            d_bus = torch.zeros((plan.n_conf, frame), dtype=torch.int32)
            conf.tick(d_in, d_out, d_bus)
            bus, out = conference_oracle(law, legs[t], plan.global_start)
            ok &= np.array_equal(d_out.numpy(), out[mine])
            if mode == "nccl":
                ok &= np.array_equal(d_bus.numpy(), bus)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("sizes,placement,mode,law", [
    ([1, 2, 3, 58, 7, 0, 5], "striped", "nccl", 0),
    ([16] * 12, "striped", "nccl", 1),
    ([9, 4, 30], "striped", "nccl", -1),
    ([16] * 12, "local", "local", 0),
    ([16] * 12, "local", "nccl", 0),
])
def test_world2_gloo_exchange_matches_single_bus(sizes, placement, mode, law):
    """two ranks over gloo: all-reduce wiring of the sharded bus == one global bus (bit for bit)"""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, sizes, placement, mode, law, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
