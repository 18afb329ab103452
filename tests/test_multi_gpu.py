"""Cross-GPU conference bus on real devices (BASELINE config 5).  The single-device cases run in the
normal `-m gpu` pass; the two-device cases need `gpurun --gpus 2` and skip themselves otherwise."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from tests._conf import conference_oracle  # noqa: E402
from wmix_b200.conference import ConferencePlan, CudaBackend, ShardedConference  # noqa: E402


def _legs(law, T, total, frame, seed):
    rng = np.random.default_rng(seed)
    if law >= 0:
        return rng.integers(0, 256, (T, total, frame)).astype(np.uint8)
    x = rng.integers(-32768, 32768, (T, total, frame)).astype(np.int16)
    x[:, ::3] //= 4                                     # mix of clipping and non-clipping sums
    return x


@pytest.mark.parametrize("law,freq,sizes", [(0, 8000, [1, 2, 3, 58, 0, 700]), (1, 8000, [16] * 300), (-1, 16000, [5, 1200, 33])])
def test_peer_bus_world1_matches_oracle(law, freq, sizes):
    """one rank, no peers: the fused kernel and the C-side NCCL bus (wmixb_nccl_bus_*, which needs no communicator here)"""
    plan = ConferencePlan(sizes, 1)
    frame = freq // 100
    legs = _legs(law, 4, plan.total, frame, 5)
    for mode in ("peer", "nccl_c"):
        conf = ShardedConference(plan, 0, law=law, freq=freq, mode=mode, device=0)
        d_bus = torch.empty((plan.n_conf, frame), dtype=torch.int32, device="cuda:0")
        for t in range(4):
            d_in = torch.from_numpy(legs[t]).to("cuda:0")
            d_out = torch.empty_like(d_in)
            conf.tick(d_in, d_out, d_bus)
            bus, out = conference_oracle(law, legs[t], plan.global_start)
            assert np.array_equal(d_bus.cpu().numpy(), bus), mode
            assert np.array_equal(d_out.cpu().numpy(), out), mode
        assert conf.status() == 0
        conf.close()


def _run_local_ranks(devices, law, sizes, opts=None, T=5):
    """several ranks inside this process (wmixb_peer_bus_connect_local), one stream per rank"""
    world = len(devices)
    opts = dict(opts or {})
    opts.setdefault("ranks_per_device", max(devices.count(d) for d in set(devices)))
    plan = ConferencePlan(sizes, world, "striped")
    frame = 80
    legs = _legs(law, T, plan.total, frame, 11)
    backs = []
    for r, dev in enumerate(devices):
        b = CudaBackend(plan.local_count(r), 8000, dev)
        b.set_conferences(plan.local_conf_start(r))
        b.peer_create(r, world, opts)
        backs.append(b)
    for b in backs:
        b.peer_connect_local(backs)
    streams = [torch.cuda.Stream(device=d) for d in devices]
    for t in range(T):
        ins, outs, buses = [], [], []
        for r, dev in enumerate(devices):
            with torch.cuda.device(dev):
                d_in = torch.from_numpy(np.ascontiguousarray(legs[t][plan.local_members(r)])).to("cuda:%d" % dev)
                ins.append(d_in)
                outs.append(torch.empty_like(d_in))
                buses.append(torch.empty((plan.n_conf, frame), dtype=torch.int32, device="cuda:%d" % dev))
        for d in set(devices):
            torch.cuda.synchronize(d)
        for r in range(world):                          # asynchronous launches: the kernels meet on the device(s)
            backs[r].peer_tick(law, ins[r], outs[r], buses[r], streams[r])
        for d in set(devices):
            torch.cuda.synchronize(d)
        bus, out = conference_oracle(law, legs[t], plan.global_start)
        for r in range(world):
            assert np.array_equal(buses[r].cpu().numpy(), bus), "tick %d rank %d bus" % (t, r)
            assert np.array_equal(outs[r].cpu().numpy(), out[plan.local_members(r)]), "tick %d rank %d legs" % (t, r)
    for b in backs:
        assert b.peer_status() == 0
    for b in backs:
        b.close()


def test_peer_bus_two_ranks_on_one_device():
    """the multi-rank protocol (mailboxes, flags, parity double-buffering) with both ranks on cuda:0:
    few conferences, so both persistent grids are resident together"""
    _run_local_ranks([0, 0], 0, [40, 7, 64, 1, 9, 200])
    _run_local_ranks([0, 0, 0], 1, [33] * 10)


@pytest.mark.parametrize("opts", [{"tile": "row"}, {"tile": "row", "reduce_scatter": 1}])
def test_peer_bus_exchange_variants_on_one_device(opts):
    """the shapes the fused kernel takes for MANY conferences — one tile per bus row, and the reduce-scatter / all-gather
    exchange it uses from 4 ranks up — forced here on 2, 3 and 4 ranks sharing cuda:0 (wmixb_peer_opts at
    creation; normally they follow from n_conf and world alone)"""
    _run_local_ranks([0, 0], 0, [40, 7, 64, 1, 9, 200], opts)
    _run_local_ranks([0, 0, 0], 1, [33] * 10, dict(opts, ranks_per_device=3))
    _run_local_ranks([0, 0, 0, 0], -1, [5, 0, 17, 3, 1, 1, 2], dict(opts, ranks_per_device=4))


def test_peer_bus_missing_peer_times_out_instead_of_hanging():
    plan = ConferencePlan([8, 8], 2, "striped")
    backs = []
    for r in range(2):
        b = CudaBackend(plan.local_count(r), 8000, 0)
        b.set_conferences(plan.local_conf_start(r))
        b.peer_create(r, 2, {"timeout_ms": 30})
        backs.append(b)
    for b in backs:
        b.peer_connect_local(backs)
    d_in = torch.zeros((plan.local_count(0), 80), dtype=torch.uint8, device="cuda:0")
    d_out = torch.empty_like(d_in)
    backs[0].peer_tick(0, d_in, d_out, None, None)  # rank 1 never ticks
    assert backs[0].peer_status() == 2               # gave up waiting for rank 1
    for b in backs:
        b.close()


needs2 = pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")


@needs2
def test_peer_bus_two_devices_one_process():
    _run_local_ranks([0, 1], 0, [1024] * 8)
    _run_local_ranks([0, 1], -1, [16] * 512 + [3, 0, 1])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _nccl_worker(rank, world, port, sizes, law, q, opts=None):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from tests._conf import conference_oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ok = True
    try:
        plan = ConferencePlan(sizes, world, "striped")
        frame, T = 80, 6
        legs = _legs(law, T, plan.total, frame, 21)
        mine = plan.local_members(rank)
        dev = "cuda:%d" % rank
        results = {}
        for mode in ("peer", "nccl", "nccl_c"):
            conf = ShardedConference(plan, rank, law=law, freq=8000, mode=mode, device=rank, peer_opts=opts if mode == "peer" else None)
            d_bus = torch.empty((plan.n_conf, frame), dtype=torch.int32, device=dev)
            outs = []
            for t in range(T):
                d_in = torch.from_numpy(np.ascontiguousarray(legs[t][mine])).to(dev)
                d_out = torch.empty_like(d_in)
                conf.tick(d_in, d_out, d_bus)
                torch.cuda.synchronize()
                bus, out = conference_oracle(law, legs[t], plan.global_start)
                ok &= np.array_equal(d_bus.cpu().numpy(), bus) and np.array_equal(d_out.cpu().numpy(), out[mine])
                outs.append(d_out.cpu().numpy())
            ok &= conf.status() == 0
            results[mode] = outs
            dist.barrier()
            conf.close()
        ok &= all(np.array_equal(a, b) for a, b in zip(results["peer"], results["nccl"]))
        ok &= all(np.array_equal(a, b) for a, b in zip(results["peer"], results["nccl_c"]))
        # soak: many back-to-back ticks (no host synchronisation in between, both mailbox parities, ranks drifting apart), the
        # fused exchange against the NCCL one on the same legs, compared on the device every tick
        a = ShardedConference(plan, rank, law=law, freq=8000, mode="peer", device=rank, peer_opts=opts)
        b = ShardedConference(plan, rank, law=law, freq=8000, mode="nccl_c", device=rank)
        g = torch.Generator(device="cpu").manual_seed(77 + rank)
        n_local = len(mine)
        pool = (torch.randint(0, 256, (16, n_local, frame), generator=g, dtype=torch.uint8) if law >= 0
                else torch.randint(-32768, 32768, (16, n_local, frame), generator=g, dtype=torch.int16)).to(dev)
        out = torch.empty_like(pool[0])
        bus = torch.empty((plan.n_conf, frame), dtype=torch.int32, device=dev)
        w_out = torch.randint(1, 1 << 20, out.shape, generator=g, dtype=torch.int64).to(dev)
        w_bus = torch.randint(1, 1 << 20, bus.shape, generator=g, dtype=torch.int64).to(dev)
        sums = torch.zeros((2, 400), dtype=torch.int64, device=dev)
        for k, conf in enumerate((b, a)):                # the fused ticks run alone, nothing else keeps the ranks in step
            for t in range(400):
                conf.tick(pool[t % 16], out, bus)
                sums[k, t] = (out.to(torch.int64) * w_out).sum() + (bus.to(torch.int64) * w_bus).sum()
            torch.cuda.synchronize()
            dist.barrier()
        bad = (sums[0] != sums[1]).sum()
        ok &= int(bad.item()) == 0 and a.status() == 0
        dist.barrier()
        a.close()
        b.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@needs2
@pytest.mark.parametrize("law,sizes,opts", [(0, [1024] * 8, None), (1, [16] * 512, None), (-1, [1, 2, 3, 58, 7, 0, 5], None),
                                            (1, [16] * 512, {"reduce_scatter": 1}),
                                            (0, [3, 40, 1, 0, 9], {"tile": "row", "reduce_scatter": 1})])
def test_two_processes_peer_and_nccl_modes_agree_with_oracle(law, sizes, opts):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, sizes, law, q, opts)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


@needs2
def test_nccl_bus_from_plain_c_two_threads_two_gpus(tmp_path):
    """The NCCL exchange of the conference bus driven from C alone (tests/c/nccl_from_c.c): two host threads, one GPU each,
    wmixb_nccl_bus_* — the two-GPU result must equal the one-GPU result of the same conferences, every tick."""
    import subprocess

    libdir = os.path.join(ROOT, "wmix_b200")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "nccl_from_c")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", cuda + "/include",
                           os.path.join(ROOT, "tests", "c", "nccl_from_c.c"), "-L", libdir, "-lwmix_b200", "-Wl,-rpath," + libdir,
                           "-L", cuda + "/lib64", "-lcudart", "-Wl,-rpath," + cuda + "/lib64", "-lpthread", "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "mismatching blocks = 0" in r.stdout
