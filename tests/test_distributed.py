"""World-size-2 gloo test of the multi-GPU window combine (bindsnet_b200/distributed.py) on CPU:
two ranks, each a replica on its batch shard driven by the oracle, one all-reduce of dW + dtheta
per window — checked against the same combination computed in one process from independent
oracle replicas (the multi-GPU oracle of SURVEY.md §8e)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make(B, n=40, seed=3):
    from bindsnet_b200.models import DiehlAndCook2015

    g = torch.Generator().manual_seed(seed)
    net = DiehlAndCook2015(n_inpt=196, n_neurons=n, batch_size=B, inpt_shape=(1, 14, 14), inh=60.0, norm=20.0)
    with torch.no_grad():
        net.connections[("X", "Ae")].w.copy_(0.3 * torch.rand(196, n, generator=g))
        net.layers["Ae"].theta.copy_(0.5 * torch.rand(n, generator=g))
    return net


def _inputs(T=60, B=8, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.bernoulli(0.08 * torch.ones(T, B, 1, 14, 14), generator=g).byte()


def _patch_cpu_combine():
    """The combine kernels are CUDA-only; on CPU ranks route them to the oracle's restatement."""
    import ctypes as C
    from bindsnet_b200 import _backend
    from oracle import oracle

    def prepare(w, w0, dw):
        dw.copy_(w - w0)

    def apply(w, w0, dws, has_clamp, wmin, wmax, has_norm, norm_abs, norm):
        assert oracle.lib().snn_oracle_delta_apply(w.data_ptr(), w0.data_ptr(), dws.contiguous().data_ptr(), w.shape[0],
                                                   w.shape[1], int(has_clamp), float(wmin), float(wmax), int(has_norm),
                                                   int(norm_abs), float(norm)) == 0

    _backend.delta_prepare, _backend.delta_apply = prepare, apply


def _worker(rank, world, port, out, emulated=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bindsnet_b200.distributed import ShardedWindowRunner
    from oracle.oracle import OracleBackend

    if emulated:   # the ranks run the library's CUDA sources (fused window with the delta epilogue, in-place apply kernel)
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import emu
        Backend = emu.EmuBackend
    else:
        _patch_cpu_combine()
        Backend = OracleBackend
    x = _inputs()
    shard = x[:, rank * 4:(rank + 1) * 4]
    net = _make(4)
    with Backend():
        runner = ShardedWindowRunner(net)
        runner._emulated = emulated
        runner.run({"X": shard}, time=60, one_spike_seed=17 + rank)
        net.reset_state_variables()
        runner.run({"X": shard}, time=60, one_spike_seed=27 + rank)
        assert not emulated or runner._delta_windows is True, "the fused delta-window path was not taken"
    torch.save({"w": net.connections[("X", "Ae")].w.detach().clone(), "theta": net.layers["Ae"].theta.clone()},
               os.path.join(out, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("emulated", [False, True])
def test_two_rank_window_combine_matches_replica_oracle(tmp_path, emulated):
    """emulated = True: the CPU twin of the 2-GPU NCCL test in tests/test_gpu_ops.py — each rank runs the fused kernel's
    delta window and the in-place apply kernel (csrc/snn_combine.cuh) under the CUDA-model emulation, gloo does the all-reduce."""
    port = 29500 + (os.getpid() % 1000) + (1000 if emulated else 0)
    mp.spawn(_worker, args=(2, port, str(tmp_path), emulated), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt"); r1 = torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["w"], r1["w"]) and torch.equal(r0["theta"], r1["theta"]), "ranks diverged"

    # single-process oracle of the same procedure: independent replicas + one exchange per window
    from oracle.oracle import OracleBackend
    from oracle import oracle

    x = _inputs()
    nets = [_make(4), _make(4)]
    w = nets[0].connections[("X", "Ae")].w.detach().clone()
    th = nets[0].layers["Ae"].theta.clone()
    with OracleBackend():
        for window, seed0 in enumerate((17, 27)):
            dws, dths = [], []
            for r, net in enumerate(nets):
                with torch.no_grad():
                    net.connections[("X", "Ae")].w.copy_(w); net.layers["Ae"].theta.copy_(th)
                if window:
                    net.reset_state_variables()
                net.run({"X": x[:, r * 4:(r + 1) * 4]}, time=60, one_spike_seed=seed0 + r, b200_normalize=False)
                dws.append(net.connections[("X", "Ae")].w.detach() - w)
                dths.append(net.layers["Ae"].theta - th)
            dsum = (dws[0] + dws[1]).contiguous()
            wn = torch.empty_like(w)
            assert oracle.lib().snn_oracle_delta_apply(wn.data_ptr(), w.data_ptr(), dsum.data_ptr(), 196, 40, 1, 0.0, 1.0, 1, 0, 20.0) == 0
            w, th = wn, th + (dths[0] + dths[1])
    assert np.array_equal(r0["w"].numpy(), w.numpy())
    assert np.array_equal(r0["theta"].numpy(), th.numpy())


# ---- learned Conv2dConnection weights (config 4's graph in small): sum + clamp by the combine kernel on the flattened filters, then the
# connection's own per-filter normalize -------------------------------------------------------------------------------------------------
def _conv_make(B):
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import cases

    ns = cases.namespace("b200")
    return cases._conv_net(ns, B, 40, (2, 9, 9), 3, 3, 2, 1, ns.learning.MSTDP, (131, 132), n_out=5, norm=1.5)


def _conv_inputs():
    g = torch.Generator().manual_seed(77)
    return torch.bernoulli(0.15 * torch.ones(40, 8, 2, 9, 9), generator=g).byte()


def _conv_worker(rank, world, port, out, emulated=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bindsnet_b200.distributed import ShardedWindowRunner
    from oracle.oracle import OracleBackend

    if emulated:   # the generic window kernel, delta_prepare / delta_apply and the conv normalize operator from their CUDA sources
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import emu
        Backend = emu.EmuBackend
    else:
        _patch_cpu_combine()
        Backend = OracleBackend
    shard = _conv_inputs()[:, rank * 4:(rank + 1) * 4]
    net = _conv_make(4)
    with Backend():
        runner = ShardedWindowRunner(net)
        for window in range(2):
            if window:
                net.reset_state_variables()
            runner.run({"X": shard}, time=40, reward=0.8 + 0.3 * window)
    torch.save({f"{s}->{t}": c.w.detach().clone() for (s, t), c in net.connections.items()}, os.path.join(out, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("emulated", [False, True])
def test_two_rank_combine_of_learned_conv_weights_matches_replica_oracle(tmp_path, emulated):
    port = 31500 + (os.getpid() % 1000) + (1000 if emulated else 0)
    mp.spawn(_conv_worker, args=(2, port, str(tmp_path), emulated), nprocs=2, join=True)
    r0 = torch.load(tmp_path / "rank0.pt"); r1 = torch.load(tmp_path / "rank1.pt")
    assert all(torch.equal(r0[k], r1[k]) for k in r0), "ranks diverged"

    from oracle.oracle import OracleBackend

    x = _conv_inputs()
    nets = [_conv_make(4), _conv_make(4)]
    keys = list(nets[0].connections)
    w = {k: nets[0].connections[k].w.detach().clone() for k in keys}
    with OracleBackend():
        for window in range(2):
            sums = {k: torch.zeros_like(w[k]) for k in keys}
            for r, net in enumerate(nets):
                with torch.no_grad():
                    for k in keys:
                        net.connections[k].w.copy_(w[k])
                if window:
                    net.reset_state_variables()
                net.run({"X": x[:, r * 4:(r + 1) * 4]}, time=40, reward=0.8 + 0.3 * window, b200_normalize=False)
                for k in keys:
                    sums[k] += net.connections[k].w.detach() - w[k]
            for k in keys:
                c = nets[0].connections[k]
                with torch.no_grad():
                    c.w.copy_(torch.clamp(w[k] + sums[k], float(c.wmin), float(c.wmax)))
                if c.norm is not None:
                    c.normalize()
                w[k] = c.w.detach().clone()
    for (s, t) in keys:
        assert np.array_equal(r0[f"{s}->{t}"].numpy(), w[(s, t)].numpy()), (s, t)
    assert not torch.equal(w[keys[0]], _conv_make(4).connections[keys[0]].w)      # something was learnt
