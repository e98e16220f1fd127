"""Single kernels behind the C ABI against the oracle on the GPU (need a B200): the multi-GPU window-combine
kernels (``snn_b200_delta_prepare / delta_apply``), the Conv2dConnection single operators, the device-side check
of the static-matrix structure hints, and the sharded window runner on two real NCCL ranks."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cases
import helpers

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape,clamp,norm,norm_abs", [((784, 1600), True, 78.4, 0), ((100, 37), True, None, 1), ((53, 129), False, 5.0, 1)])
def test_delta_prepare_apply_bit_exact_vs_oracle(shape, clamp, norm, norm_abs):
    """dW = W - W0 and W = clamp(W0 + sum dW) + normalize(): CUDA == oracle, bit for bit."""
    from bindsnet_b200 import _backend
    from oracle import oracle

    g = torch.Generator().manual_seed(11)
    w0 = (0.3 * torch.rand(*shape, generator=g)).contiguous()
    w = (w0 + 0.05 * torch.randn(*shape, generator=g)).clamp(0, 1).contiguous()
    other = 0.05 * torch.randn(*shape, generator=g)
    # prepare
    dw_gpu = torch.empty_like(w, device="cuda")
    _backend.delta_prepare(w.cuda(), w0.cuda(), dw_gpu)
    assert np.array_equal(dw_gpu.cpu().numpy(), (w - w0).numpy())
    # apply (what every rank does with the all-reduced sum)
    dsum = ((w - w0) + other).contiguous()
    out_gpu = torch.empty_like(w, device="cuda")
    _backend.delta_apply(out_gpu, w0.cuda(), dsum.cuda(), clamp, 0.0, 1.0, norm is not None, norm_abs, norm or 0.0)
    out_cpu = torch.empty_like(w)
    assert oracle.lib().snn_oracle_delta_apply(out_cpu.data_ptr(), w0.data_ptr(), dsum.data_ptr(), shape[0], shape[1], int(clamp),
                                               C.c_float(0.0), C.c_float(1.0), int(norm is not None), norm_abs, C.c_float(norm or 0.0)) == 0
    a, b = out_gpu.cpu().numpy(), out_cpu.numpy()
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"max |d| {np.abs(a - b).max():.3e}"


def test_conv2d_single_operators_bit_exact_vs_oracle():
    """Conv2dConnection.compute / normalize through the C ABI (topology.py:799-815, 824-837) equal the oracle's and,
    within fp32 summation tolerance, torch's conv2d."""
    from bindsnet_b200.network import nodes, topology
    from oracle.oracle import OracleBackend

    res = []
    for dev in ("cuda", "cpu"):
        g = torch.Generator().manual_seed(4)
        X = nodes.Input(shape=[2, 11, 9], traces=True); H = nodes.LIFNodes(shape=[5, 6, 5], traces=True)
        c = topology.Conv2dConnection(X, H, kernel_size=(3, 3), stride=2, padding=1, wmin=-1.0, wmax=1.0, norm=0.4,
                                      w=torch.rand(5, 2, 3, 3, generator=g) - 0.3, b=torch.rand(5, generator=g))
        s = torch.bernoulli(0.3 * torch.ones(7, 2, 11, 9), generator=g).byte()
        c.to(dev)
        ctx = OracleBackend() if dev == "cpu" else None
        if ctx: ctx.__enter__()
        out = c.compute(s.to(dev))
        c.normalize()
        if ctx: ctx.__exit__(None, None, None)
        res.append((out.cpu().numpy(), c.w.detach().cpu().numpy(), s, c))
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    g = torch.Generator().manual_seed(4)
    w0 = torch.rand(5, 2, 3, 3, generator=g) - 0.3; b0 = torch.rand(5, generator=g)
    ref = torch.nn.functional.conv2d(res[0][2].float(), w0, b0, stride=2, padding=1)
    assert np.allclose(res[0][0], ref.numpy(), atol=1e-5)


def test_static_structure_is_verified_on_the_device():
    """The fused kernels replace DiehlAndCook2015's exc / inh matrices by their constants (structure hints verified
    on the host and cached).  A matrix modified behind that cache — `.data` writes do not bump the version counter —
    must be detected on the device, not silently ignored."""
    from bindsnet_b200 import _backend
    from bindsnet_b200.models import DiehlAndCook2015

    net = DiehlAndCook2015(n_inpt=784, n_neurons=64, batch_size=4, inpt_shape=(1, 28, 28), inh=120.0).to("cuda")
    x = torch.bernoulli(0.05 * torch.ones(20, 4, 1, 28, 28)).byte().cuda()
    net.run({"X": x}, time=20)
    net.check_errors()
    assert _backend.last_tier in (2, 3)
    net.connections[("Ai", "Ae")].w.data[3, 5] = -1.0   # no longer constant off the diagonal
    net.run({"X": x}, time=20)
    with pytest.raises(_backend.BackendError):
        net.check_errors()
    # the cache was dropped: the next window re-verifies on the host and takes the generic kernel
    net.run({"X": x}, time=20)
    net.check_errors()
    assert _backend.last_tier == 1


_NCCL_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
import torch, torch.distributed as dist
rank = int(os.environ["RANK"]); torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
import test_distributed as td
from bindsnet_b200.distributed import ShardedWindowRunner
x = td._inputs()
shard = x[:, rank * 4:(rank + 1) * 4].cuda()
net = td._make(4).to("cuda")
runner = ShardedWindowRunner(net)
runner.run({{"X": shard}}, time=60, one_spike_seed=17 + rank)
net.reset_state_variables()
runner.run({{"X": shard}}, time=60, one_spike_seed=27 + rank)
net.check_errors()
assert runner._delta_windows is True, "the fused delta window (epilogue writes dW / dtheta) was not taken"
torch.save({{"w": net.connections[("X", "Ae")].w.detach().cpu(), "theta": net.layers["Ae"].theta.cpu()}}, os.path.join({out!r}, f"rank{{rank}}.pt"))
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_runner_on_two_nccl_ranks_matches_replica_oracle(tmp_path):
    """ShardedWindowRunner on two real NCCL ranks (CUDA delta_prepare / delta_apply, one all-reduce per window)
    against the single-process combination of independent oracle replicas — the same reference the gloo test uses."""
    import test_distributed as td
    from oracle import oracle
    from oracle.oracle import OracleBackend

    script = tmp_path / "worker.py"
    script.write_text(_NCCL_WORKER.format(root=ROOT, out=str(tmp_path)))
    port = 29600 + (os.getpid() % 300)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                    "--master-port", str(port), str(script)], check=True, timeout=600)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0["w"], r1["w"]) and torch.equal(r0["theta"], r1["theta"]), "ranks diverged"
    x = td._inputs()
    nets = [td._make(4), td._make(4)]
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
            assert oracle.lib().snn_oracle_delta_apply(wn.data_ptr(), w.data_ptr(), dsum.data_ptr(), 196, 40, 1, C.c_float(0.0), C.c_float(1.0), 1, 0,
                                                       C.c_float(20.0)) == 0
            w, th = wn, th + (dths[0] + dths[1])
    assert np.array_equal(r0["w"].numpy(), w.numpy())
    assert np.array_equal(r0["theta"].numpy(), th.numpy())


def test_readout_kernels_match_the_torch_formulas():
    """snn_b200_assign_labels / snn_b200_predict (csrc/snn_readout.cu) against bindsnet_b200.evaluation's CPU formulas
    (which tests/test_evaluation.py pins against the live reference), from SpikeCounter-style int32 counts."""
    from bindsnet_b200 import evaluation as ev

    g = torch.Generator().manual_seed(8)
    S, n, L = 128, 1600, 10
    counts = torch.poisson(3.0 * torch.rand(S, n, generator=g), generator=g).int()
    labels = torch.randint(0, L, (S,), generator=g)
    a_c, p_c, r_c = ev.assign_labels(counts, labels, L)
    a_g, p_g, r_g = ev.assign_labels(counts.cuda(), labels.cuda(), L)
    assert torch.equal(a_g.cpu(), a_c) and torch.allclose(p_g.cpu(), p_c, atol=1e-6) and torch.allclose(r_g.cpu(), r_c, atol=1e-5)
    a2_c, p2_c, r2_c = ev.assign_labels(counts.flip(0), labels, L, rates=r_c.clone(), alpha=0.8)
    a2_g, p2_g, r2_g = ev.assign_labels(counts.flip(0).cuda(), labels.cuda(), L, rates=r_g.clone(), alpha=0.8)
    assert torch.equal(a2_g.cpu(), a2_c) and torch.allclose(r2_g.cpu(), r2_c, atol=1e-5)
    assert torch.equal(ev.all_activity(counts.cuda(), a_g, L).cpu(), ev.all_activity(counts, a_c, L))
    pw_g, pw_c = ev.proportion_weighting(counts.cuda(), a_g, p_g, L).cpu(), ev.proportion_weighting(counts, a_c, p_c, L)
    assert (pw_g != pw_c).sum() <= 1   # weighted float sums: a near-tie may fall the other way


def test_delta_window_writes_the_change_and_leaves_the_weights():
    """snn_run_opts_t.delta_w / delta_theta (fused DiehlAndCook2015 kernel): the window writes W_end - W_start and
    theta_end - theta_start into the caller's buffer and leaves W / theta untouched; snn_b200_delta_apply_fused then
    equals the snapshot-based combine, bit for bit."""
    from bindsnet_b200 import _backend
    from bindsnet_b200.models import DiehlAndCook2015

    def make():
        torch.manual_seed(3)
        return DiehlAndCook2015(n_inpt=784, n_neurons=96, batch_size=8, inpt_shape=(1, 28, 28), norm=78.4, theta_plus=0.05).to("cuda")

    x = torch.bernoulli(0.04 * torch.ones(80, 8, 1, 28, 28), generator=torch.Generator().manual_seed(5)).byte().cuda()
    a, b = make(), make()
    w0, th0 = a.connections[("X", "Ae")].w.detach().clone(), a.layers["Ae"].theta.clone()
    a.run({"X": x}, time=80, one_spike_seed=9, b200_normalize=False)
    a.check_errors()
    assert _backend.last_tier == 2
    wb, thb = b.connections[("X", "Ae")].w.detach(), b.layers["Ae"].theta
    flat = torch.full((wb.numel() + thb.numel(),), float("nan"), device="cuda")
    dw, dth = flat[:wb.numel()].view_as(wb), flat[wb.numel():]
    b.run({"X": x}, time=80, one_spike_seed=9, b200_normalize=False, b200_delta=(dw, dth))
    b.check_errors()
    assert _backend.last_tier == 2
    assert torch.equal(wb, w0) and torch.equal(thb, th0), "a delta window must not touch W / theta"
    wa, tha = a.connections[("X", "Ae")].w.detach(), a.layers["Ae"].theta
    assert torch.equal(dw, wa - w0) and torch.equal(dth, tha - th0)
    assert float(dw.abs().sum()) > 0 and float(dth.abs().sum()) > 0, "nothing learned: nothing tested"
    for lname in ("Ae", "Ai"):   # everything else is written back as usual
        assert torch.equal(a.layers[lname].v, b.layers[lname].v) and torch.equal(a.layers[lname].s, b.layers[lname].s)
    # combine: in place == snapshot based
    ref_w = torch.empty_like(w0)
    _backend.delta_apply(ref_w, w0, (2.0 * dw).contiguous(), True, 0.0, 1.0, True, 0, 78.4)
    two = (2.0 * flat).contiguous()
    _backend.delta_apply_fused(wb, two[:wb.numel()].view_as(wb), True, 0.0, 1.0, True, 0, 78.4, theta=thb, dtheta_sum=two[wb.numel():])
    assert torch.equal(wb, ref_w) and torch.equal(thb, th0 + 2.0 * dth)
    # a graph the fused kernel does not take has no delta window
    c = make(); c.force_tier = 1
    with pytest.raises(_backend.BackendError):
        c.run({"X": x}, time=80, one_spike_seed=9, b200_normalize=False, b200_delta=(dw, dth))
