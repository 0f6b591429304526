"""Single-process multi-device entry points (nsb_multi_*, include/nsb200.h): the calls a single-threaded host -- the Julia
shim -- makes to drive several GPUs.  With one device the fan-out path is exercised on any box; with two or more the replicas
run the sharded region step and must reproduce the single-GPU sweep (energies 1e-10, truncation errors 1e-8)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("ndev", [1, 2])
def test_multi_device_region_steps_match_single_device(ndev):
    if _ndev() < ndev:
        pytest.skip(f"needs {ndev} GPUs")
    import networksolvers_b200 as ns
    from networksolvers_b200.parallel import MultiDeviceNetwork
    g = ns.path_graph(14)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    psi = ns.random_state(sites, 64, seed=3)
    ref = ns.EigsolveProblem(state=psi, operator=H).net
    multi = MultiDeviceNetwork(H, psi, devices=list(range(ndev)))
    regions = [[7, 8], [8, 9], [9, 8], [8, 7]]
    for reg in regions:
        out = []
        for net in (ref, multi):
            net.extract(reg)
            val, info = net.update_eigsolve()
            th, _ = net.local_download()
            ins = net.insert((1e-10, 1, 48))
            out.append((val, info.nmatvec, np.linalg.norm(th), ins.newdim, ins.truncerr))
        a, b = out
        assert abs(a[0] - b[0]) <= 1e-10 * max(1.0, abs(a[0])), (reg, a, b)
        assert a[1] == b[1] == 3 and abs(a[2] - b[2]) < 1e-12
        assert a[3] == b[3] and abs(a[4] - b[4]) <= 1e-8
    assert multi.maxlinkdim() == ref.maxlinkdim()
    if ndev > 1:
        assert multi.shard_active
    multi.close()
