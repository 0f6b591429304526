"""The arithmetic of the multi-GPU partition of the H_eff application on ONE device (nsb_shard_emulate): for every rank of a
G-way partition the partial result (reduce-scatter / all-reduce positions, regions swept to the right) or the result slab
(all-gather positions, regions swept to the left) is formed exactly as that rank would form it, the collective is replaced by a
local sum / concatenation, and the result must equal the single-GPU application.  Covers what the 1-GPU test tier cannot reach
through NCCL: rank counts 2 / 4 / 8, uneven bonds, identity channels skipped and not, real and complex, a long chain (the
environments of an N = 100 chain carry O(N) energy terms next to the identity channel)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("chi,nsites,cplx", [(64, 16, False), (96, 100, False), (48, 14, True), (40, 12, False)])
@pytest.mark.parametrize("skip", [1, 0])
def test_sharded_arithmetic_equals_single_gpu(chi, nsites, cplx, skip):
    import networksolvers_b200 as ns
    g = ns.path_graph(nsites)
    sites = ns.siteinds("S=1/2", g)
    H = ns.ttno(ns.heisenberg(g), sites)
    dt = np.complex128 if cplx else np.float64
    ctx = ns.default_context()
    ctx.set_option("skip_identity", skip)
    try:
        mid = nsites // 2
        net = ns.DeviceNetwork.synthetic(H, sites, chi, seed=11, dtype=dt, ctx=ctx, ortho_region=[mid, mid + 1], canonical=True)
        for region, want_mode in (([mid, mid + 1], (1, 3)), ([mid + 1, mid + 2], (1, 3)), ([mid + 2, mid + 1], (2, 3)), ([mid + 1, mid], (2, 3))):
            net.extract(region)
            ref = net.matvec_device(1, download=True)
            for G in (2, 3, 4, 8):
                out, mode = net.shard_emulate(G)
                _, dims = net.local_info()
                if dims[-1] % G == 0 and dims[-1] >= G:
                    assert mode == want_mode[0], (region, G, mode)          # even bond: reduce-scatter / all-gather form
                else:
                    assert mode in (0, 3), (region, G, mode)                 # uneven: all-reduce form (right) or unsharded (left)
                err = np.abs(out - ref).max() / np.abs(ref).max()
                assert err <= 1e-13, (region, G, mode, err)
            net.update_eigsolve()
            net.insert((0.0, 1, chi))
    finally:
        ctx.set_option("skip_identity", 1)
