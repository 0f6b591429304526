"""CPU oracle for the NetworkSolvers.jl sweep hot path  --  TEST INFRASTRUCTURE ONLY.

This package is a NumPy restatement of the reference algorithm (emstoudenmire/
NetworkSolvers, Julia).  It exists to *check* the CUDA path; it is never the
thing shipped or measured as the product.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it.  `networksolvers_b200/` must never import it.

PARITY STATUS: **parity unpinned against an executed reference.**  The
reference is Julia and all of its arithmetic lives in un-vendored packages
(ITensors / NDTensors / ITensorNetworks / KrylovKit, Project.toml:6-27, no
Manifest); neither Julia nor those packages exist in this environment, so the
reference cannot be run to generate vectors.  The oracle is pinned instead on
  * the one numeric constant the reference ships: E0 = -12.8945601 for the S=1
    N=10 Heisenberg chain (examples/dmrg.jl:41-43),
  * the quantities the reference's own tests assert: |E_dmrg - E_ED| < 1e-5 on
    the 10-vertex tree (test/dmrg/test_tree_dmrg.jl:53,67), TDVP norm / overlap /
    phase checks (test/tdvp/test_tree_tdvp.jl:65-77), Euler-tour structure
    (test/test_euler_tour.jl:13-25),
  * exact diagonalisation / dense expm computed independently (oracle/ed.py),
  * mathematics that does not depend on any restatement: H_eff = B^dagger H B from the dense
    Hamiltonian for optimal_map / operator_map / the on-edge map (tests/test_oracle_heff_definition.py);
    Lanczos = Rayleigh-Ritz on the Krylov space, RK2 / RK4 = Taylor polynomials, exponentiate = dense expm
    (tests/test_oracle_local_solvers.py); hand-worked answers of the truncation / docut / expansion-size
    rules (tests/test_oracle_truncation.py); invariants of the subspace expansion
    (tests/test_oracle_expansion_properties.py).
These pin the arithmetic; what stays unpinned is agreement with the *choices* of the un-vendored packages
where they are conventions rather than mathematics (gauge signs, tie-breaking in the truncation at
degenerate values, KrylovKit's stopping tests), restated from their published behaviour.
Each function cites the reference file:line it follows; rules of the un-vendored
upstream packages are cited as "UPSTREAM" with the SURVEY.md appendix entry.
"""
from .tensor import Tensor, contract, prime, noprime, dag  # noqa: F401
