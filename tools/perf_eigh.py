"""Times the Gram + eigh factorisation route on device-resident data (nsb_factorize_host includes the H2D/D2H of the
matrices; the [eigh] phase lines on stderr with NSB_DEBUG_EIGH=1 are device times).
    NSB_DEBUG_EIGH=1 python tools/perf_eigh.py 2048 4096 8192"""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
ctx = ns.default_context()
rng = np.random.default_rng(0)
for a in [a for a in sys.argv[1:] if "=" in a]:   # context options, e.g. eigh_sym_tc=16
    k_, v_ = a.split("="); ctx.set_option(k_, int(v_))
sizes = [int(a) for a in sys.argv[1:] if a.isdigit()] or [1024, 2048, 4096]
NOCHECK = "nocheck" in sys.argv   # skip the host LAPACK SVD (minutes at n = 8192)
kinds = [a for a in sys.argv[1:] if not a.isdigit() and a != "nocheck" and "=" not in a] or ["gauss", "graded"]
for n in sizes:
    for kind in kinds:
        if kind == "gauss":
            M = rng.standard_normal((n, n))
        else:   # DMRG-like: exponentially decaying singular values over 12 decades
            Q1, _ = np.linalg.qr(rng.standard_normal((n, n)))
            Q2, _ = np.linalg.qr(rng.standard_normal((n, n)))
            M = (Q1 * np.exp(-28.0 * np.arange(n) / n)) @ Q2.T
        ctx.factorize(M, cutoff=0.0, maxdim=n // 2)   # warm-up: grows the memory pool (first-touch cost is not the kernels')
        t0 = time.perf_counter()
        ctx.reset_counters()
        U, C, spec, info = ctx.factorize(M, cutoff=0.0, maxdim=n // 2)
        dt = time.perf_counter() - t0
        c = ctx.counters()
        k = info["newdim"]
        ortho = np.abs(U.T @ U - np.eye(k)).max()
        terr_ref = None
        if not NOCHECK:
            s = np.linalg.svd(M, compute_uv=False)
            terr_ref = np.sum(s[k:] ** 2) / np.sum(s ** 2)
        rec = np.linalg.norm(U @ C - M) ** 2 / np.linalg.norm(M) ** 2
        print(json.dumps(dict(bench="factorize_eigh", kind=kind, n=n, s=dt, newdim=k, gemm_tflop=c["gemm_flops"] / 1e12,
                              launches=c["kernel_launches"], ortho_err=ortho, truncerr=info["truncerr"], truncerr_lapack=terr_ref,
                              recon_err2=rec)), flush=True)
