"""GEMM / factorisation timing probe (run under gpurun)."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns

ctx = ns.default_context()
out = []
print("dmma peak", ctx.dmma_peak_tflops())
for chi in (512, 1024, 2048, 4096):
    for name, (m, n, k, oa, ob) in {"K1_TN": (5 * chi, 4 * chi, chi, "T", "N"), "K3_NN": (4 * chi, chi, 5 * chi, "N", "N"),
                                    "envK3_TN": (5 * chi, chi, 2 * chi, "T", "N"), "R_NT": (5 * chi, chi, 2 * chi, "N", "T")}.items():
        for impl in (2, 3):
            ms = ctx.gemm_bench(m, n, k, oa, ob, impl=impl, reps=3)
            tf = 2.0 * m * n * k / ms * 1e-9
            rec = dict(bench="gemm", chi=chi, name=name, impl=impl, m=m, n=n, k=k, ms=ms, tflops=tf)
            print(json.dumps(rec)); out.append(rec)
for chi in (1024,):
    m, n, k = 5 * chi, 4 * chi, chi
    for impl in (2, 3):
        ms = ctx.gemm_bench(m, n, k, "C", "N", dtype=np.complex128, impl=impl, reps=3)
        rec = dict(bench="zgemm", chi=chi, impl=impl, ms=ms, real_tflops=8.0 * m * n * k / ms * 1e-9)
        print(json.dumps(rec)); out.append(rec)
rng = np.random.default_rng(0)
for n in (128, 256, 512, 1024):
    M = rng.standard_normal((n, n))
    t0 = time.perf_counter()
    U, C, spec, info = ctx.factorize(M, cutoff=0.0, maxdim=n // 2)
    dt = time.perf_counter() - t0
    rec = dict(bench="factorize", n=n, s=dt, sweeps=info["sweeps"])
    print(json.dumps(rec)); out.append(rec)
json.dump(out, open("gpurun_out/perf_probe.json", "w"))
