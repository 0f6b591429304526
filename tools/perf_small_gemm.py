"""Small GEMMs in the launch-bound regime: time per call of a stream of back-to-back calls (device events around 50 calls:
whichever of host issue rate and device time is slower) for the three implementations -- naive, DMMA + cp.async, DMMA + TMA
(tensor maps encoded on the host per call).  Decides the size thresholds of GEMM_AUTO."""
import sys, json
sys.path.insert(0, ".")
import networksolvers_b200 as ns
ctx = ns.default_context()
shapes = [(8, 32, 8), (16, 64, 16), (32, 128, 32), (49, 196, 196), (64, 256, 64), (100, 400, 100), (128, 512, 128), (166, 664, 166),
          (256, 1024, 256), (320, 1280, 320), (512, 2048, 512), (1024, 4096, 1024)]
for (m, n, k) in shapes:
    for opa, opb in (("T", "N"), ("N", "N")):
        rec = dict(bench="small_gemm", m=m, n=n, k=k, opa=opa, opb=opb, mflop=2e-6 * m * n * k)
        for name, impl in (("naive", 1), ("cpasync", 2), ("tma", 3)):
            if impl == 1 and m * n * k > 64e6:
                continue
            rec[name + "_us"] = round(ctx.gemm_bench(m, n, k, opa, opb, impl=impl, reps=50) * 1e3, 2)
        print(json.dumps(rec), flush=True)
