"""Diagnostic: where does the time of a device-resident H_eff application go (kernel time vs gaps)?"""
import sys, time, subprocess
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns
from bench import build_problem, ClockSampler
chi = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
nsites = int(sys.argv[2]) if len(sys.argv) > 2 else 26
ctx = ns.Context(0)
def run(tag):
    net, region = build_problem(chi, nsites, ctx)
    net.extract(region)
    ctx.synchronize()
    for _ in range(3):
        net.matvec_device(1)
    ctx.synchronize()
    single = []
    for _ in range(5):
        ctx.tic(); net.matvec_device(1); single.append(ctx.toc())
    with ClockSampler(0) as clk:
        ctx.tic()
        for _ in range(10):
            net.matvec_device(1)
        b2b = ctx.toc() / 10
    t0 = time.perf_counter()
    for _ in range(10):
        net.matvec_device(1)
    host_enqueue = (time.perf_counter() - t0) / 10
    ctx.synchronize()
    print(tag, "single", [round(x, 2) for x in single], "back-to-back", round(b2b, 2), "host enqueue ms/matvec", round(host_enqueue * 1e3, 3),
          clk.summary(), ctx.mem_info(), flush=True)
    del net
run(f"nsites={nsites}")
print(subprocess.run(["nvidia-smi", "--query-gpu=power.draw,power.limit,clocks.sm,temperature.gpu", "--format=csv"], capture_output=True, text=True).stdout)
