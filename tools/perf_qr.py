"""Timing + accuracy of the blocked compact-WY Householder QR (qr_thin) on device-resident data.
    python tools/perf_qr.py [rows cols] ..."""
import sys, time, json
sys.path.insert(0, ".")
import numpy as np
import networksolvers_b200 as ns

ctx = ns.default_context()
shapes = [(8192, 4096), (4096, 2048), (2048, 1024), (1024, 2048)]
if len(sys.argv) > 2:
    shapes = [(int(sys.argv[i]), int(sys.argv[i + 1])) for i in range(1, len(sys.argv) - 1, 2)]
rng = np.random.default_rng(0)
for rows, cols in shapes:
    M = rng.standard_normal((rows, cols))
    for blk in (64, 0):
        if blk == 0 and rows * cols > 2048 * 1024:
            continue
        ctx.set_option("qr_block_min", blk)
        t0 = time.perf_counter(); Q, R = ctx.qr(M); dt = time.perf_counter() - t0
        k = min(rows, cols)
        orth = float(np.abs(Q.T @ Q - np.eye(k)).max())
        res = float(np.abs(Q @ R - M).max())
        low = float(np.abs(np.tril(R[:, :k], -1)).max())
        ms = ctx.qr_bench(rows, cols, reps=3)
        print(json.dumps(dict(rows=rows, cols=cols, blocked=bool(blk), device_ms=ms, host_call_s=dt, orth=orth, resid=res, lower=low)), flush=True)
    ctx.set_option("qr_block_min", 64)
