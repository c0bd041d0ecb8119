"""Sustained cuBLAS GEMM throughput under the board power cap, bf16 vs fp16 operands and random vs
zero data: calibrates what "tensor peak" means for fp16 x fp16 -> fp32 MMAs fed real activations.
nvidia-smi samples clocks / power in the background.  python tools/gemm_power_calib.py [seconds=6]"""
import json, subprocess, sys, time
import torch

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
n = 8192
out = {}
for dt, name in ((torch.bfloat16, "bf16"), (torch.float16, "fp16")):
    for data in ("randn", "zeros"):
        a = (torch.randn if data == "randn" else torch.zeros)((n, n), device="cuda", dtype=dt)
        b = (torch.randn if data == "randn" else torch.zeros)((n, n), device="cuda", dtype=dt)
        c = torch.empty((n, n), device="cuda", dtype=dt)
        for _ in range(10):
            torch.matmul(a, b, out=c)
        torch.cuda.synchronize()
        smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active",
                                "--format=csv,noheader,nounits", "-lms", "200", "-i", "0"], stdout=subprocess.PIPE, text=True)
        chunks = []
        t0 = time.time()
        while time.time() - t0 < secs:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(400):
                torch.matmul(a, b, out=c)
            e1.record()
            chunks.append((e0, e1))
            while not e0.query():          # keep ~one chunk queued ahead, never drain the GPU
                time.sleep(0.001)
        torch.cuda.synchronize()
        smi.terminate()
        rows = [l.split(",") for l in smi.communicate()[0].strip().splitlines() if l.count(",") >= 2]
        rows = rows[len(rows) // 2:]
        tf = [2 * n ** 3 * 400 / e0.elapsed_time(e1) / 1e9 for e0, e1 in chunks]
        out[f"{name}_{data}"] = {"tflops_last_half": sum(tf[len(tf) // 2:]) / max(1, len(tf) - len(tf) // 2), "tflops_first_chunk": tf[0],
                                 "sm_mhz": sum(float(r[0]) for r in rows) / max(1, len(rows)),
                                 "watts": sum(float(r[1]) for r in rows) / max(1, len(rows)),
                                 "reasons": sorted({r[2].strip() for r in rows})}
        time.sleep(1.0)
print(json.dumps(out))
