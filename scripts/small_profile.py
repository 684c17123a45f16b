#!/usr/bin/env python
"""Latency-bound instances through the drop-in binary with LORADS_PROFILE=1 (per-kernel-class device time at exit), three
runs each, plus an FP64 GEMM yardstick (cuBLAS DGEMM through torch) for the dense-cone kernels.  GPU box; JSON lines."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ltr-lowrank-sdp_b200"))
GSET = ["--phase1Tol", "1e-2", "--heuristicFactor", "10"]


def parse(out):
    res = {"classes": {}}
    for line in out.splitlines():
        if line.startswith("profile "):
            t = line.split()
            res["classes"][t[1]] = {"launches": int(t[3]), "device_ms": float(t[6])}
        elif "kernel launches" in line:
            res["launches"] = int(line.split("lorads_b200:")[1].split()[0])
        elif line.startswith("all_time:"):
            res["all_time_s"] = float(line.split(":")[1])
        elif line.startswith("all_time - all_dual_infea:"):
            res["alm_admm_s"] = float(line.split(":")[1])
        elif "OuterIter:" in line and "InnerIter:" in line:
            res["alm_inner"] = int(line.split("InnerIter:")[1].split()[0])
        elif line.startswith("ADMM Iter:"):
            res["admm"] = int(line.split("Iter:")[1].split()[0]) + 1
            if "cgIter:" in line:
                res["cg_iters"] = int(line.split("cgIter:")[1].split()[0])
        elif "1.Primal Objective:" in line:
            res["obj"] = float(line.split(":")[-1])
    return res


def main():
    import lorads_b200 as lb
    ei, ej, w = lb.torus_graph(100, 200, 81)
    torus = "/tmp/torus_100x200.dat-s"
    lb.write_sdpa(torus, lb.maxcut_problem(20000, ei, ej, w))
    cases = [("G1", os.path.join(ROOT, "tests/golden/instances/G1.dat-s"), GSET + ["--reoptLevel", "0"]),
             ("G11", os.path.join(ROOT, "tests/golden/instances/G11.dat-s"), GSET),
             ("torus_100x200", torus, GSET + ["--reoptLevel", "0"]),
             ("delaunay_n14", os.path.join(ROOT, "bench_data/delaunay_n14.dat-s"), ["--phase1Tol", "1e+1", "--heuristicFactor", "100", "--timesLogRank", "0.25"]),
             ("MC_500", os.path.join(ROOT, "bench_data/MC_500.dat-s"), []),
             ("cphil12", os.path.join(ROOT, "bench_data/cphil12.dat-s"), []),
             ("control_like_12_6", os.path.join(ROOT, "tests/golden/instances/control_like_12_6.dat-s"), [])]
    if "--theta" in sys.argv:
        cases.append(("theta102", os.path.join(ROOT, "bench_data/theta102.dat-s"), ["--timeSecLimit", "150"]))
    for name, path, flags in cases:
        if not os.path.exists(path):
            continue
        runs = []
        for k in range(1 if name == "theta102" else 4):
            env = dict(os.environ)
            if k == 0:
                env["LORADS_PROFILE"] = "1"   # run 0: per-class device time (events around every launch inflate its wall time)
            t0 = time.perf_counter()
            out = subprocess.run([lb.BINARY_PATH, path] + flags, capture_output=True, text=True, env=env, timeout=900)
            r = parse(out.stdout + out.stderr)
            r["process_wall_s"] = time.perf_counter() - t0
            r["profiled"] = k == 0
            runs.append(r)
        print(json.dumps({"instance": name, "flags": " ".join(flags), "runs": runs}), flush=True)
    try:
        import torch
        res = {}
        for n in (4096, 8192):
            a = torch.randn(n, n, dtype=torch.float64, device="cuda")
            b = torch.randn(n, n, dtype=torch.float64, device="cuda")
            for _ in range(2):
                a @ b
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                a @ b
            e1.record()
            torch.cuda.synchronize()
            res[str(n)] = 5 * 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12
        # the shapes of the dense cones: tall-skinny products (n = 500, r = 32 .. 256)
        for (n, r) in ((512, 32), (512, 256), (2048, 64)):
            u = torch.randn(n, r, dtype=torch.float64, device="cuda")
            for _ in range(3):
                u @ u.T
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                u @ u.T
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 50
            res[f"syrk_like_{n}x{r}"] = {"ms": ms, "tflops": 2.0 * n * n * r / (ms * 1e-3) / 1e12}
        print(json.dumps({"dgemm_yardstick_tflops": res, "note": "cuBLAS DGEMM via torch.matmul, CUDA events, 1 B200"}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"dgemm_yardstick_error": repr(e)}))


if __name__ == "__main__":
    main()
