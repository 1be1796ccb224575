#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 Householder-QR hot path.

Metric (BASELINE.json): batched 32x32 Float64 QR, matrices/s, 1,048,576 matrices per GPU, sharded by
matrix index with no data-path collective (weak scaling: every rank owns its own 2^20-matrix slab).
One "step" = one pass of the hot path over the rank's slab = ONE launch of batched_qr32_reg_kernel.

  value      whole-job matrices/s with the slab resident in HBM (CUDA events, max over ranks)
  e2e        the same through the host-pointer C ABI call (gla_dgeqr_batched) on PINNED host buffers:
             H2D of the slab, kernel, D2H of factors + tau, all inside the timed region
  roofline   HBM-bound: algorithmic bytes = 16,640 B per matrix (8192 in + 8192 factors + 256 tau)
  cpu_baseline  the oracle (C++ restatement of the reference's qrBlocked!, blocksize 12, OpenMP over
             matrices) timed on this box's host cores on a bounded sample
  also       FP64 qrBlocked! n=16384 TFLOP/s (the other half of BASELINE.json's metric), Cholesky
             n=4096 and TSQR 8,388,608x64, each measured once after the timed region (rank 0 / all ranks
             for TSQR), reported under "other_configs"

`--impl reference` times the reference's CPU path (the oracle port; no Julia runtime exists in the
image) on the host cores for the same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M = N_ = 32
BATCH_PER_GPU = 1 << 20
BYTES_PER_MATRIX = 8192 + 8192 + 256
FLOPS_PER_MATRIX = 4.0 / 3.0 * 32 ** 3
# dram__bytes_read.sum + dram__bytes_write.sum of one launch (2^20 matrices), ncu --set full: profiles/r01_ncu_batched_ll4_s5.txt
NCU_DRAM_BYTES_PER_LAUNCH = 8.590219e9 + 8.802413e9
FP64_TENSOR_PEAK_TFLOPS = 37.08               # measured on this pool: tools/fp64_peak.cu (profiles/fp64_peak_r01.txt)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_baseline(seconds=12.0):
    """Oracle port of the reference (qrBlocked!, blocksize 12) on the host cores, OpenMP over matrices."""
    import numpy as np
    from oracle import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    rng = np.random.default_rng(123)
    chunk = 16384
    src = rng.standard_normal((chunk, 32, 32))
    tau = np.zeros((chunk, 32))
    buf = src.copy()
    oracle.qr_batched_raw(buf, 32, 32, 1024, tau)      # warm-up (thread pool, page faults)
    t_fact, reps = 0.0, 0
    while t_fact < seconds and reps < 64:
        buf[...] = src
        t1 = time.perf_counter()
        oracle.qr_batched_raw(buf, 32, 32, chunk, tau)
        t_fact += time.perf_counter() - t1
        reps += 1
    rate = reps * chunk / t_fact
    return {"value": rate, "unit": "matrices/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {chunk} random 32x32 Float64 matrices, oracle qr_blocked (blocksize 12), "
                      f"OMP threads = {cores}; Julia unavailable in image (JULIA_NUM_THREADS n/a)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(seconds=max(5.0, 2.0 * args.steps))
    line = {"impl": "reference", "metric": "batched 32x32 Float64 QR matrices/s", "value": cb["value"],
            "unit": "matrices/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": BATCH_PER_GPU / cb["value"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched_qr_32x32_f64", "matrices_per_gpu": BATCH_PER_GPU, "m": 32, "n": 32,
                       "note": "CPU port timed on a bounded sample; ms_per_step extrapolated to 2^20 matrices"},
            "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="matrices per GPU (default = BASELINE config)")
    ap.add_argument("--skip-other", action="store_true", help="skip the n=16384 QR / Cholesky / TSQR side measurements")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    g = ge.load()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    g.set_device(local)
    dev = torch.device("cuda", local)
    batch = args.batch
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident slabs: ring of 2 buffers (each 8.6 GB >> 126 MB L2, so no step sees a warm L2)
    gen = torch.Generator(device=dev).manual_seed(123 + rank)
    NBUF = 2
    slabs = [torch.randn((batch, 32, 32), generator=gen, device=dev, dtype=torch.float64) for _ in range(NBUF)]
    taus = torch.zeros((batch, 32), device=dev, dtype=torch.float64)

    def step(i):
        g.qr_batched_dev(slabs[i % NBUF].data_ptr(), 32, 32, batch, taus.data_ptr(), stream)

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = ms_total / args.steps
    value = world * batch / (ms_per_step * 1e-3)
    kernel_ms = ms_per_step  # one launch per step: the step IS the kernel

    # ---- e2e through the host-pointer C ABI on pinned buffers (H2D + kernel + D2H inside the timed region)
    e2e_steps = max(1, min(3, args.steps))
    hA = torch.empty((batch, 32, 32), dtype=torch.float64, pin_memory=True)
    ht = torch.empty((batch, 32), dtype=torch.float64, pin_memory=True)
    hA.copy_(slabs[0])   # dense data (timing is data independent)
    g.qr_batched_ptr(hA.data_ptr(), 32, 32, min(batch, 65536), ht.data_ptr())   # warm-up of the staged path
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g.qr_batched_ptr(hA.data_ptr(), 32, 32, batch, ht.data_ptr())
    torch.cuda.synchronize()
    el = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    e2e_value = world * batch * e2e_steps / el.item()
    h2d = batch * 8192
    d2h = batch * (8192 + 256)

    # ---- smoke-level parity of what was just timed (oracle on a few matrices of the last slab is not
    # possible after in-place factorisation; instead check the Gram identity R^T R = A^T A on a fresh sample)
    chk = torch.randn((256, 32, 32), generator=gen, device=dev, dtype=torch.float64)
    chk0 = chk.clone()
    g.qr_batched_dev(chk.data_ptr(), 32, 32, 256, taus.data_ptr(), stream)
    torch.cuda.synchronize()
    Rm = torch.triu(chk.transpose(1, 2))
    A0 = chk0.transpose(1, 2)
    gram_err = ((Rm.transpose(1, 2) @ Rm - A0.transpose(1, 2) @ A0).abs().amax() /
                (A0.transpose(1, 2) @ A0).abs().amax()).item()

    other = {}
    del slabs, hA, ht
    torch.cuda.empty_cache()
    if not args.skip_other:
        other = other_configs(g, torch, dist, dev, rank, world, stream)

    hbm_peak, peak_src = _peaks()
    achieved = BYTES_PER_MATRIX * batch / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": "batched 32x32 Float64 QR matrices/s", "value": value, "unit": "matrices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched_qr_32x32_f64 (BASELINE configs[2])", "matrices_per_gpu": batch, "m": 32,
                   "n": 32, "shard": "matrix index, no collective", "l2": "inputs larger than L2 (8.6 GB slab per step, ring of 2)",
                   "seed": 123, "gram_check": gram_err},
        "e2e": {"value": e2e_value, "unit": "matrices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "gla_dgeqr_batched (host pointers, pinned)"},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak, "traffic": NCU_DRAM_BYTES_PER_LAUNCH * batch / BATCH_PER_GPU,
                     "kernel": "batched_qr32_ll4_kernel<double>", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_MATRIX * batch,
                     "fp64_tflops": FLOPS_PER_MATRIX * batch / (kernel_ms * 1e-3) / 1e12},
        "clocks": clocks,
        "other_configs": other,
    }
    if rank == 0:
        if not args.skip_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _time(torch, fn, reps=3):
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), ts


def other_configs(g, torch, dist, dev, rank, world, stream):
    """The remaining BASELINE configs, measured once each after the timed region (device-resident)."""
    out = {}
    if rank == 0:
        # metric part 1: FP64 qrBlocked! n = 16384
        n = 16384
        src = torch.randn((n, n), device=dev, dtype=torch.float64)
        dA = torch.empty_like(src)
        dtau = torch.zeros(n, device=dev, dtype=torch.float64)

        def qr():
            g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, stream)
        best = 1e30
        for _ in range(3):
            dA.copy_(src)
            ms, _ = _time(torch, qr, reps=1)
            best = min(best, ms)
        tf = 4.0 / 3.0 * n ** 3 / (best * 1e-3) / 1e12
        # validate what was timed: R^T R x = A^T A x for a random x (storage is column-major: dA[j, i] = F[i, j])
        xv = torch.randn(n, device=dev, dtype=torch.float64)
        Rm = torch.triu(dA.t())
        y1 = Rm.t() @ (Rm @ xv)
        y2 = src @ (src.t() @ xv)
        probe = ((y1 - y2).abs().max() / y2.abs().max()).item()
        del Rm, y1, y2
        out["qr_f64_n16384"] = {"ms": best, "tflops": tf, "unit": "TFLOP/s", "gram_probe": probe,
                                "roofline": {"bound": "tensor", "achieved": tf, "peak": FP64_TENSOR_PEAK_TFLOPS,
                                             "unit": "TFLOP/s", "frac": tf / FP64_TENSOR_PEAK_TFLOPS,
                                             "peak_source": "measured DMMA.8x8x4 peak, tools/fp64_peak.cu"}}
        del src, dA
        # config 5: ComplexF64 n = 16384 (complex reflectors, contraction on the FP64 tensor pipe as 2 real DMMA products)
        import numpy as np
        src = torch.randn((n, n), device=dev, dtype=torch.complex128)
        dA = torch.empty_like(src)
        ztau = torch.zeros(n, device=dev, dtype=torch.complex128)
        best = 1e30
        for _ in range(2):
            dA.copy_(src)
            ms, _ = _time(torch, lambda: g.qr_blocked_dev(dA.data_ptr(), n, n, n, ztau.data_ptr(), 0, stream, np.complex128), reps=1)
            best = min(best, ms)
        tf = 4.0 * 4.0 / 3.0 * n ** 3 / (best * 1e-3) / 1e12
        xv = torch.randn(n, device=dev, dtype=torch.complex128)
        Rm = torch.triu(dA.t())
        y1 = Rm.conj().t() @ (Rm @ xv)
        y2 = src.conj() @ (src.t() @ xv)
        probe = ((y1 - y2).abs().max() / y2.abs().max()).item()
        del Rm, y1, y2
        out["qr_c128_n16384"] = {"ms": best, "tflops_real": tf, "gram_probe": probe, "unit": "real TFLOP/s (4 per complex FMA pair)",
                                 "roofline": {"bound": "tensor", "achieved": tf, "peak": FP64_TENSOR_PEAK_TFLOPS,
                                              "unit": "TFLOP/s", "frac": tf / FP64_TENSOR_PEAK_TFLOPS,
                                              "peak_source": "measured DMMA.8x8x4 peak, tools/fp64_peak.cu"}}
        del src, dA, ztau
        torch.cuda.empty_cache()
        # config 1: n = 1024
        n = 1024
        src = torch.randn((n, n), device=dev, dtype=torch.float64)
        dA = torch.empty_like(src)
        best = 1e30
        for _ in range(4):
            dA.copy_(src)
            ms, _ = _time(torch, lambda: g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, stream), reps=1)
            best = min(best, ms)
        out["qr_f64_n1024"] = {"ms": best, "tflops": 4.0 / 3.0 * n ** 3 / (best * 1e-3) / 1e12}
        # config 2: Cholesky n = 4096
        n = 4096
        X = torch.randn((n, n), device=dev, dtype=torch.float64)
        S = X.t() @ X + n * torch.eye(n, device=dev, dtype=torch.float64)
        dS = torch.empty_like(S)
        info = torch.zeros(1, device=dev, dtype=torch.int32)
        best = 1e30
        for _ in range(4):
            dS.copy_(S)
            ms, _ = _time(torch, lambda: g.chol_recursive_dev(dS.data_ptr(), n, n, info.data_ptr(), 1, stream), reps=1)
            best = min(best, ms)
        out["chol_f64_n4096"] = {"ms": best, "tflops": n ** 3 / 3.0 / (best * 1e-3) / 1e12, "info": int(info.item())}
        del X, S, dS
        torch.cuda.empty_cache()
    # config 4: TSQR 8,388,608 x 64, row-sharded; the 64x64 R factors are exchanged by ONE ncclAllGather issued
    # inside the library (gla_dtsqr_allreduce_dev on a library-owned communicator) and reduced on every rank
    m_total, n = 1 << 23, 64
    _, rows = g.shard_range(m_total, rank, world)
    A = torch.randn((n, rows), device=dev, dtype=torch.float64)       # column-major rows x n
    Rloc = torch.zeros((n, n), device=dev, dtype=torch.float64)
    Rall = torch.zeros((world, n, n), device=dev, dtype=torch.float64)
    R = torch.zeros((n, n), device=dev, dtype=torch.float64)
    comm = None
    if world > 1:
        def bcast(raw):
            t = torch.zeros(128, dtype=torch.uint8, device=dev)
            if raw is not None:
                t.copy_(torch.tensor(list(raw), dtype=torch.uint8))
            dist.broadcast(t, 0)
            return bytes(t.cpu().tolist())
        comm = g.TsqrComm(rank, world, bcast)

    def tsqr():
        g.tsqr_local_dev(A.data_ptr(), rows, n, rows, (Rloc if world > 1 else R).data_ptr(), n, stream)
        if world > 1:
            comm.allreduce_R(Rloc.data_ptr(), n, Rall.data_ptr(), R.data_ptr(), n, stream)
    for _ in range(2):
        tsqr()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms, _ = _time(torch, tsqr, reps=5)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    # Gram identity of the local shard (sign independent): R_loc^T R_loc = A^T A
    Rl = torch.triu((Rloc if world > 1 else R).t())
    G = A @ A.t()
    gram = ((Rl.t() @ Rl - G).abs().amax() / G.abs().amax()).item()
    tf = 2.0 * m_total * n * n / (ms * 1e-3) / 1e12
    out["tsqr_f64_8388608x64"] = {"ms": ms, "rows_per_s": m_total / (ms * 1e-3), "gb_per_s": m_total * n * 8 / (ms * 1e-3) / 1e9,
                                  "tflops": tf, "scaling": "strong", "gram_check_local": gram,
                                  "roofline": {"bound": "fp64 (AI 16 flop/B)", "achieved": tf, "peak": FP64_TENSOR_PEAK_TFLOPS * world,
                                               "unit": "TFLOP/s", "frac": tf / (FP64_TENSOR_PEAK_TFLOPS * world),
                                               "hbm_frac": m_total * n * 8 / (ms * 1e-3) / 1e9 / (_peaks()[0] * world)},
                                  "collective": "ncclAllGather of 64x64 R factors inside gla_dtsqr_allreduce_dev" if world > 1 else "none"}
    if comm is not None:
        torch.cuda.synchronize()
        comm.destroy()
    return out


if __name__ == "__main__":
    main()
