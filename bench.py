#!/usr/bin/env python
"""bench.py -- headline measurement of the B200 Householder-QR hot path.

Metric (BASELINE.json): batched 32x32 Float64 QR, matrices/s, 1,048,576 matrices per GPU, sharded by
matrix index with no data-path collective (weak scaling: every rank owns its own 2^20-matrix slabs).
One "step" = one pass of the hot path over one FRESH slab = ONE launch of batched_qr32_ll4_kernel.

  value      whole-job matrices/s with the slab resident in HBM (CUDA events, max over ranks)
  e2e        the same through the host-pointer C ABI call (gla_dgeqr_batched) on PINNED host buffers:
             H2D of the slab, kernel, D2H of factors + tau, all inside the timed region
  roofline   HBM-bound: algorithmic bytes = 16,640 B per matrix (8192 in + 8192 factors + 256 tau);
             traffic = dram bytes of the committed ncu capture (profiles/ncu_batched_current.txt)
  cpu_baseline  the oracle (C++ restatement of the reference's qrBlocked!, blocksize 12, OpenMP over
             matrices) timed on this box's host cores on a bounded sample; OpenMP team size forced to
             the core count and the team size actually in effect reported
  parity     what was timed is checked in the same run: sampled matrices of the LAST TIMED slab against the
             oracle elementwise, and the whole slab through the Gram identity
  also       FP64 qrBlocked! n=16384 TFLOP/s (the other half of BASELINE.json's metric), ComplexF64
             n=16384, n=1024, Cholesky n=4096 and TSQR 8,388,608x64 under "other_configs", each with the
             reference's CPU path (oracle port) timed beside it (rank 0, N=1) and an end-to-end
             host-pointer figure for the single-matrix paths

`--impl reference` times the reference's CPU path (the oracle port; no Julia runtime exists in the
image) on the host cores on the FULL per-GPU config (2^20 matrices per step, no extrapolation).
"""
from __future__ import annotations

import argparse
import json
import os
import re
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
NCU_SUMMARY = os.path.join(ROOT, "profiles", "ncu_batched_current.txt")   # ncu --set full capture of the shipped kernel
FP64_TENSOR_PEAK_TFLOPS = 37.08               # measured on this pool: tools/fp64_peak.cu (profiles/r01_fp64_peak.txt)
PARITY_NOTE = ("oracle = C++ restatement of the reference (oracle/gla_oracle.cpp); parity UNPINNED: the reference is "
               "100 % Julia, no Julia runtime in the image, no golden vectors in the reference's tests; pinned only by "
               "hand-derived KATs (tests/golden) and LAPACK cross-checks")


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch (2^20 matrices) from the committed ncu summary."""
    try:
        txt = open(NCU_SUMMARY).read()
        tot, scale = 0.0, {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(re.escape(key) + r" \[(\w+)\] = ([0-9.eE+-]+)", txt)
            tot += float(m.group(2)) * scale[m.group(1)]
        return tot, os.path.relpath(NCU_SUMMARY, ROOT)
    except Exception:
        return None, None


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


# ------------------------------------------------------------------------------------------------- CPU side
def _oracle():
    """The oracle with its OpenMP team forced to every host core (torchrun exports OMP_NUM_THREADS=1, which made the
    round-1 CPU arm run on ONE thread under a 32-thread label).  Returns (module, cores requested, team size in effect)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    from oracle import oracle
    oracle.build()
    oracle.set_threads(cores)
    return oracle, cores, oracle.max_threads()


def _tiled_slab(np, batch, base=1 << 16, seed=123):
    """`batch` random 32x32 matrices: a base of 2^16 distinct standard-normal matrices tiled (the CPU path's time is
    data independent; generating 8.6 GB of normals single-threaded would take longer than the measurement)."""
    rng = np.random.default_rng(seed)
    src = rng.standard_normal((min(base, batch), 32, 32))
    return src


def _fill(np, buf, src):
    b, n = buf.shape[0], src.shape[0]
    for o in range(0, b, n):
        k = min(n, b - o)
        buf[o:o + k] = src[:k]


def cpu_baseline(seconds=12.0):
    """Headline config on the host cores, bounded: as many passes over a 2^16-matrix sample as fit in `seconds`."""
    import numpy as np
    oracle, cores, team = _oracle()
    chunk = 1 << 16
    src = _tiled_slab(np, chunk)
    tau = np.zeros((chunk, 32))
    buf = src.copy()
    oracle.qr_batched_raw(buf, 32, 32, chunk, tau)      # warm-up (thread pool, page faults)
    t_fact, reps = 0.0, 0
    while t_fact < seconds and reps < 64:
        buf[...] = src
        t1 = time.perf_counter()
        oracle.qr_batched_raw(buf, 32, 32, chunk, tau)
        t_fact += time.perf_counter() - t1
        reps += 1
    rate = reps * chunk / t_fact
    return {"value": rate, "unit": "matrices/s", "cores": team, "kind": "port",
            "sample": f"{reps} x {chunk} random 32x32 Float64 matrices, oracle qr_blocked (blocksize 12), OpenMP team = {team} "
                      f"(requested {cores}); Julia unavailable in image (JULIA_NUM_THREADS n/a)"}


def cpu_other_configs():
    """The reference's CPU path (oracle port) beside every other reported shape: C1 and C2 in full, C4 on a
    2^20-row sample, C5 / the n=16384 metric at n=2048 with the n^3 extrapolation labelled."""
    import numpy as np
    oracle, cores, team = _oracle()
    rng = np.random.default_rng(123)
    out = {}

    def best(fn, reps=2):
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return min(ts)

    A = np.asfortranarray(rng.standard_normal((1024, 1024)))
    oracle.qr_blocked(A[:256, :256].copy(order="F"), 12)   # warm the thread pool
    t = best(lambda: oracle.qr_blocked(A, 12), 3)
    out["qr_f64_n1024"] = {"ms": t * 1e3, "tflops": 4.0 / 3.0 * 1024 ** 3 / t / 1e12, "cores": team, "kind": "port",
                           "sample": "full config: oracle qr_blocked(1024x1024, blocksize 12)"}
    X = rng.standard_normal((4096, 4096))
    S = np.asfortranarray(X.T @ X + 4096 * np.eye(4096))
    t = best(lambda: oracle.chol_recursive(S, 1, mt=False), 1)
    t_mt = best(lambda: oracle.chol_recursive(S, 1, mt=True), 1)
    out["chol_f64_n4096"] = {"ms": t * 1e3, "tflops": 4096 ** 3 / 3.0 / t / 1e12, "cores": team, "kind": "port",
                             "sample": "full config: oracle chol_recursive(4096), generic rankUpdate! single-threaded as in the "
                                       "reference (src/juliaBLAS.jl:89-112), rdiv! multithreaded (BLAS trsm in the reference)",
                             "ms_with_threaded_rank_update": t_mt * 1e3}
    del X, S
    rows = 1 << 20
    T = np.asfortranarray(rng.standard_normal((rows, 64)))
    t = best(lambda: oracle.qr_blocked(T, 12), 1)
    out["tsqr_f64_8388608x64"] = {"ms": t * 8 * 1e3, "tflops": 2.0 * rows * 64 * 64 / t / 1e12, "cores": team, "kind": "port",
                                  "sample": "oracle qr_blocked on a 2^20 x 64 sample; ms EXTRAPOLATED x8 (linear in m)"}
    del T
    n = 2048
    A = np.asfortranarray(rng.standard_normal((n, n)))
    t = best(lambda: oracle.qr_blocked(A, 12), 1)
    out["qr_f64_n16384"] = {"ms": t * 512 * 1e3, "tflops": 4.0 / 3.0 * n ** 3 / t / 1e12, "cores": team, "kind": "port",
                            "sample": "oracle qr_blocked at n=2048; ms EXTRAPOLATED x512 ((4/3)n^3)"}
    # f3 rows: the two-sided reductions at n = 2048 (full config)
    t_b = best(lambda: oracle.bidiagonalize(A.copy(order="F")), 1)
    t_h = best(lambda: oracle.hessenberg(A.copy(order="F")), 1)
    S2 = np.asfortranarray(A + A.T)
    t_s = best(lambda: oracle.symtri(S2.copy(order="F"), "L"), 1)
    out["twosided_f64_n2048"] = {"bidiagonalize_ms": t_b * 1e3, "hessenberg_ms": t_h * 1e3, "symtri_ms": t_s * 1e3,
                                 "cores": team, "kind": "port",
                                 "sample": "full config: oracle bidiagonalize / hessenberg / symtri(:L) at n=2048 (OpenMP over the "
                                           "independent columns / rows of each reflector application)"}
    del S2
    Z = np.asfortranarray(A + 1j * rng.standard_normal((n, n)))
    t = best(lambda: oracle.qr_blocked(Z, 12), 1)
    out["qr_c128_n16384"] = {"ms": t * 512 * 1e3, "tflops_real": 4.0 * 4.0 / 3.0 * n ** 3 / t / 1e12, "cores": team,
                             "kind": "port", "sample": "oracle qr_blocked (ComplexF64) at n=2048; ms EXTRAPOLATED x512"}
    return out


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (oracle port) on the FULL per-GPU config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    oracle, cores, team = _oracle()
    batch = args.batch
    src = _tiled_slab(np, batch)
    buf = np.empty((batch, 32, 32))
    tau = np.zeros((batch, 32))
    _fill(np, buf, src)
    oracle.qr_batched_raw(buf, 32, 32, min(batch, 1 << 16), tau)     # thread pool, page faults
    times = []
    for i in range(max(1, min(args.warmup, 1)) + args.steps):
        _fill(np, buf, src)                                            # fresh input, outside the timed region
        t0 = time.perf_counter()
        oracle.qr_batched_raw(buf, 32, 32, batch, tau)
        dt = time.perf_counter() - t0
        if i >= max(1, min(args.warmup, 1)):
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    rate = batch / (ms * 1e-3)
    cb = {"value": rate, "unit": "matrices/s", "cores": team, "kind": "port",
          "sample": f"full config: {args.steps} timed passes over {batch} 32x32 Float64 matrices (2^16 distinct matrices tiled), "
                    f"oracle qr_blocked blocksize 12, OpenMP team = {team} (requested {cores}); one per-GPU slab per step "
                    f"whatever --gpus says; Julia unavailable in image (JULIA_NUM_THREADS n/a)"}
    line = {"impl": "reference", "metric": "batched 32x32 Float64 QR matrices/s", "value": rate,
            "unit": "matrices/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched_qr_32x32_f64 (BASELINE configs[2])", "matrices_per_gpu": batch, "m": 32, "n": 32,
                       "shard": "matrix index, no collective", "seed": 123},
            "cpu_baseline": cb, "parity": PARITY_NOTE,
            "e2e": {"value": rate, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- GPU side
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

    # ---- resident slabs: every timed step factorises a FRESH slab (8.6 GB >> 126 MB L2, so no step sees a warm L2 and
    # no step re-factorises factors); up to 12 slabs (103 GB of the 180 GB), beyond that the ring wraps
    slab_bytes = batch * 8192
    nslab = max(1, min(args.steps, 12, int(140e9 // max(slab_bytes, 1))))

    def make_slab(i):
        gen = torch.Generator(device=dev).manual_seed(123 + 1000 * rank + i)
        return torch.randn((batch, 32, 32), generator=gen, device=dev, dtype=torch.float64)

    warm = make_slab(10_000)
    slabs = [make_slab(i) for i in range(nslab)]
    taus = torch.zeros((batch, 32), device=dev, dtype=torch.float64)

    def step(t):
        g.qr_batched_dev(t.data_ptr(), 32, 32, batch, taus.data_ptr(), stream)

    for i in range(args.warmup):
        step(warm)
    del warm
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
        step(slabs[i % nslab])
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

    # ---- parity of WHAT WAS TIMED: the last timed slab (factorised exactly once when steps <= nslab)
    last = (args.steps - 1) % nslab
    once = args.steps <= nslab
    F = slabs[last]
    tau_last = taus.clone()
    if not once:    # the ring wrapped: this slab was factorised more than once; redo it once from its seed, untimed
        F = make_slab(last)
        step(F)
        tau_last = taus.clone()
    orig = make_slab(last)
    torch.cuda.synchronize()
    Rm = torch.triu(F.transpose(1, 2))
    A0 = orig.transpose(1, 2)
    gram_ref = torch.matmul(A0.transpose(1, 2), A0)
    gram_err = ((torch.matmul(Rm.transpose(1, 2), Rm) - gram_ref).abs().amax() / gram_ref.abs().amax()).item()
    del Rm, gram_ref
    parity = {"oracle": PARITY_NOTE, "slab": "last timed slab" if once else "last slab, re-run once untimed (ring wrapped)",
              "gram_all_matrices": gram_err}
    if rank == 0 and not args.skip_cpu:
        oracle, _, _ = _oracle()
        idx = torch.linspace(0, batch - 1, min(batch, 512), device=dev).long()
        a_h = orig[idx].cpu().numpy()                    # (s, col, row)
        f_h = F[idx].cpu().numpy()
        t_h = tau_last[idx].cpu().numpy()
        rf, rt = oracle.qr_batched(np.transpose(a_h, (0, 2, 1)))
        rf = np.transpose(rf, (0, 2, 1))
        sc = np.max(np.abs(rf), axis=(1, 2), keepdims=True)
        parity["oracle_sample"] = int(idx.numel())
        parity["oracle_max_rel_factors"] = float(np.max(np.abs(f_h - rf) / sc))
        parity["oracle_max_abs_tau"] = float(np.max(np.abs(t_h - rt)))
        parity["ok"] = bool(parity["oracle_max_rel_factors"] < 1e-10 and parity["oracle_max_abs_tau"] < 1e-10 and gram_err < 1e-10)
    del orig, F, slabs
    torch.cuda.empty_cache()

    # ---- e2e through the host-pointer C ABI on pinned buffers (H2D + kernel + D2H inside the timed region)
    e2e_steps = max(1, min(3, args.steps))
    hA = torch.empty((batch, 32, 32), dtype=torch.float64, pin_memory=True)
    ht = torch.empty((batch, 32), dtype=torch.float64, pin_memory=True)
    hA.copy_(make_slab(0))   # dense data (timing is data independent)
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
    # the ceiling of that path on this box: raw pinned copies in both directions at once (PCIe is full duplex), all ranks
    # together; e2e cannot exceed bytes / these rates whatever the kernel does
    nb = min(1 << 31, batch * 8192)
    dbuf_in = torch.empty(nb, dtype=torch.uint8, device=dev)
    dbuf_out = torch.empty(nb, dtype=torch.uint8, device=dev)
    hin = hA.view(torch.uint8).reshape(-1)[:nb]
    hout = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        with torch.cuda.stream(s1):
            dbuf_in.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dbuf_out, non_blocking=True)
    torch.cuda.synchronize()
    tc = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
    duplex_gbs = 3 * nb / tc.item() / 1e9            # per direction, per rank, with every rank copying
    e2e_ceiling = world * batch / (max(h2d, d2h) / (duplex_gbs * 1e9))
    del hA, ht, dbuf_in, dbuf_out, hout
    torch.cuda.empty_cache()

    other = {}
    if not args.skip_other:
        other = other_configs(g, torch, dist, dev, rank, world, stream, cpu=(rank == 0 and world == 1 and not args.skip_cpu))

    hbm_peak, peak_src = _peaks()
    achieved = BYTES_PER_MATRIX * batch / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic_per_launch()
    line = {
        "metric": "batched 32x32 Float64 QR matrices/s", "value": value, "unit": "matrices/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched_qr_32x32_f64 (BASELINE configs[2])", "matrices_per_gpu": batch, "m": 32,
                   "n": 32, "shard": "matrix index, no collective",
                   "l2": f"inputs larger than L2: every timed step factorises a fresh 8.6 GB slab ({nslab} slabs resident)",
                   "seed": 123},
        "e2e": {"value": e2e_value, "unit": "matrices/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "gla_dgeqr_batched (host pointers, pinned)",
                "pcie_duplex_gbs_per_direction_per_gpu": duplex_gbs, "pcie_ceiling_matrices_per_s": e2e_ceiling,
                "frac_of_pcie_ceiling": e2e_value / e2e_ceiling},
        "gpu_launches": args.steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved / hbm_peak,
                     "traffic": traffic * batch / BATCH_PER_GPU if traffic else None, "traffic_source": traffic_src,
                     "kernel": "batched_qr32_ll4_kernel<double,12,1,300,2,true>", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_MATRIX * batch,
                     "fp64_tflops": FLOPS_PER_MATRIX * batch / (kernel_ms * 1e-3) / 1e12,
                     "note": "co-limited by the FP64 pipe's operand bandwidth: a DFMA with three distinct register-pair operands "
                             "issues every ~3.5 cycles per sub-partition (profiles/r02_fp64_operand_pattern.txt)"},
        "clocks": clocks,
        "parity": parity,
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


def _panel_check(np, torch, oracle, dA, src, k, complex_):
    """Leading n x k panel of the factors / taus against the oracle run on the leading panel of the input (the first k
    columns of qrBlocked!'s result depend on the first k columns of A only, src/qr.jl:113-146)."""
    panel = np.asfortranarray(src[:k].cpu().numpy().T.copy())
    F = np.asfortranarray(dA[0][:k].cpu().numpy().T.copy())
    tau = dA[1][:k].cpu().numpy()
    rf, rt = oracle.qr_unblocked(panel) if complex_ else oracle.qr_blocked(panel, 12)
    sc = np.max(np.abs(rf))
    return {"columns": k, "max_rel_factors": float(np.max(np.abs(F - rf)) / sc), "max_abs_tau": float(np.max(np.abs(tau - rt)))}


def other_configs(g, torch, dist, dev, rank, world, stream, cpu=False):
    """The remaining BASELINE configs, measured once each after the timed region (device-resident), with the host-pointer
    end-to-end time of the single-matrix paths and (N=1) the reference's CPU path beside each."""
    import numpy as np
    out = {}
    oracle = _oracle()[0] if cpu else None
    if rank == 0:
        # live FP64 GEMM rate of the box (cuBLAS DGEMM 8192^3 through torch) as context for the tensor-pipe denominators
        a = torch.randn((8192, 8192), device=dev, dtype=torch.float64)
        b = torch.randn((8192, 8192), device=dev, dtype=torch.float64)
        torch.matmul(a, b)
        ms, _ = _time(torch, lambda: torch.matmul(a, b), reps=3)
        dgemm_live = 2.0 * 8192 ** 3 / (ms * 1e-3) / 1e12
        del a, b
        peak_note = {"peak": FP64_TENSOR_PEAK_TFLOPS, "peak_source": "measured DMMA.8x8x4 peak, tools/fp64_peak.cu "
                     "(MEASURED_PEAKS.json has no FP64 entry)", "cublas_dgemm_8192_live_tflops": dgemm_live}

        def roof(tf):
            return dict({"bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "frac": tf / FP64_TENSOR_PEAK_TFLOPS}, **peak_note)

        def e2e_qr(n, dt, npdt):
            """host-pointer gla_*geqr_blocked on a pinned matrix: H2D + factorisation + D2H, second call (workspace pool warm)"""
            hA = torch.empty((n, n), dtype=dt, pin_memory=True)
            htau = torch.empty(n, dtype=dt, pin_memory=True)
            best = 1e30
            for _ in range(2):
                hA.copy_(torch.randn((n, n), device=dev, dtype=dt))
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                g.qr_blocked_ptr(hA.data_ptr(), n, n, n, htau.data_ptr(), 0, npdt)
                best = min(best, (time.perf_counter() - t0) * 1e3)
            return best

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
        out["qr_f64_n16384"] = {"ms": best, "tflops": tf, "unit": "TFLOP/s", "gram_probe": probe, "roofline": roof(tf)}
        if oracle is not None:
            out["qr_f64_n16384"]["oracle_leading_panel"] = _panel_check(np, torch, oracle, (dA, dtau), src, 128, False)
        del src, dA
        torch.cuda.empty_cache()
        out["qr_f64_n16384"]["e2e_ms"] = e2e_qr(n, torch.float64, np.float64)
        # config 5: ComplexF64 n = 16384 (complex reflectors, contraction on the FP64 tensor pipe as 2 real DMMA products)
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
                                 "roofline": roof(tf)}
        if oracle is not None:
            out["qr_c128_n16384"]["oracle_leading_panel"] = _panel_check(np, torch, oracle, (dA, ztau), src, 128, True)
        del src, dA, ztau
        torch.cuda.empty_cache()
        # Float32 n = 16384
        src = torch.randn((n, n), device=dev, dtype=torch.float32)
        dA = torch.empty_like(src)
        stau = torch.zeros(n, device=dev, dtype=torch.float32)
        best = 1e30
        for _ in range(2):
            dA.copy_(src)
            ms, _ = _time(torch, lambda: g.qr_blocked_dev(dA.data_ptr(), n, n, n, stau.data_ptr(), 0, stream, np.float32), reps=1)
            best = min(best, ms)
        tf32 = 4.0 / 3.0 * n ** 3 / (best * 1e-3) / 1e12
        bf16_peak = None
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                bf16_peak = float(json.load(f).get("bf16_tflops_sustained", 0.0)) or None
        except OSError:
            pass
        out["qr_f32_n16384"] = {"ms": best, "tflops": tf32, "unit": "TFLOP/s (Float32)",
                                "contraction": "tcgen05.mma kind::tf32 (UTCHMMA), TMA-fed, TMEM accumulators, 3xTF32 split with "
                                               "per-slab rounded accumulation; products with a dimension < 128 on the mma.sync kernel"}
        if bf16_peak:
            # 3xTF32 costs three TF32 MMAs per product and TF32 runs at half the dense bf16 rate
            peak32 = bf16_peak / 6.0
            out["qr_f32_n16384"]["roofline"] = {"bound": "tensor", "achieved": tf32, "peak": peak32, "unit": "TFLOP/s", "frac": tf32 / peak32,
                                                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 rate) / 3 (3xTF32)"}
        if oracle is not None:
            out["qr_f32_n16384"]["oracle_leading_panel"] = _panel_check(np, torch, oracle, (dA, stau), src, 128, False)
        del src, dA, stau
        torch.cuda.empty_cache()
        # config 1: n = 1024
        n = 1024
        src = torch.randn((n, n), device=dev, dtype=torch.float64)
        dA = torch.empty_like(src)
        best = 1e30
        for _ in range(6):
            dA.copy_(src)
            ms, _ = _time(torch, lambda: g.qr_blocked_dev(dA.data_ptr(), n, n, n, dtau.data_ptr(), 0, stream), reps=1)
            best = min(best, ms)
        tf = 4.0 / 3.0 * n ** 3 / (best * 1e-3) / 1e12
        out["qr_f64_n1024"] = {"ms": best, "tflops": tf, "roofline": roof(tf), "e2e_ms": e2e_qr(n, torch.float64, np.float64)}
        # config 2: Cholesky n = 4096
        n = 4096
        X = torch.randn((n, n), device=dev, dtype=torch.float64)
        S = X.t() @ X + n * torch.eye(n, device=dev, dtype=torch.float64)
        dS = torch.empty_like(S)
        info = torch.zeros(1, device=dev, dtype=torch.int32)
        best = 1e30
        for _ in range(6):
            dS.copy_(S)
            ms, _ = _time(torch, lambda: g.chol_recursive_dev(dS.data_ptr(), n, n, info.data_ptr(), 1, stream), reps=1)
            best = min(best, ms)
        tf = n ** 3 / 3.0 / (best * 1e-3) / 1e12
        L = torch.tril(dS.t())
        resid = ((L @ L.t() - S).norm() / S.norm()).item()
        hS = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        e2e = 1e30
        for _ in range(2):
            hS.copy_(S)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            g.cholRecursive_(hS.numpy().T)     # a pinned host matrix through the host-pointer ABI (symmetric: layout is immaterial)
            e2e = min(e2e, (time.perf_counter() - t0) * 1e3)
        out["chol_f64_n4096"] = {"ms": best, "tflops": tf, "info": int(info.item()), "residual": resid, "roofline": roof(tf),
                                 "e2e_ms": e2e}
        del X, S, dS, L, hS
        torch.cuda.empty_cache()
        # f3 rows: two-sided reductions n = 2048 (BLAS-2: every step streams the trailing matrix; L2 resident at this size)
        n = 2048
        src = torch.randn((n, n), device=dev, dtype=torch.float64)
        sym = src + src.t()
        dW = torch.empty_like(src)
        t1 = torch.zeros(n, device=dev, dtype=torch.float64)
        t2 = torch.zeros(n, device=dev, dtype=torch.float64)
        ts = {}
        for name, base, call, passes in (
                ("bidiagonalize", src, lambda: g.bidiagonalize_dev(dW.data_ptr(), n, n, n, t1.data_ptr(), t2.data_ptr(), stream), 6.0),
                ("hessenberg", src, lambda: g.hessenberg_dev(dW.data_ptr(), n, n, t1.data_ptr(), stream), 7.5),
                ("symtri", sym, lambda: g.symtri_dev(dW.data_ptr(), n, n, "L", t1.data_ptr(), stream), 2.0)):
            best = 1e30
            for _ in range(3):
                dW.copy_(base)
                ms, _ = _time(torch, call, reps=1)
                best = min(best, ms)
            # algorithmic traffic: passes x 8 B x sum over the steps of the trailing size (n^3/3 elements; hessenberg's right
            # application runs over all n rows: n^3/2), see DESIGN.md
            ts[name + "_ms"] = best
            ts[name + "_stream_gbps"] = passes * 8.0 * n ** 3 / 3.0 / (best * 1e-3) / 1e9
        out["twosided_f64_n2048"] = dict(ts, note="GB/s = algorithmic trailing-matrix traffic of the unblocked reductions / time "
                                         "(reads for the dots + read-modify-write of the updates); the matrix (32 MiB) is L2 resident")
        if oracle is not None:
            dW.copy_(src)
            g.bidiagonalize_dev(dW.data_ptr(), n, n, n, t1.data_ptr(), t2.data_ptr(), stream)
            torch.cuda.synchronize()
            F, tl, tr, dv, ev, _ = oracle.bidiagonalize(np.asfortranarray(src.cpu().numpy().T))
            got = dW.cpu().numpy().T
            out["twosided_f64_n2048"]["oracle_bidiagonal_max_abs_err"] = float(max(np.max(np.abs(np.diagonal(got) - dv)),
                                                                                   np.max(np.abs(np.diagonal(got, 1) - ev))))
            out["twosided_f64_n2048"]["oracle_factors_max_abs_err"] = float(np.max(np.abs(got - F)))
        del src, sym, dW
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
    # parity of the COMBINED R: Gram identity against the all-reduced A^T A of every shard (sign independent), and every
    # rank must hold bitwise the same R
    Rc = torch.triu(R.t())
    G = A @ A.t()
    if world > 1:
        dist.all_reduce(G)
    gram = ((Rc.t() @ Rc - G).abs().amax() / G.abs().amax()).item()
    same = True
    if world > 1:
        allR = [torch.empty_like(R) for _ in range(world)]
        dist.all_gather(allR, R)
        same = all(bool(torch.equal(allR[0], x)) for x in allR)
    tf = 2.0 * m_total * n * n / (ms * 1e-3) / 1e12
    out["tsqr_f64_8388608x64"] = {"ms": ms, "rows_per_s": m_total / (ms * 1e-3), "gb_per_s": m_total * n * 8 / (ms * 1e-3) / 1e9,
                                  "tflops": tf, "scaling": "strong", "gram_check_combined_R": gram,
                                  "combined_R_bitwise_identical_on_all_ranks": same,
                                  "roofline": {"bound": "fp64 (AI 16 flop/B)", "achieved": tf, "peak": FP64_TENSOR_PEAK_TFLOPS * world,
                                               "unit": "TFLOP/s", "frac": tf / (FP64_TENSOR_PEAK_TFLOPS * world),
                                               "hbm_frac": m_total * n * 8 / (ms * 1e-3) / 1e9 / (_peaks()[0] * world)},
                                  "collective": "ncclAllGather of 64x64 R factors inside gla_dtsqr_allreduce_dev" if world > 1 else "none"}
    if world == 1:
        # end to end through the host-pointer entry point gla_dtsqr on a pinned host matrix (4.3 GB): the call streams 2^20-row
        # chunks (H2D under the reduction of the previous chunk), so this is the upload rate of the box
        hA = torch.empty((n, rows), dtype=torch.float64, pin_memory=True)
        hA.copy_(A)
        torch.cuda.synchronize()
        e2e = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            Rh = g.tsqr_R(hA.numpy().T)
            e2e = min(e2e, (time.perf_counter() - t0) * 1e3)
        out["tsqr_f64_8388608x64"]["e2e_ms"] = e2e
        out["tsqr_f64_8388608x64"]["e2e_host_gb_per_s"] = m_total * n * 8 / (e2e * 1e-3) / 1e9
        out["tsqr_f64_8388608x64"]["e2e_gram_check"] = float(np.max(np.abs(Rh.T @ Rh - G.cpu().numpy())) / G.abs().amax().item())
        del hA
    if oracle is not None and world == 1:
        # oracle on a leading row block is not comparable (R depends on all rows); compare |R| of a 2^16-row problem instead
        mm = 1 << 16
        sub = np.asfortranarray(A[:, :mm].cpu().numpy().T.copy())
        Rs = torch.zeros((n, n), device=dev, dtype=torch.float64)
        g.tsqr_local_dev(A.data_ptr(), mm, n, rows, Rs.data_ptr(), n, stream)
        torch.cuda.synchronize()
        rf, _ = oracle.qr_blocked(sub, 12)
        Ro = np.triu(rf[:n])
        Rg = np.triu(Rs.cpu().numpy().T)
        D = np.sign(np.diag(Ro)) * np.sign(np.diag(Rg))       # row-phase normalisation (a TSQR tree does not see the sequential pivots)
        out["tsqr_f64_8388608x64"]["oracle_65536x64_max_rel_R"] = float(np.max(np.abs(D[:, None] * Rg - Ro)) / np.max(np.abs(Ro)))
    if comm is not None:
        torch.cuda.synchronize()
        comm.destroy()
    if cpu:
        for k, v in cpu_other_configs().items():
            if k in out:
                out[k]["cpu_baseline"] = v
    return out


if __name__ == "__main__":
    main()
