#!/usr/bin/env python
"""Benchmark of the token-grouping hot path (BASELINE.json metric: images/sec of NCut token clustering + pooling).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one pass of the whole hot path (affinity+degree -> top-k NCut eigenvectors -> k-means -> label
composition -> cluster-mean pooling) over one batch of synthetic planted-mixture ViT tokens.  At N=1 the
workload is BASELINE.json configs[1] (ViT-B/16, 196 tokens, d=768, batch 1024, k=8).  For N>1 every rank runs
the same per-GPU batch on its own images (weak scaling, no collective on the data path); the step time is
the max over ranks, `value` = images all ranks processed / that time.

Printed JSON (rank 0, one line): the base contract keys plus
  roofline      the dominant kernel of the step (by CUDA-event share): algorithmic bytes (or flops) per launch
                / its average launch duration, against MEASURED_PEAKS.json
  stages        per-stage average ms and achieved rate (same events)
  cpu_baseline  the CPU oracle (a port: the reference's arithmetic lives in packages absent from the image)
                timed on this box's host cores on a bounded sample of the same workload
  e2e           same metric through the public host-buffer API (msvit.HostClusterer): pinned host tokens ->
                H2D -> kernels -> D2H of labels / pooled tokens / counts, all inside the timed region
`--impl reference` times the CPU oracle alone (rank 0 only) and prints the same line shape.

Only the cpu_baseline leg and `--impl reference` import oracle/; the product path never does.
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
for _p in (ROOT, os.path.join(ROOT, "multi-state-vit_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "images/sec NCut token clustering+pooling, ViT-B/16 196 tok, at 1/2/4/8 B200"
UNIT = "images/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        p["_source"] = "measured (MEASURED_PEAKS.json)"
        return p
    p = dict(FALLBACK_PEAKS)
    p["_source"] = "fallback (B200_PROFILING.md)"
    return p


def workload(name: str):
    from msvit.synthetic import CONFIGS
    B, N, D, K = CONFIGS[name]
    k = {"C1": 8, "C2": 8, "C3": 16, "C4": 8}[name]
    return B, N, D, K, k


def workload_string(name, B, N, D, K, k, dtype_name):
    model = {"C1": "ViT-B/16 224px", "C2": "ViT-B/16 224px", "C3": "ViT-L/14 336px", "C4": "ViT-B/14 448px (1024 tokens)"}[name]
    return (f"{name}: {model} tokens [B={B} per GPU, N={N}, D={D}] {dtype_name}, rbf NCut affinity "
            f"(gamma=3, scale=D/4), k={k} eigenvectors, K={K} k-means clusters, cluster-mean pooling")


# ------------------------------------------------------------------------------------------ algorithmic work
def stage_work(B, N, D, K, k, esz, fused=False, block=16, eig_iters=0.0):
    """Algorithmic bytes / flops of ONE launch of each stage kernel over B images (DESIGN.md section 4;
    per-image figures are SURVEY.md section 8d's)."""
    lda = (N + 3) & ~3
    w = {}
    if fused:
        # fused affinity + subspace iteration: read X once; out: degree, the basis U [N, 16] and H [16, 16].  The
        # affinity never leaves tensor memory.  flops: Gram 2 N^2 D + (products 2 N^2 m + Grams 4 N m^2) per iteration
        w["affinity"] = {"bytes": B * (esz * N * D + 4 * N + 4 * N * block + 4 * block * block),
                         "flops": B * (2.0 * N * N * D + eig_iters * (2.0 * N * N * block + 4.0 * N * block * block))}
        # Ritz rotation + k-means: read U, H, deg; write V, lambda, int64 labels, child count
        w["kmeans"] = {"bytes": B * (4 * N * block + 4 * block * block + 4 * N + 4 * N * k + 4 * k + 8 * N + 4),
                       "flops": 0.0}
    else:
        # affinity: read X once, write A (fp32, padded rows) and deg; flops = Gram 2 N^2 D
        w["affinity"] = {"bytes": B * (esz * N * D + 4 * N * lda + 4 * N), "flops": 2.0 * B * N * N * D}
        # eigensolver: compulsory traffic = A and deg once, V and lambda out (A is re-read from L2 every iteration)
        w["eig"] = {"bytes": B * (4 * N * lda + 4 * N + 4 * N * k + 4 * k), "flops": 0.0}
        # k-means: read V, lambda, deg; write labels (int32) and the child count
        w["kmeans"] = {"bytes": B * (4 * N * k + 4 * k + 4 * N + 4 * N + 4), "flops": 0.0}
        # label composition: read int32 labels + counts, write int64 labels
        w["compose"] = {"bytes": B * (4 * N + 4 + 8 * N), "flops": 0.0}
    # pooling: read X and int64 labels, write pooled fp32 and counts
    w["pool"] = {"bytes": B * (esz * N * D + 8 * N + 4 * K * D + 4 * K), "flops": 0.0}
    return w


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi sampled every 100 ms while the timed region runs (B200_PROFILING.md clocks line)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None          # (module, handle) when NVML is importable: sub-ms polling instead of nvidia-smi's 100 ms
        self.samples = []
        self.running = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(index))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        power, n = 0.0, 0
        while self.running:
            try:
                if n % 16 == 0:          # the power query is the slow one
                    power = nv.nvmlDeviceGetPowerUsage(h) / 1e3
                self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, power, int(reasons_fn(h))))
            except Exception:
                pass
            n += 1
            time.sleep(0.0005)

    def sample_now(self):
        """One sample taken by the calling thread (used after the last launch of the timed region has been enqueued,
        while the GPU is still working through it)."""
        if self.nvml is None or not self.running:
            return
        nv, h = self.nvml
        try:
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                                 nv.nvmlDeviceGetPowerUsage(h) / 1e3, int(reasons_fn(h))))
        except Exception:
            pass

    def start(self):
        if self.nvml is not None:
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.running = False
            self.thread.join(timeout=2)
            if not self.samples:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            sm = sorted(x[0] for x in self.samples)
            bits = 0
            for x in self.samples:
                bits |= x[3]
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            return {"sm_mhz": float(sm[len(sm) // 2]), "sm_max_mhz": float(max(x[1] for x in self.samples)),
                    "reasons": sorted(nm for bit, nm in names.items() if bits & bit), "samples": len(sm),
                    "power_w_max": round(max(x[2] for x in self.samples), 2), "source": "nvml polled inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_rate(N, D, K, k, n_images: int, batch: int = 8, threads=None):
    """Times the CPU oracle (torch CPU restatement, all host threads unless `threads` says otherwise) on `n_images`
    images of the workload, processed in batches of `batch` (BASELINE.json configs[0] is batch 8).
    Returns (images/s, threads, seconds, images)."""
    from oracle import ncut_oracle as O
    from msvit.synthetic import default_scale, planted_tokens
    cores = threads or os.cpu_count() or 1
    torch.set_num_threads(cores)
    pool_n = min(n_images, 64)
    x, _ = planted_tokens(pool_n, N, D, K)
    s = default_scale(D)

    def one(b0):
        xb = x[b0:b0 + batch]
        child, _, _, _ = O.cluster_tokens(xb, None, ncut_dim=k, n_clusters=K, scale=s)
        O.pool(xb, child, K)
        return xb.shape[0]

    one(0)  # warm-up
    done = 0
    t0 = time.perf_counter()
    while done < n_images:
        done += one((done % pool_n) // batch * batch if pool_n >= batch else 0)
    dt = time.perf_counter() - t0
    return done / dt, torch.get_num_threads(), dt, done


def torch_cuda_eager_rate(N, D, K, k, n_images: int, dev, gamma=3.0):
    """BASELINE.md section 3's extra reference row: the reference path's arithmetic as plain eager torch on the GPU
    (cuBLAS matmul + elementwise exp + cuSOLVER eigh + eager k-means + one-hot matmul pooling), fp32, CUDA-event
    timed.  None of this repo's kernels run here.  Returns (images/s, ms)."""
    from msvit.synthetic import default_scale, planted_tokens
    x, _ = planted_tokens(min(n_images, 64), N, D, K)
    x = x.repeat((n_images + x.shape[0] - 1) // x.shape[0], 1, 1)[:n_images].to(dev)
    scale = default_scale(D)

    def step():
        sq = (x * x).sum(-1)
        d2 = (sq[:, :, None] + sq[:, None, :] - 2.0 * torch.bmm(x, x.transpose(1, 2))).clamp_min(0) / scale
        A = torch.exp(-d2 / gamma)
        deg = A.sum(-1)
        dis = deg.rsqrt()
        lam, vec = torch.linalg.eigh(A * dis[:, :, None] * dis[:, None, :])
        V = vec[:, :, -k:].flip(-1)[:, :, :K]                      # leading eigenvectors, descending
        # farthest-point seeding + Lloyd, batched
        first = deg.argmax(1)
        bidx = torch.arange(x.shape[0], device=dev)
        cen = V[bidx, first][:, None, :]
        mind = ((V - cen) ** 2).sum(-1)
        for _ in range(1, K):
            nxt = mind.argmax(1)
            c = V[bidx, nxt][:, None, :]
            cen = torch.cat([cen, c], 1)
            mind = torch.minimum(mind, ((V - c) ** 2).sum(-1))
        lab = None
        for _ in range(20):
            new = torch.cdist(V, cen).argmin(-1)
            if lab is not None and torch.equal(new, lab):
                break
            lab = new
            oh = torch.nn.functional.one_hot(lab, K).to(V.dtype)
            cnt = oh.sum(1)
            cen = torch.where(cnt[:, :, None] > 0, torch.bmm(oh.transpose(1, 2), V) / cnt.clamp_min(1)[:, :, None], cen)
        oh = torch.nn.functional.one_hot(lab, K).to(x.dtype)
        pooled = torch.bmm(oh.transpose(1, 2), x) / oh.sum(1).clamp_min(1)[:, :, None]
        return pooled

    for _ in range(2):
        step()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize(dev)
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[1]
    return n_images / ms * 1e3, ms


def run_reference(args):
    """`--impl reference`: the reference's CPU path for this workload.  The reference's own arithmetic
    (ncut-pytorch / cuML) is not installable here, so this is the oracle port, on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B, N, D, K, k = workload(args.config)
    per_step = args.ref_images_per_step
    for _ in range(args.warmup):
        cpu_oracle_rate(N, D, K, k, 8)
    t_total, n_total, cores = 0.0, 0, 1
    for _ in range(args.steps):
        rate, cores, dt, done = cpu_oracle_rate(N, D, K, k, per_step)
        t_total += dt
        n_total += done
    value = n_total / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.config, B, N, D, K, k, args.dtype),
                   "sample_per_step": f"{per_step} images in batches of 8 (CPU oracle port of the reference path)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{n_total} images of {args.config} in batches of 8, torch CPU {cores} threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch.distributed as dist
    import msvit
    from msvit.functional import ClusterPlan, HostClusterer
    from msvit.sharding import max_over_ranks
    from msvit.synthetic import default_scale, planted_tokens

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    msvit._lib.load()  # fail loudly if libmsvit.so is missing
    # host threads and the pinned buffers they allocate stay on the NUMA node of this rank's GPU
    from msvit.sharding import bind_host_thread_to_gpu
    numa_cores = bind_host_thread_to_gpu(local)

    B, N, D, K, k = workload(args.config)
    dtype = torch.float32 if args.dtype == "float32" else torch.bfloat16
    esz = 4 if dtype == torch.float32 else 2
    scale = default_scale(D)

    # this rank's images (weak scaling: B per GPU, image ids rank*B .. rank*B+B-1), generated on the host
    host = torch.empty(B, N, D, dtype=dtype).pin_memory()
    planted_tokens(B, N, D, K, first=rank * B, out=host) if dtype == torch.float32 else host.copy_(
        planted_tokens(B, N, D, K, first=rank * B)[0])
    x = host.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    plan = ClusterPlan(B, N, D, dtype, dev, ncut_dim=k, n_clusters=K, scale=scale)
    n_st = len(ClusterPlan.STAGES)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def reduce_max(v: float) -> float:
        return max_over_ranks(v, dev)

    # ---- device-resident timing: K steps, per-stage events recorded inside the same region
    for _ in range(args.warmup):
        out = plan.run(x)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(n_st + 1)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_begin.record()
    for s in range(args.steps):
        out = plan.run(x)
    t_end.record()
    # the launches are enqueued far ahead of the GPU: sample the clocks from this thread until the timed region has
    # drained (an NVML query takes about a millisecond, so it must not sit between the launches)
    if rank == 0:
        while not t_end.query():
            sampler.sample_now()
    barrier()
    ms_total = reduce_max(t_begin.elapsed_time(t_end))
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel durations: the same K steps again with an event after every launch.  The events are kept out of the
    # timed region above: seven records per step cost 18 us of the 0.5 ms step (tools/microbench/event_overhead.py)
    for s in range(args.steps):
        out = plan.run(x, events=events[s])
    barrier()
    stage_ms = {name: sum(ev[i].elapsed_time(ev[i + 1]) for ev in events) / args.steps
                for i, name in enumerate(ClusterPlan.STAGES)}
    iters = out.iters.float()
    eig_iters = {"mean": float(iters.mean()), "max": int(iters.max())}

    # ---- end to end through the public host-buffer API
    hc = HostClusterer(B, N, D, dtype, dev, ncut_dim=k, n_clusters=K, scale=scale, chunk=args.e2e_chunk)
    for _ in range(max(1, min(args.warmup, 3))):
        res = hc.run(host)
    barrier()
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = args.e2e_steps or args.steps
    barrier()
    wall0 = time.perf_counter()
    e_begin.record()
    for s in range(e2e_steps):
        res = hc.run(host)
    e_end.record()
    barrier()
    e2e_wall = time.perf_counter() - wall0
    e2e_ms = reduce_max(max(e_begin.elapsed_time(e_end), 1e3 * e2e_wall))
    h2d_bytes, d2h_bytes = hc.h2d_bytes, hc.d2h_bytes   # of the fp32 run (hc is rebuilt for the bf16 leg below)
    # the end-to-end results agree with the device-resident ones
    if not torch.equal(res.labels, out.labels.cpu()):
        raise RuntimeError("end-to-end labels differ from the device-resident run")
    # same call with the host tokens held in bf16 (the caller's choice of a narrower host dtype): half the H2D bytes,
    # kind::f16 tensor-core path on the device
    e2e_bf16 = None
    if dtype == torch.float32 and args.extras:
        del hc
        host16 = host.to(torch.bfloat16).pin_memory()
        hc = HostClusterer(B, N, D, torch.bfloat16, dev, ncut_dim=k, n_clusters=K, scale=scale, chunk=args.e2e_chunk)
        for _ in range(2):
            res16 = hc.run(host16)
        barrier()
        b0_, b1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        b0_.record()
        for s in range(e2e_steps):
            res16 = hc.run(host16)
        b1_.record()
        barrier()
        ms16 = reduce_max(max(b0_.elapsed_time(b1_), 1e3 * (time.perf_counter() - w0)))
        e2e_bf16 = {"value": round(world * B * e2e_steps / ms16 * 1e3, 1), "unit": UNIT, "h2d_bytes_per_step": hc.h2d_bytes,
                    "d2h_bytes_per_step": hc.d2h_bytes, "ms_per_step": round(ms16 / e2e_steps, 3),
                    "labels_equal_fp32_run": bool(torch.equal(res16.labels, res.labels))}
        del host16
        # device-resident step with bf16 tokens (north_star's "TMA-staged bf16 tiles"): kind::f16 Gram, half the bytes
        x16 = x.to(torch.bfloat16)
        plan16 = ClusterPlan(B, N, D, torch.bfloat16, dev, ncut_dim=k, n_clusters=K, scale=scale)
        for _ in range(args.warmup):
            out16 = plan16.run(x16)
        ev16 = [[torch.cuda.Event(enable_timing=True) for _ in range(n_st + 1)] for _ in range(args.steps)]
        barrier()
        b0_.record()
        for s in range(args.steps):
            out16 = plan16.run(x16)
        b1_.record()
        barrier()
        ms16d = reduce_max(b0_.elapsed_time(b1_)) / args.steps
        for s in range(args.steps):
            out16 = plan16.run(x16, events=ev16[s])
        barrier()
        e2e_bf16["device_resident"] = {
            "value": round(world * B / ms16d * 1e3, 1), "unit": UNIT, "ms_per_step": round(ms16d, 4),
            "stages_ms": {name: round(sum(ev[i].elapsed_time(ev[i + 1]) for ev in ev16) / args.steps, 4)
                          for i, name in enumerate(ClusterPlan.STAGES) if any(ev[i].elapsed_time(ev[i + 1]) > 0 for ev in ev16)},
            "labels_equal_fp32_run": bool(torch.equal(out16.labels, out.labels)),
            "note": "same step as `value` with the tokens held in bf16 on the device"}
        del x16, plan16, out16

    # ---- the other BASELINE.json configs as sub-records (C3, C4 with its three levels, C5 with its NCCL all-reduce)
    extras = {}
    if args.extras and args.config == "C2":
        del hc
        torch.cuda.empty_cache()
        for nm in ("C3", "C4"):
            try:
                extras[nm] = extra_per_image(nm, dev, world, rank, max(3, args.steps // 4), 3, barrier, reduce_max)
            except Exception as exc:  # a sub-record must not take the headline down
                extras[nm] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            extras["C2_smooth_spectrum"] = smooth_spectrum_record(B, N, D, K, k, scale, dev, world, rank,
                                                                  max(3, args.steps // 4), barrier, reduce_max)
        except Exception as exc:
            extras["C2_smooth_spectrum"] = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            c5 = c5_measure(args, dev, world, rank, local, max(5, args.steps // 2), 3, with_e2e=False)
            if c5 is not None:
                extras["C5"] = {k_: c5[k_] for k_ in ("metric", "value", "unit", "ms_per_step", "scaling", "roofline", "stages")}
                extras["C5"]["workload"] = c5["config"]["workload"]
        except Exception as exc:
            extras["C5"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    work = stage_work(B, N, D, K, k, esz, fused=plan.fused, block=plan.block, eig_iters=eig_iters["mean"])
    kernel_of = ({"affinity": "ncut_fused_kernel (affinity + subspace iteration, affinity kept in TMEM)",
                  "kmeans": "ritz_kmeans_kernel (Rayleigh-Ritz + k-means + labels)", "pool": "pool_kernel"}
                 if plan.fused else {n: n for n in ClusterPlan.STAGES})
    stages = {}
    for name in ClusterPlan.STAGES:
        ms = stage_ms[name]
        if name not in work or ms <= 0:
            continue
        w = work[name]
        stages[name] = {"kernel": kernel_of.get(name, name), "ms": round(ms, 4), "GB/s": round(w["bytes"] / ms / 1e6, 1)}
        if w["flops"]:
            stages[name]["TFLOP/s"] = round(w["flops"] / ms / 1e9, 1)
    dominant = max(stages, key=lambda n: stages[n]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tkey = {"affinity": "ncut_fused", "kmeans": "ritz_kmeans"}.get(dominant, dominant) if plan.fused else dominant
            traffic = json.load(f).get(args.config, {}).get(tkey)
    if dominant == "affinity" and not plan.fused:
        # the Gram contraction is the one tensor-core-bound kernel of the path
        ach = stages[dominant]["TFLOP/s"]
        peak = peaks["bf16_tflops_sustained"]
        roof = {"kernel": kernel_of[dominant], "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peaks["_source"] + ", sustained bf16"}
    else:
        ach = stages[dominant]["GB/s"]
        peak = peaks["hbm_gbs"]
        roof = {"kernel": kernel_of[dominant], "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4), "traffic": traffic, "peak_source": peaks["_source"]}
    roof["share_of_step"] = round(stages[dominant]["ms"] / sum(s["ms"] for s in stages.values()), 3)
    if dominant == "affinity" and plan.fused:
        roof["note"] = ("compulsory HBM traffic = the tokens once (the affinity stays in tensor memory); the kernel is bound by "
                        "the per-image dependency chain of the subspace iteration (one image per SM at a time), "
                        "see profiles/r2_summary.md")
        roof["tensor_tflops"] = stages[dominant].get("TFLOP/s")
    if dominant == "eig":
        # the eigensolver is bound by per-segment dependency latency, not by a pipe (profiles/r1c_summary.md): its
        # HBM figure is the compulsory traffic (affinity once); its tensor-core products are reported next to it
        m = plan.block
        fl = 2.0 * B * N * N * m * eig_iters["mean"]
        roof["note"] = "latency-bound (4 CTAs/SM, two waves); hbm = compulsory bytes (affinity read once)"
        roof["products_tflops"] = round(fl / stages["eig"]["ms"] / 1e9, 2)

    cpu_rate, cores, cpu_s, cpu_n = cpu_oracle_rate(N, D, K, k, args.cpu_images)

    ms_step = ms_total / args.steps
    value = world * B / ms_step * 1e3
    n_launch_step = sum(1 for n in ClusterPlan.STAGES if stage_ms[n] > 0 and n in work)
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32" if dtype == torch.float32 else "bf16", "data": "synthetic",
        "config": {"workload": workload_string(args.config, B, N, D, K, k, args.dtype),
                   "global_batch": world * B, "parallelism": f"batch-sharded x{world}, no collective",
                   "l2_policy": f"inputs larger than L2 ({B * N * D * esz / 1e6:.0f} MB tokens per step vs 126 MB L2)",
                   "path": "fused (affinity in tensor memory)" if plan.fused else "affinity + eig kernels (affinity through L2)",
                   "gram_operands": ("fp16, converted in place from the fp32 tokens: the 11-bit significand TF32 keeps; fp32 "
                                     "accumulation, fp32 everywhere else" if (plan.fused and dtype == torch.float32)
                                     else ("bf16 tokens" if dtype == torch.bfloat16 else "tf32")),
                   "eig_iters": eig_iters,
                   "stage_timing": "`stages` / `roofline` come from a second pass of the same K steps with a CUDA event after "
                                   "every launch; the timed region holds the launches only (the event records cost 18 us per step)"},
        "roofline": roof, "stages": stages,
        "cpu_baseline": {"value": round(cpu_rate, 2), "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cpu_n} images of {args.config} in batches of 8 ({cpu_s:.1f} s), "
                                   f"torch CPU oracle, {cores} threads"},
        "e2e": {"value": round(world * B * e2e_steps / e2e_ms * 1e3, 1), "unit": UNIT,
                "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes, "steps": e2e_steps,
                "ms_per_step": round(e2e_ms / e2e_steps, 3), "api": "msvit.HostClusterer.run (pinned host buffers)",
                "h2d_GB/s_per_gpu": round(h2d_bytes * e2e_steps / e2e_ms / 1e6, 1),
                "host_cores_bound": None if numa_cores is None else len(numa_cores),
                "note": "fp32 host tokens (the reference hands fp32 hidden states): 617 MB per GPU and step cross PCIe, "
                        "which is the ceiling of this number (profiles/r2_h2d_probe.md); e2e_bf16_host halves the bytes"},
        "e2e_bf16_host": e2e_bf16,
        "gpu_launches": n_launch_step * args.steps,
        "clocks": clocks,
    }
    if extras:
        line["extra_configs"] = extras
    if rank == 0 and args.extras:
        # the single-thread CPU row and the plain torch-CUDA row, beside cpu_baseline
        r1, _, s1, n1 = cpu_oracle_rate(N, D, K, k, max(8, args.cpu_images // 32), threads=1)
        n_eager = min(B, 256)
        re, ms_e = torch_cuda_eager_rate(N, D, K, k, n_eager, dev)
        line["extra_baselines"] = {
            "cpu_oracle_1_thread": {"value": round(r1, 2), "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n1} images of {args.config} in batches of 8 ({s1:.1f} s)"},
            "torch_cuda_eager": {"value": round(re, 1), "unit": UNIT, "kind": "plain torch on 1 B200 (cuBLAS bmm, exp, "
                                 "cuSOLVER eigh, eager k-means, one-hot bmm pooling), fp32, none of this repo's kernels",
                                 "sample": f"{n_eager} images of {args.config} in one batch, median of 3 ({ms_e:.1f} ms)"},
        }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------ the other configs
def smooth_spectrum_record(B, N, D, K, k, scale, dev, world, rank, steps, barrier, reduce_max):
    """The C2 step on tokens WITHOUT planted clusters (msvit.synthetic.smooth_tokens: smooth random fields over the patch
    grid, NCut eigenvalues decaying gradually): iteration counts and throughput where the solver has no spectral gap to
    lean on.  Same kernels, same plan arguments as the headline."""
    from msvit.functional import ClusterPlan
    from msvit.synthetic import smooth_tokens
    pool_n = 64
    xs = smooth_tokens(pool_n, N, D, first=rank * pool_n)
    x = xs.repeat((B + pool_n - 1) // pool_n, 1, 1)[:B].contiguous().to(dev)
    plan = ClusterPlan(B, N, D, torch.float32, dev, ncut_dim=k, n_clusters=K, scale=scale)
    for _ in range(3):
        out = plan.run(x)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(steps):
        out = plan.run(x)
    t1.record()
    barrier()
    ms = reduce_max(t0.elapsed_time(t1)) / steps
    it = out.iters.float()
    return {"value": round(world * B / ms * 1e3, 1), "unit": UNIT, "ms_per_step": round(ms, 4), "scaling": "weak",
            "workload": f"C2 shape [B={B} per GPU, N={N}, D={D}] float32, smooth-spectrum tokens (no planted clusters), "
                        f"k={k} eigenvectors, K={K} k-means clusters, cluster-mean pooling",
            "path": "fused (affinity in tensor memory)" if plan.fused else "affinity + eig kernels",
            "eig_iters": {"mean": round(float(it.mean()), 2), "max": int(it.max())},
            "converged_fraction": round(float(out.converged.float().mean()), 4)}


def extra_per_image(name, dev, world, rank, steps, warmup, barrier, reduce_max):
    """C3 (ViT-L/14, 576 tokens, k=16) and C4 (1024 tokens, THREE hierarchical levels: 1 -> 4 -> 16 parents per image,
    every parent segment re-clustered into 4 children from that level's hidden states) as sub-records of the main
    line: device-resident images/s per GPU batch, weak scaling, per-stage ms."""
    from msvit.functional import ClusterPlan
    from msvit.synthetic import default_scale, hierarchical_tokens, planted_tokens
    B, N, D, K, k = workload(name)
    pool_n = min(B, 32)
    if name == "C4":
        # a planted 4 x 4 x 4 tree: every level of re-clustering has four sub-clusters to find (a flat mixture makes
        # levels 1 and 2 cluster iid noise, which only measures the iteration cap)
        xs, _ = hierarchical_tokens(pool_n, N, D, branch=K, depth=3, first=rank * pool_n)
    else:
        xs, _ = planted_tokens(pool_n, N, D, K, first=rank * pool_n)
    x = xs.repeat((B + pool_n - 1) // pool_n, 1, 1)[:B].contiguous().to(dev)
    scale = default_scale(D)
    if name != "C4":
        plans = [ClusterPlan(B, N, D, torch.float32, dev, ncut_dim=k, n_clusters=K, scale=scale)]
        levels = [x]
    else:
        # level l clusters every parent of level l-1 into K children: P = 1, K, K^2 parents
        plans = [ClusterPlan(B, N, D, torch.float32, dev, ncut_dim=k, n_clusters=K, scale=scale, n_parents=K ** l,
                             want_pool=(l == 2), pool_k=K ** 3) for l in range(3)]
        g = torch.Generator(device=dev).manual_seed(77 + rank)
        levels = [x] + [x + 0.1 * torch.randn(x.shape, generator=g, device=dev) for _ in range(2)]

    def step(ev=None):
        parent = None
        outs = []
        for l, plan in enumerate(plans):
            out = plan.run(levels[l], parent, events=None if ev is None else ev[l])
            parent = out.labels
            outs.append(out)
        return outs

    for _ in range(warmup):
        outs = step()
    barrier()
    n_st = len(ClusterPlan.STAGES)
    events = [[[torch.cuda.Event(enable_timing=True) for _ in range(n_st + 1)] for _ in plans] for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for s_ in range(steps):
        outs = step()
    t1.record()
    barrier()
    ms = reduce_max(t0.elapsed_time(t1)) / steps
    for s_ in range(steps):          # per-kernel durations from a second, instrumented pass (see run_ours)
        outs = step(events[s_])
    barrier()
    stage_ms = {}
    for l in range(len(plans)):
        for i, nm in enumerate(ClusterPlan.STAGES):
            v = sum(ev[l][i].elapsed_time(ev[l][i + 1]) for ev in events) / steps
            if v > 0:
                stage_ms[nm if len(plans) == 1 else f"L{l}.{nm}"] = round(v, 4)
    rec = {"value": round(world * B / ms * 1e3, 1), "unit": UNIT, "ms_per_step": round(ms, 4), "scaling": "weak",
           "workload": workload_string(name, B, N, D, K, k, "float32") +
                       (" -- 3 hierarchical levels (1, 4, 16 parents per image), tokens from a planted 4x4x4 tree" if name == "C4" else ""),
           "stages_ms": stage_ms, "eig_iters_mean": [round(float(o.iters.float().mean()), 2) for o in outs],
           "converged": [bool(o.converged.all()) for o in outs]}
    if name == "C4":
        rec["children_per_image"] = [round(float((o.labels.max(dim=1).values + 1).float().mean()), 2) for o in outs]
    del plans, levels, x
    torch.cuda.empty_cache()
    return rec


# ------------------------------------------------------------------------------------------ C5: dataset-level k-means
C5_METRIC = "rows/sec per Lloyd iteration, global k-means k=1000 over 1M x 768 features (BASELINE.json configs[4])"


def run_c5(args):
    """`--config C5`: one step = one Lloyd iteration (assign -> sort -> accumulate -> all-reduce -> finalize) over
    all 1 000 000 rows, sharded over the ranks (strong scaling: the total is fixed); the only collective is the
    all-reduce of the packed [k, D+1] sums|counts buffer (3.08 MB) over NCCL."""
    import torch.distributed as dist
    import msvit

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    msvit._lib.load()
    line = c5_measure(args, dev, world, rank, local, args.steps, args.warmup, with_e2e=True)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def c5_measure(args, dev, world, rank, local, steps, warmup, with_e2e):
    """The C5 measurement proper (all ranks call it; rank 0 gets the record, the others None)."""
    import torch.distributed as dist
    from msvit.global_kmeans import GlobalKMeansPlan, broadcast_init, global_kmeans
    from msvit.sharding import max_over_ranks, shard_bounds
    n_total, D, k = args.c5_rows, 768, 1000
    first, n = shard_bounds(n_total, rank, world)
    dtype = torch.bfloat16   # the dataset-level path takes bf16 features
    # planted mixture: centres1000[label] + 0.5 randn, seed 1212; every rank generates its own rows
    g = torch.Generator(device=dev).manual_seed(1212)
    centres = torch.randn(k, D, generator=g, device=dev)
    g.manual_seed(1212 + 7919 * (rank + 1))
    x = torch.empty(n, D, dtype=dtype, device=dev)
    for r0 in range(0, n, 1 << 16):
        r1 = min(n, r0 + (1 << 16))
        lab = torch.randint(0, k, (r1 - r0,), generator=g, device=dev)
        x[r0:r1] = (centres[lab] + 0.5 * torch.randn(r1 - r0, D, generator=g, device=dev)).to(dtype)
    plan = GlobalKMeansPlan(n, D, k, dtype, dev)
    plan.set_centroids(broadcast_init(x, k))
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    names = ("assign", "sort", "accumulate", "allreduce", "finalize")

    def step(ev=None):
        packed = plan.local_step(x, events=ev)
        if world > 1:
            dist.all_reduce(packed)
        if ev is not None:
            ev[4].record()
        plan.finalize(packed)
        if ev is not None:
            ev[5].record()

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    events = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for s in range(steps):
        step(events[s])
    t1.record()
    barrier()
    ms_total = max_over_ranks(t0.elapsed_time(t1), dev)
    clocks = sampler.stop() if rank == 0 else None
    stage_ms = {nm: sum(ev[i].elapsed_time(ev[i + 1]) for ev in events) / steps for i, nm in enumerate(names)}

    # end to end: the public call on HOST features (H2D of the shard, `steps` iterations, D2H of centroids and labels)
    e2e = None
    if with_e2e:
        host = x.cpu().pin_memory()
        barrier()
        w0 = time.perf_counter()
        res = global_kmeans(host.to(dev, non_blocking=True), k, steps, init=plan.centroids.clone())
        cent_h, lab_h = res.centroids.cpu(), res.labels.cpu()
        barrier()
        e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - w0), dev)
        e2e = {"value": round(n_total * steps / e2e_ms * 1e3, 1), "unit": "rows/s",
               "h2d_bytes_per_step": host.numel() * 2 // steps,
               "d2h_bytes_per_step": (cent_h.numel() * 4 + lab_h.numel() * 8) // steps,
               "api": "msvit.global_kmeans on host features (one H2D, steps iterations, D2H of centroids + labels)"}
    del x
    if rank != 0:
        return None
    peaks = load_peaks()
    ms_step = ms_total / steps
    flops = 2.0 * n * k * D
    ach = flops / stage_ms["assign"] / 1e9
    line = {
        "metric": C5_METRIC, "value": round(n_total / ms_step * 1e3, 1), "unit": "rows/s", "n_gpus": world,
        "steps": steps, "warmup": warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"C5: {n_total} x {D} bf16 features, k={k}, rows sharded x{world} ({n} on rank 0), "
                               f"one all-reduce of {plan.allreduce_bytes} B per iteration",
                   "l2_policy": f"inputs larger than L2 ({n * D * 2 / 1e6:.0f} MB features per rank vs 126 MB L2)"},
        "roofline": {"kernel": "gkm assign", "bound": "tensor", "achieved": round(ach, 1),
                     "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                     "frac": round(ach / peaks["bf16_tflops_sustained"], 4), "traffic": None,
                     "peak_source": peaks["_source"] + ", sustained bf16",
                     "share_of_step": round(stage_ms["assign"] / sum(stage_ms.values()), 3)},
        "stages": {nm: {"ms": round(v, 4)} for nm, v in stage_ms.items()},
        "e2e": e2e, "gpu_launches": 7 * steps, "clocks": clocks,
    }
    line["stages"]["accumulate"]["GB/s"] = round((n * D * 2 + n * 4 + k * (D + 1) * 4) / stage_ms["accumulate"] / 1e6, 1)
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--c5-rows", type=int, default=1_000_000)
    ap.add_argument("--dtype", default="float32", choices=["float32", "bfloat16"],
                    help="dtype of the token tensor handed to the path (the reference hands fp32 hidden states)")
    ap.add_argument("--cpu-images", type=int, default=2048, help="images in the cpu_baseline sample")
    ap.add_argument("--ref-images-per-step", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    ap.add_argument("--e2e-chunk", type=int, default=148, help="images per H2D chunk (one CTA per SM per chunk)")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the C3 / C4 / C5 sub-records of the default (C2) run")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "C5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
