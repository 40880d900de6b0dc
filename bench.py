#!/usr/bin/env python
"""bench.py -- buildGraph sample-pairs/s on synthetic panels (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            # this repository's B200 path
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle), host cores

A step is one full Conos.buildGraph over the panel (per-pair reduction, projection, exact two-directional kNN,
mutual-neighbour matching, local k.self edges, graph assembly).  `value` is measured with the panel resident
in HBM and the graph left on the device; `e2e` goes through the public host API with the panel in host memory
(uploaded inside the timed region) and the graph copied back.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N = 1
    2: dict(workload="synthetic 8 samples x 50k cells x 2000 od-genes, space=PCA ncomps=30 k=30 (configs[1])",
            n_samples=8, n_cells=50000, n_genes=2000, space="PCA", ncomps=30, k=30, k_self=10),
    # configs[2]: covariance + common-PCA kernels (profiling / parity runs, not a bench line)
    3: dict(workload="synthetic 32 samples x 20k cells, space=CPCA ncomps=50 (configs[2])",
            n_samples=32, n_cells=20000, n_genes=2000, space="CPCA", ncomps=50, k=15, k_self=10),
    # configs[4]: CCA with alignment.strength = 0.3 (k1 = round(0.09 * max cells) >> k: hub paring ladder)
    5: dict(workload="synthetic 2 x 40 samples x 25k cells, space=CCA ncomps=30 alignment.strength=0.3 (configs[4])",
            n_samples=80, n_cells=25000, n_genes=2000, space="CCA", ncomps=30, k=15, k_self=10,
            alignment_strength=0.3),
    # configs[3]: the atlas, pair-sharded
    4: dict(workload="synthetic atlas 100 samples x 10k cells (4950 pairs), space=PCA k=15 k.self=5 (configs[3])",
            n_samples=100, n_cells=10000, n_genes=2000, space="PCA", ncomps=30, k=15, k_self=5),
}


# dram__bytes_read.sum + dram__bytes_write.sum of the cross-kNN launch of this workload (56 problems of 50k x 50k),
# main kernel + finish kernel, from the ncu --set full capture summarised in profiles/ (None until captured)
KNN_DRAM_BYTES_PER_LAUNCH = 0.354947e9 + 0.564263e9 + 1.114550e9 + 0.640961e9
KNN_DRAM_SOURCE = ("profiles/r01_ncu_knn_tc2_bench_config2.txt: dram read+write of knn_tc2_kernel + "
                   "knn_tc2_finish_kernel for the cross-kNN launch (56 problems of 50k x 50k); algorithmic bytes of "
                   "that launch 1.39e9 (fp32 panels in, idx+dist out); the excess is the candidate lists (written by "
                   "the filter phase, read by the finish kernel) and the fp16 operand copies")


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm_gbs=d.get("hbm_gbs", 6650.0), bf16_tflops=d.get("bf16_tflops_sustained", 1400.0),
                    source="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML from a background thread (the library nvidia-smi
    itself is built on; one light call per 50 ms instead of a polling CLI that takes driver locks for milliseconds),
    `nvidia-smi -lms 200` (the profiling recipe's line) when the NVML binding is not importable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.stop_flag = False
        self.sm, self.reasons, self.sm_max = [], set(), None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nvml
        masks = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self.handle))
                for name, m in masks.items():
                    if r & m:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def pin_panel(panel):
    """Same panel with every array the library uploads living in pinned host memory."""
    import scipy.sparse as sp
    import torch
    from conos_b200 import Pagoda2

    def pin(a, dt):
        t = torch.from_numpy(np.ascontiguousarray(a, dt)).pin_memory()
        return t.numpy()
    out = {}
    for n, s in panel.items():
        m = sp.csc_matrix(s.counts)
        m.sort_indices()
        pm = sp.csc_matrix((pin(m.data, np.float64), pin(m.indices, np.int32), pin(m.indptr, np.int32)), shape=m.shape)
        pm.has_sorted_indices = True
        pca = np.asfortranarray(s.pca, np.float64)
        ppca = pin(pca.T, np.float64).T   # column-major view over pinned storage
        out[n] = Pagoda2(counts=pm, genes=s.genes, cells=s.cells, varinfo_v=pin(s.varinfo_v, np.float64),
                         varinfo_qv=pin(s.varinfo_qv, np.float64), pca=ppca, od_genes=s.od_genes)
    return out


def knn_recall(ctx, k, d):
    """recall@k of the library's kNN against an FP64 brute force (torch on the GPU, outside every timed region) on
    the kernel-level micro-input of SURVEY.md 8(d): 20-component Gaussian mixture, 2000 queries x 50000 references."""
    import torch
    from conos_b200 import ops
    rng = np.random.default_rng(SEED_RECALL)
    cent = rng.normal(0, 3.0, (20, d))
    A = cent[rng.integers(0, 20, 2000)] + rng.normal(0, 1, (2000, d))
    B = cent[rng.integers(0, 20, 50000)] + rng.normal(0, 1, (50000, d))
    idx, _ = ops.crossKnn(A, B, k, "angular", ctx=ctx)
    ta = torch.as_tensor(A, device="cuda", dtype=torch.float64)
    tb = torch.as_tensor(B, device="cuda", dtype=torch.float64)
    ta = ta / ta.norm(dim=1, keepdim=True)
    tb = tb / tb.norm(dim=1, keepdim=True)
    true = torch.topk(ta @ tb.T, k, dim=1).indices.cpu().numpy()
    hit = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(idx, true))
    return hit / float(idx.size)


SEED_RECALL = 20261019


def panel_bytes(panel):
    b = 0
    for s in panel.values():
        b += s.counts.indptr.size * 4 + s.counts.indices.size * 4 + s.counts.data.size * 8
        b += s.varinfo_v.size * 8 + s.varinfo_qv.size * 8 + s.pca.size * 8
    return b


def graph_bytes(g):
    return (g.vertices.nbytes + g.edge_from.nbytes + g.edge_to.nbytes + g.weight.nbytes + g.type.nbytes +
            g.adj_p.nbytes + g.adj_i.nbytes + g.adj_x.nbytes)


def cpu_reference_step(oracle, osamples, names, cfg, pair=(0, 1)):
    """The reference algorithm (oracle restatement) for ONE sample pair at full size: reduction, projection,
    exact kNN both ways, matching.  Returns seconds."""
    pr = {names[pair[0]]: osamples[names[pair[0]]], names[pair[1]]: osamples[names[pair[1]]]}
    t0 = time.perf_counter()
    oracle.build_graph(pr, k=cfg["k"], k_self=0, space=cfg["space"], ncomps=cfg["ncomps"], n_odgenes=cfg["n_genes"])
    return time.perf_counter() - t0


def to_oracle(oracle, panel):
    return {n: oracle.Sample(counts=s.counts, genes=s.genes, cells=s.cells, varinfo_v=s.varinfo_v,
                             varinfo_qv=s.varinfo_qv, pca=s.pca, od_genes=s.od_genes) for n, s in panel.items()}


_RESULT_FD = None


def guard_stdout():
    """Library chatter (NCCL prints its version on stdout) must not share the channel of the one JSON line: file
    descriptor 1 is pointed at stderr for the rest of the process and the result goes to a private duplicate."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 4])
    ap.add_argument("--n-samples", type=int, default=None, help="override (debug runs only)")
    ap.add_argument("--n-cells", type=int, default=None, help="override (debug runs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    reduced = False
    if args.n_samples:
        cfg["n_samples"], reduced = args.n_samples, True
    if args.n_cells:
        cfg["n_cells"], reduced = args.n_cells, True
    if reduced:
        cfg["workload"] += f" [REDUCED to {cfg['n_samples']} x {cfg['n_cells']}: not a valid bench line]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_pairs = cfg["n_samples"] * (cfg["n_samples"] - 1) // 2
    W = max(args.warmup, 0)
    K = max(args.steps, 1)

    # ------------------------------------------------------------------ reference arm: CPU, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return 0
        import conos_oracle as oracle
        oracle.build()
        from conos_b200.panel import make_panel
        cores = os.cpu_count() or 1
        # two samples of the workload are enough for a one-pair step
        panel = make_panel(args.config, 2, cfg["n_cells"], cfg["n_genes"], use_torch=None)
        osamples = to_oracle(oracle, panel)
        names = list(panel)
        times = []
        W = min(W, 1)  # a CPU run needs no clock warm-up; one untimed step pages the libraries in
        for it in range(W + K):
            dt = cpu_reference_step(oracle, osamples, names, cfg)
            if it >= W:
                times.append(dt)
            if sum(times) > 150 and len(times) >= 1 and it >= W:  # keep the whole run within a few minutes
                break
        steps_done = len(times)
        tot = sum(times)
        val = steps_done / tot
        sample = (f"1 of {n_pairs} sample pairs per step at full size (2 x {cfg['n_cells']} cells, {cfg['n_genes']} "
                  f"genes): reduction + projection + exact kNN both ways + matching; oracle port (exact float32 "
                  f"brute force, OpenMP), {steps_done} timed steps")
        line = {"impl": "reference", "metric": "buildGraph sample-pairs/s", "value": val, "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": steps_done, "warmup": W, "ms_per_step": 1e3 * tot / steps_done,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": cfg["workload"]},
                "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ this repository's arm
    import torch
    import conos_b200
    from conos_b200.panel import make_panel
    torch.cuda.set_device(local_rank)
    dist_group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = conos_b200.Context(local_rank)
    lib = ctx.lib
    t_gen = time.perf_counter()
    panel = make_panel(args.config, cfg["n_samples"], cfg["n_cells"], cfg["n_genes"])
    t_gen = time.perf_counter() - t_gen
    panel = pin_panel(panel)   # host copy of the inputs in page-locked memory (what e2e uploads every step)
    con = conos_b200.Conos(panel, ctx=ctx)
    wtuple = (rank, world, dist_group) if world > 1 else None
    kw = dict(k=cfg["k"], k_self=cfg["k_self"], space=cfg["space"], ncomps=cfg["ncomps"], n_odgenes=cfg["n_genes"])

    def barrier():
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()
        if os.environ.get("CONOS_BENCH_STAGES"):
            sys.stderr.write(f"[barrier rank {rank}] {1e3 * (time.perf_counter() - t1):.1f} ms\n")

    def device_step():
        con.pairs = {}                                   # no cached reductions (updatePairs recomputes every pair)
        if con._panel is not None:
            lib.cg_panel_drop_cache(con._panel)          # nor cached Gram matrices / PCA panels: inputs only
        return con.buildGraph(fetch=False, world=wtuple, **kw)

    def e2e_step():
        con.pairs = {}
        con._drop_panel()       # the panel is re-uploaded from host memory inside the timed region
        return con.buildGraph(fetch=True, world=wtuple, **kw)

    import ctypes as C
    import gc
    # ---- device-resident throughput
    for _ in range(W):
        device_step()
    # the panel holds ~10^6 Python strings (cell barcodes, gene names): a generation-2 collection of the cyclic GC walks
    # all of them and lands, at random, as a 100-600 ms pause inside a step.  Everything allocated so far is parked in
    # the permanent generation; the collector stays enabled for what the steps themselves allocate.
    gc.collect()
    gc.freeze()
    barrier()
    lib.cg_knn_timing(ctx.h, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    n_launch0 = ctx.kernel_launches()
    lib.cg_timer_start(ctx.h)
    t0 = time.perf_counter()
    step_ms = []
    for _ in range(K):
        t1 = time.perf_counter()
        g = device_step()
        step_ms.append(round(1e3 * (time.perf_counter() - t1), 1))
    if os.environ.get("CONOS_BENCH_STAGES"):
        sys.stderr.write(f"[steps rank {rank}] {step_ms}\n")
    ms = C.c_double(0)
    lib.cg_timer_stop(ctx.h, C.byref(ms))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.kernel_launches() - n_launch0
    kms, kl, kc, kb = C.c_double(0), C.c_uint64(0), C.c_double(0), C.c_double(0)
    lib.cg_knn_timing_read(ctx.h, C.byref(kms), C.byref(kl), C.byref(kc), C.byref(kb))
    lib.cg_knn_timing(ctx.h, 0)
    dev_ms = max(ms.value, 1e3 * wall)  # the host gaps between kernels belong to the step
    if world > 1:
        import torch.distributed as dist
        tmax = torch.tensor([dev_ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_ms = float(tmax.item())
        lsum = torch.tensor([float(launches)], device="cuda")
        dist.all_reduce(lsum)
        launches = int(lsum.item())
    if os.environ.get("CONOS_BENCH_STAGES"):   # debugging aid: one more step with the per-stage profile, every rank
        sys.stderr.write(f"[timed rank {rank}] last timed step py " +
                         json.dumps({k: round(v, 1) for k, v in con.timings.items()}) + "\n")
        lib.cg_profile(ctx.h, 1)
        t1 = time.perf_counter()
        device_step()
        dt = 1e3 * (time.perf_counter() - t1)
        buf = C.create_string_buffer(8192)
        lib.cg_profile_dump(ctx.h, buf, 8192)
        lib.cg_profile(ctx.h, 0)
        st = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in buf.value.decode().split(";") if kv}
        sys.stderr.write(f"[stages rank {rank}] wall {dt:.1f} ms " + json.dumps({k: round(v, 1) for k, v in st.items() if "." not in k}) +
                         " py " + json.dumps({k: round(v, 1) for k, v in con.timings.items()}) + "\n")
    if os.environ.get("CONOS_BENCH_TC_PROJECTION"):   # debugging aid: A/B of the projection engine in the e2e loop
        lib.cg_set_option(ctx.h, b"tc_projection", float(os.environ["CONOS_BENCH_TC_PROJECTION"]))
    n_inter = con._n_inter_edges
    if world > 1:
        import torch.distributed as dist
        ni = torch.tensor([float(n_inter)], device="cuda")
        dist.all_reduce(ni)
        n_inter = int(ni.item())
    # ---- end to end through the host API
    for _ in range(min(W, 2)):   # the first two passes through the host path still grow pinned / pooled buffers
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    K2 = max(1, min(K, 10))
    e2e_each = []
    for _ in range(K2):
        if os.environ.get("CONOS_BENCH_E2E_STAGES"):   # debugging aid (adds synchronisation: not a bench number)
            lib.cg_profile(ctx.h, 1)
        redo0 = (ctx.stat("knn_redo_blocks"), ctx.stat("eig_redo_pairs"))
        t1 = time.perf_counter()
        g_host = e2e_step()
        if os.environ.get("CONOS_BENCH_E2E_STAGES"):
            buf = C.create_string_buffer(8192)
            lib.cg_profile_dump(ctx.h, buf, 8192)
            lib.cg_profile(ctx.h, 0)
            sys.stderr.write(f"[e2e rank {rank}] {1e3 * (time.perf_counter() - t1):.1f} ms " + buf.value.decode() + "\n")
        e2e_each.append({"ms": round(1e3 * (time.perf_counter() - t1), 2),
                         "knn_redo_blocks": int(ctx.stat("knn_redo_blocks") - redo0[0]),
                         "eig_redo_pairs": int(ctx.stat("eig_redo_pairs") - redo0[1]),
                         "scratch_pool_mb": round(ctx.stat("scratch_pool_reserved_bytes") / 2**20),
                         "panel_pool_mb": round(ctx.stat("panel_pool_reserved_bytes") / 2**20),
                         **{k: round(v, 1) for k, v in con.timings.items() if k != "total_ms"}})
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        import torch.distributed as dist
        tmax = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    peaks = read_peaks()
    d = cfg["ncomps"]
    knn_s = kms.value * 1e-3
    flops = 2.0 * kc.value * d
    achieved = flops / knn_s / 1e12 if knn_s > 0 else 0.0
    roofline = {"bound": "tensor",
                "kernel": "knn_tc2_kernel + knn_tc2_finish_kernel (exact cross-kNN: fp16 tcgen05 tiles, group-max "
                          "thresholds, sign-bit filter, fp32 re-score + warp sort)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": KNN_DRAM_BYTES_PER_LAUNCH,
                "traffic_source": KNN_DRAM_SOURCE, "peak_source": peaks["source"],
                "launches": int(kl.value), "avg_launch_ms": kms.value / max(int(kl.value), 1),
                "candidates_per_s": kc.value / knn_s if knn_s > 0 else 0.0,
                "algorithmic_bytes_per_launch": kb.value / max(int(kl.value), 1),
                "share_of_step": knn_s / (dev_ms * 1e-3) if dev_ms > 0 else None,
                "note": "achieved = 2*n_query*n_ref*ncomps per launch / CUDA-event time on the launching stream "
                        "(both kernels and the overflow check); every distance tile is computed twice (threshold "
                        "phase, filter phase) with K padded 30 -> 32, so the tensor pipe issues 2.13x the algorithmic "
                        "flops; see DESIGN.md K2"}
    line = {"metric": "buildGraph sample-pairs/s", "value": n_pairs * K / (dev_ms * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "k": cfg["k"], "k_self": cfg["k_self"], "space": cfg["space"],
                       "ncomps": cfg["ncomps"], "pairs": n_pairs,
                       "l2": "inputs larger than L2 (panel and per-pair operands exceed 126 MB per step)",
                       "panel_generation_s": t_gen},
            "mnn_edges_per_s": n_inter * K / (dev_ms * 1e-3), "mnn_edges": n_inter,
            "graph": {"vertices": getattr(g, "n_vertices_device", None), "edges": getattr(g, "n_edges_device", None)},
            "e2e": {"value": n_pairs * K2 / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": panel_bytes(panel),
                    "d2h_bytes_per_step": graph_bytes(g_host), "steps": K2},
            "e2e_breakdown_ms": {k: round(v, 2) for k, v in con.timings.items()}, "e2e_steps_ms": e2e_each,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
    if rank == 0:
        line["knn_recall_at_k"] = knn_recall(ctx, cfg["k"], cfg["ncomps"])
    if not args.no_cpu_baseline:
        import conos_oracle as oracle
        oracle.build()
        names = list(panel)
        osamples = to_oracle(oracle, {n: panel[n] for n in names[:2]})
        dt = cpu_reference_step(oracle, osamples, names, cfg)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "pairs/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"1 of {n_pairs} pairs at full size (2 x {cfg['n_cells']} cells): reduction "
                                          f"+ projection + exact kNN both ways + matching, {dt:.1f} s"}
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
