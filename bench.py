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

HOST_THREADS = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
if "reference" in sys.argv or os.environ.get("RANK", "0") == "0":
    # torch.distributed.run exports OMP_NUM_THREADS=1 when nproc > 1; the CPU legs (reference arm, cpu_baseline) use
    # every host core the process may run on, whatever the launcher set -- before NumPy / OpenMP read the variable
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(HOST_THREADS)

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


# dram__bytes_read.sum + dram__bytes_write.sum of the cross-kNN launch of configs[1] on one GPU (56 problems of
# 50k x 50k = 1.4e11 scores), main kernel + finish kernel, from the ncu --set full capture summarised in profiles/.
# A launch on a rank of an N-GPU run scans fewer scores: `traffic` is this figure scaled by the scores of the launch.
KNN_DRAM_BYTES_CAPTURED = 0.352433e9 + 0.561278e9 + 1.113896e9 + 0.642316e9
KNN_DRAM_SCORES_CAPTURED = 56 * 50000.0 * 50000.0
KNN_DRAM_SOURCE = ("profiles/r02_ncu_knn_tc2_config2.txt: dram read+write of knn_tc2_kernel + "
                   "knn_tc2_finish_kernel for the cross-kNN launch of configs[1] on one GPU (1.4e11 scores), scaled by "
                   "the scores of this run's average launch; the excess over the algorithmic bytes is the candidate "
                   "lists (written by the filter phase, read by the finish kernel) and the fp16 operand copies")


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
        self.period = float(os.environ.get("CONOS_BENCH_CLOCK_PERIOD_S", "0.05"))
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
            time.sleep(self.period)

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
    exact kNN both ways, matching.  Returns (seconds, result of oracle.build_graph)."""
    pr = {names[pair[0]]: osamples[names[pair[0]]], names[pair[1]]: osamples[names[pair[1]]]}
    t0 = time.perf_counter()
    res = oracle.build_graph(pr, k=cfg["k"], k_self=0, space=cfg["space"], ncomps=cfg["ncomps"],
                             n_odgenes=cfg["n_genes"], alignment_strength=cfg.get("alignment_strength"))
    return time.perf_counter() - t0, res


def to_oracle(oracle, panel):
    return {n: oracle.Sample(counts=s.counts, genes=s.genes, cells=s.cells, varinfo_v=s.varinfo_v,
                             varinfo_qv=s.varinfo_qv, pca=s.pca, od_genes=s.od_genes) for n, s in panel.items()}


def oracle_threads(oracle):
    try:
        return int(oracle.lib().oracle_max_threads())
    except Exception:
        return HOST_THREADS


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


def edge_set_jaccard(ctx, lib, panel2, cfg, oracle_res):
    """Inter-sample edge set of ONE full-size pair, default mode of the library against the oracle (exact eigenvectors,
    FP64 projection, exact kNN): |intersection| / |union|."""
    import ctypes as C
    import conos_b200
    con = conos_b200.Conos(panel2, ctx=ctx)
    edges = con.buildGraph(k=cfg["k"], k_self=0, space=cfg["space"], ncomps=cfg["ncomps"], n_odgenes=cfg["n_genes"],
                           alignment_strength=cfg.get("alignment_strength"), assemble=False)
    n = int(lib.cg_edges_count(edges))
    f, t = np.empty(n, np.int32), np.empty(n, np.int32)
    ctx.check(lib.cg_edges_fetch(ctx.h, edges, f.ctypes.data_as(C.POINTER(C.c_int32)),
                                 t.ctypes.data_as(C.POINTER(C.c_int32)), None, None))
    lib.cg_edges_free(edges)
    con.close()
    of, ot, _, ty = oracle_res["edges"]
    nc = int(1 << 32)
    mine = np.unique(f.astype(np.int64) * nc + t.astype(np.int64))
    theirs = np.unique(of[ty == 1].astype(np.int64) * nc + ot[ty == 1].astype(np.int64))
    inter = np.intersect1d(mine, theirs, assume_unique=True).size
    return inter / max(mine.size + theirs.size - inter, 1), int(mine.size), int(theirs.size)


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="2 = BASELINE configs[1] (the bench line); 3 / 4 / 5 = configs[2] (CPCA), configs[3] (atlas), "
                         "configs[4] (CCA, alignment.strength): profiling runs")
    ap.add_argument("--n-samples", type=int, default=None, help="override (debug runs only)")
    ap.add_argument("--n-cells", type=int, default=None, help="override (debug runs only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end loop (profiling runs)")
    ap.add_argument("--no-check", action="store_true", help="skip the single-rank / oracle graph checks")
    ap.add_argument("--no-atlas", action="store_true",
                    help="skip the atlas sub-record (configs[3]: 100 samples x 10k cells, the scaling workload)")
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
        oracle.lib().oracle_set_threads(HOST_THREADS)   # explicit: never the launcher's OMP_NUM_THREADS
        threads = oracle_threads(oracle)
        # two samples of the workload are enough for a one-pair step
        panel = make_panel(args.config, 2, cfg["n_cells"], cfg["n_genes"], use_torch=None)
        osamples = to_oracle(oracle, panel)
        names = list(panel)
        times = []
        W = min(W, 1)  # a CPU run needs no clock warm-up; one untimed step pages the libraries in
        for it in range(W + K):
            dt, _ = cpu_reference_step(oracle, osamples, names, cfg)
            if it >= W:
                times.append(dt)
            if sum(times) > 150 and len(times) >= 1 and it >= W:  # keep the whole run within a few minutes
                break
        steps_done = len(times)
        tot = sum(times)
        val = steps_done / tot
        sample = (f"1 of {n_pairs} sample pairs per step at full size (2 x {cfg['n_cells']} cells, {cfg['n_genes']} "
                  f"genes): reduction + projection + exact kNN both ways + matching; oracle port of the reference's "
                  f"algorithm with its HNSW search replaced by an exact float32 brute force (O(n^2): the reference "
                  f"itself, ef = 50 k, would be faster at this size -- the ratio to this line is a speed-up over the "
                  f"exact-search baseline, not over conos's approximate search), OpenMP x {threads} threads, "
                  f"{steps_done} timed steps, k.self = 0")
        line = {"impl": "reference", "metric": "buildGraph sample-pairs/s", "value": val, "unit": "pairs/s",
                "n_gpus": args.gpus, "steps": steps_done, "warmup": W, "ms_per_step": 1e3 * tot / steps_done,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": {"workload": cfg["workload"]},
                "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample,
                                 "host_cores_available": HOST_THREADS,
                                 "launcher_omp_num_threads_ignored": True},
                "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ this repository's arm
    import torch
    import conos_b200
    from conos_b200.panel import make_panel
    torch.cuda.set_device(local_rank)
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
    wtuple = (rank, world, None) if world > 1 else None   # the library owns the communicator (cg_comm_init)
    kw = dict(k=cfg["k"], k_self=cfg["k_self"], space=cfg["space"], ncomps=cfg["ncomps"], n_odgenes=cfg["n_genes"],
              alignment_strength=cfg.get("alignment_strength"))

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
    lib.cg_knn_timing(ctx.h, 0 if os.environ.get("CONOS_BENCH_NO_KNN_TIMING") else 1)
    sampler = ClockSampler(local_rank)
    if rank == 0 and not os.environ.get("CONOS_BENCH_NO_SAMPLER"):
        sampler.start()
    n_launch0 = ctx.kernel_launches()
    lib.cg_timer_start(ctx.h)
    t0 = time.perf_counter()
    step_ms = []
    gc.disable()    # no collector pauses inside the timed steps (a 35 ms stall on one rank stalls all of them)
    trace = bool(os.environ.get("CONOS_BENCH_TRACE"))   # host-clock stage times of every step (no synchronisation added)
    traces = []
    for _ in range(K):
        if trace:
            lib.cg_profile(ctx.h, 2)
        t1 = time.perf_counter()
        g = device_step()
        step_ms.append(round(1e3 * (time.perf_counter() - t1), 1))
        if trace:
            buf = C.create_string_buffer(8192)
            lib.cg_profile_dump(ctx.h, buf, 8192)
            py_ = {k: round(v, 1) for k, v in con.timings.items()}
            py_["scratch_pool_mb"] = round(ctx.stat("scratch_pool_reserved_bytes") / 2**20)
            py_["panel_pool_mb"] = round(ctx.stat("panel_pool_reserved_bytes") / 2**20)
            traces.append((step_ms[-1], buf.value.decode(), py_))
    if trace:
        lib.cg_profile(ctx.h, 0)
        med = sorted(step_ms)[len(step_ms) // 2]
        pools = [(t_[2]["scratch_pool_mb"], t_[2]["panel_pool_mb"]) for t_ in traces]
        sys.stderr.write(f"[trace rank {rank}] pool MB per step (scratch, panel): {pools}\n")
        for ms_, st_, py_ in traces:
            if ms_ > med + 4 or ms_ == med:
                sys.stderr.write(f"[trace rank {rank}] {ms_} ms (median {med}) {st_} py {json.dumps(py_)}\n")
    gc.enable()
    sys.stderr.write(f"[steps rank {rank}] {step_ms}\n")
    ms = C.c_double(0)
    lib.cg_timer_stop(ctx.h, C.byref(ms))
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 and not os.environ.get("CONOS_BENCH_NO_SAMPLER") else None
    launches = ctx.kernel_launches() - n_launch0
    knn_scores = ctx.stat("knn_scores")
    kms, kl, kc, kb = C.c_double(0), C.c_uint64(0), C.c_double(0), C.c_double(0)
    lib.cg_knn_timing_read(ctx.h, C.byref(kms), C.byref(kl), C.byref(kc), C.byref(kb))
    lib.cg_knn_timing(ctx.h, 0)
    dev_ms = max(ms.value, 1e3 * wall)  # the host gaps between kernels belong to the step
    graph_sizes = (getattr(g, "n_vertices_device", None), getattr(g, "n_edges_device", None)) if g is not None \
        else (None, None)
    n_inter = con._n_inter_edges                         # after the allgatherv: the whole graph's, on every rank
    rank_stats = None
    if world > 1:
        import torch.distributed as dist
        tmax = torch.tensor([dev_ms], device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dev_ms = float(tmax.item())
        lsum = torch.tensor([float(launches)], device="cuda")
        dist.all_reduce(lsum)
        launches = int(lsum.item())
        # per-rank picture of the last timed step (pairs, samples uploaded, kNN time): what scaling is made of
        mine = torch.tensor([float(len(con._last_pair_counts)), float(len([n for n in con._panel_key[2]])),
                             kms.value / K, float(step_ms[-1])], device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_stats = [{"pairs": int(t[0]), "samples_resident": int(t[1]), "knn_ms_per_step": round(float(t[2]), 2),
                       "last_step_ms": round(float(t[3]), 1)} for t in allr]
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
    # ---- end to end through the host API
    e2e_each, e2e_s, K2, g_host, h2d = [], None, 0, None, 0
    if not args.no_e2e:
        for _ in range(min(W, 2)):   # the first two passes through the host path still grow pinned / pooled buffers
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        K2 = max(1, min(K, 10))
        for _ in range(K2):
            if os.environ.get("CONOS_BENCH_E2E_STAGES"):   # debugging aid (adds synchronisation: not a bench number)
                lib.cg_profile(ctx.h, 1)
            redo0 = (ctx.stat("knn_redo_blocks"), ctx.stat("eig_redo_pairs"))
            t1 = time.perf_counter()
            gh_ = e2e_step()
            if gh_ is not None:
                g_host = gh_
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
        h2d = con._h2d_bytes
        if world > 1:
            import torch.distributed as dist
            tmax = torch.tensor([e2e_s], device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            e2e_s = float(tmax.item())
            hsum = torch.tensor([float(h2d)], device="cuda")
            dist.all_reduce(hsum)
            h2d = int(hsum.item())
    # ---- atlas sub-record: BASELINE configs[3], the workload north_star names for 1 -> 8 GPU scaling
    atlas = None
    if args.config == 2 and not args.no_atlas and not reduced:
        acfg = CONFIGS[4]
        con._drop_panel()
        t_a = time.perf_counter()
        apanel = make_panel(4, acfg["n_samples"], acfg["n_cells"], acfg["n_genes"])
        t_a = time.perf_counter() - t_a
        acon = conos_b200.Conos(apanel, ctx=ctx)
        akw = dict(k=acfg["k"], k_self=acfg["k_self"], space=acfg["space"], ncomps=acfg["ncomps"],
                   n_odgenes=acfg["n_genes"])

        def atlas_step():
            acon.pairs = {}
            if acon._panel is not None:
                lib.cg_panel_drop_cache(acon._panel)
            return acon.buildGraph(fetch=False, world=wtuple, **akw)
        atlas_step()
        barrier()
        gc.disable()
        t0 = time.perf_counter()
        KA = 2
        for _ in range(KA):
            ga = atlas_step()
        barrier()
        a_s = time.perf_counter() - t0
        gc.enable()
        if world > 1:
            import torch.distributed as dist
            tmax = torch.tensor([a_s], device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            a_s = float(tmax.item())
        n_ap = acfg["n_samples"] * (acfg["n_samples"] - 1) // 2
        atlas = {"workload": acfg["workload"], "value": n_ap * KA / a_s, "unit": "pairs/s", "steps": KA, "warmup": 1,
                 "ms_per_step": 1e3 * a_s / KA, "n_gpus": world, "pairs": n_ap,
                 "graph_edges": getattr(ga, "n_edges_device", None) if ga is not None else None,
                 "samples_resident_rank0": None if acon._panel_key[2] is None else len(acon._panel_key[2]),
                 "panel_generation_s": t_a,
                 "note": "device-resident, one buildGraph per step, wall clock between barriers, max over ranks"}
        acon.close()
        del apanel
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        con.close()
        ctx.close()
        return 0

    peaks = read_peaks()
    d = cfg["ncomps"]
    knn_s = kms.value * 1e-3
    # 2 * n_a * n_b * d flop per sample pair -- ONE distance matrix serves both directions -- plus
    # 2 * n^2 * d for a sample against itself; kc already counts a mirrored pair of problems once
    flops = 2.0 * kc.value * d
    achieved = flops / knn_s / 1e12 if knn_s > 0 else 0.0
    n_l = max(int(kl.value), 1)
    roofline = {"bound": "tensor",
                "kernel": "knn_tc2_kernel + knn_tc2_finish_kernel (exact kNN, cross and self: fp16 tcgen05 tiles, "
                          "group-max thresholds, sign-bit filter, fp32 re-score + warp sort)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"],
                "traffic": KNN_DRAM_BYTES_CAPTURED * (knn_scores / n_l) / KNN_DRAM_SCORES_CAPTURED,
                "traffic_source": KNN_DRAM_SOURCE, "peak_source": peaks["source"],
                "launches": int(kl.value), "avg_launch_ms": kms.value / n_l,
                "algorithmic_flop_per_launch": flops / n_l,
                "scores_scanned_per_s": knn_scores / knn_s if knn_s > 0 else 0.0,
                "algorithmic_bytes_per_launch": kb.value / n_l,
                "share_of_step": knn_s / (dev_ms * 1e-3) if dev_ms > 0 else None,
                "tensor_pipe_active_pct_ncu": 28.3,
                "note": "achieved = SURVEY 8(d) flops (2*n_a*n_b*ncomps per sample pair, one distance matrix for both "
                        "directions; 2*n^2*ncomps for the self search) / CUDA-event time of the kNN launches on the "
                        "launching stream (both kernels and the overflow check); the kernel computes each direction "
                        "separately, twice (threshold phase, filter phase), with K padded 30 -> 32, so the tensor pipe "
                        "issues 4.27x these flops; tensor_pipe_active_pct_ncu is sm__pipe_tensor_subunit busy of the "
                        "committed ncu capture (profiles/r02_ncu_knn_tc2_config2.txt), see DESIGN.md K2"}
    line = {"metric": "buildGraph sample-pairs/s", "value": n_pairs * K / (dev_ms * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": dev_ms / K, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"], "k": cfg["k"], "k_self": cfg["k_self"], "space": cfg["space"],
                       "ncomps": cfg["ncomps"], "pairs": n_pairs,
                       "l2": "inputs larger than L2 (panel and per-pair operands exceed 126 MB per step)",
                       "panel_generation_s": t_gen},
            "steps_ms": step_ms,
            "mnn_edges_per_s": n_inter * K / (dev_ms * 1e-3), "mnn_edges": n_inter,
            "graph": {"vertices": graph_sizes[0], "edges": graph_sizes[1]},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
    if rank_stats is not None:
        line["ranks"] = rank_stats
    if atlas is not None:
        line["atlas"] = atlas
    if e2e_s is not None:
        each = sorted(x["ms"] for x in e2e_each)
        line["e2e"] = {"value": n_pairs * K2 / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": graph_bytes(g_host) if g_host is not None else 0, "steps": K2,
                       "ms_per_step_median": each[len(each) // 2], "ms_per_step_max": each[-1],
                       "ms_per_step_min": each[0]}
        line["e2e_breakdown_ms"] = {k: round(v, 2) for k, v in con.timings.items()}
        line["e2e_steps_ms"] = e2e_each
    line["knn_recall_at_k"] = knn_recall(ctx, min(cfg["k"], 30), cfg["ncomps"])
    # ---- diffusion sub-record: propagate_labels over the graph this run built, HBM-bound
    if g_host is not None and world == 1 and not args.no_check and g_host.ecount() > 0:
        from conos_b200 import ops
        nv_, ne_, n_lab = g_host.vcount(), g_host.ecount(), 20
        rng_ = np.random.default_rng(1)
        lab_ = np.where(rng_.random(nv_) < 0.1, rng_.integers(0, n_lab, nv_), -1).astype(np.int32)

        def run_diffusion(iters):
            torch.cuda.synchronize()
            t_ = time.perf_counter()
            ops.propagate_labels(g_host.edge_from, g_host.edge_to, g_host.weight, nv_, lab_, n_lab, max_n_iters=iters,
                                 tol=0.0, fixed_initial_labels=True, ctx=ctx)
            return time.perf_counter() - t_
        run_diffusion(1)
        t_2 = min(run_diffusion(2) for _ in range(3))
        t_22 = min(run_diffusion(22) for _ in range(3))
        per_iter = (t_22 - t_2) / 20.0
        # algorithmic HBM bytes of one iteration: the incidence lists (int32 neighbour + f64 weight per entry, two
        # entries per edge) + the matrix read once and written once (the gathers of neighbour rows hit L2: 64 MB)
        by = 2.0 * ne_ * 12 + 2.0 * nv_ * n_lab * 8
        line["diffusion"] = {"workload": f"propagate_labels on this graph ({nv_} vertices, {ne_} edges), {n_lab} labels, "
                                         f"10 % of the vertices labelled and fixed",
                             "ms_per_iteration": 1e3 * per_iter, "algorithmic_bytes_per_iteration": by,
                             "achieved_gbs": by / per_iter / 1e9, "peak_gbs": peaks["hbm_gbs"],
                             "frac": by / per_iter / 1e9 / peaks["hbm_gbs"],
                             "gathered_bytes_per_iteration": 2.0 * ne_ * n_lab * 8,
                             "ms_call_22_iterations_host_edge_list": 1e3 * t_22,
                             "note": "per iteration = (22-iteration call - 2-iteration call) / 20, wall clock around "
                                     "the C-ABI call with the edge list in host memory"}
        if not args.no_cpu_baseline:
            import conos_oracle as oracle_d
            oracle_d.build()
            t_ = time.perf_counter()
            oracle_d.propagate_labels(g_host.edge_from, g_host.edge_to, g_host.weight, nv_, lab_, n_lab, max_n_iters=2,
                                      tol=0.0, fixed_initial_labels=True)
            line["diffusion"]["cpu_oracle_ms_per_iteration"] = 1e3 * (time.perf_counter() - t_) / 2.0
    # ---- graph checks (outside every timed region)
    if not args.no_check and g_host is not None and world > 1:
        # the N-rank graph against the graph one rank builds alone: every array, bit for bit
        solo = conos_b200.Conos(panel, ctx=ctx)
        g1 = solo.buildGraph(fetch=True, **kw)
        same = all(np.array_equal(getattr(g_host, a), getattr(g1, a))
                   for a in ("vertices", "edge_from", "edge_to", "weight", "type"))
        line["graph_equals_single_rank"] = bool(same)
        line["graph_single_rank_edges"] = int(g1.ecount())
        solo.close()
    if not args.no_cpu_baseline and world == 1:
        import conos_oracle as oracle
        oracle.build()
        oracle.lib().oracle_set_threads(HOST_THREADS)
        names = list(panel)
        panel2 = {n: panel[n] for n in names[:2]}
        osamples = to_oracle(oracle, panel2)
        dt, ores = cpu_reference_step(oracle, osamples, names, cfg)
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "pairs/s", "cores": oracle_threads(oracle), "kind": "port",
                                "sample": f"1 of {n_pairs} pairs at full size (2 x {cfg['n_cells']} cells): reduction "
                                          f"+ projection + exact float32 brute-force kNN both ways + matching (the "
                                          f"reference's HNSW replaced by exact search), {dt:.1f} s"}
        if not args.no_check:
            jac, n_mine, n_theirs = edge_set_jaccard(ctx, lib, panel2, cfg, ores)
            line["graph_jaccard_vs_oracle"] = {"value": jac, "pair": "samples 0 and 1 at full size, default mode",
                                               "edges": n_mine, "oracle_edges": n_theirs}
    emit(line)
    con.close()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
