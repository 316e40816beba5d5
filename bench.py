#!/usr/bin/env python
"""bench.py — the headline benchmark: PRM roadmap build on a 7-D joint box, 1M samples,
k = 10, capsule-obstacle world (BASELINE.json config 4), at 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one complete roadmap build over the resident store: the k-NN phase (every node
against all nodes) and the motion check of every candidate edge.  At N > 1 the query rows and
the candidate edges are sharded over the ranks against a replicated store (strong scaling:
the 1M-node job is fixed) and the results are exchanged with two NCCL all-gathers.

One JSON line is printed by rank 0 (see the task contract): `value` is nodes/s with the
store resident in HBM, `e2e` is the same build through the host-buffer C ABI with the H2D
upload of the samples and the D2H read of the roadmap inside the timed region, `roofline`
is the k-NN kernel's algorithmic FP64 rate against the unfused-FP64 issue peak measured in
the same run, `cpu_baseline` is the CPU oracle (single thread, as the reference is) on a
bounded sample.  `--impl reference` times the CPU oracle with all host threads instead.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "PRM edges checked/sec + k-NN queries/sec (7-D, 1M nodes) at 1/2/4/8 B200"
UNIT = "nodes/s (one k-NN query + the motion checks of its candidate edges per node)"
DIM = 7
SEED = 4
# dram__bytes_read.sum + dram__bytes_write.sum of one k-NN launch at this config, from the
# `ncu --set full` capture summarised in profiles/ (None until captured for the current kernel)
KNN_DRAM_BYTES_NCU = 2249420000  # profiles/r01_ncu_full_knn_rv_dot_7_2.csv (2.046 GB read + 0.204 GB written)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--capsules", type=int, default=256)
    ap.add_argument("--cpu-queries", type=int, default=2400,
                    help="k-NN queries in the single-thread cpu_baseline sample")
    ap.add_argument("--cpu-edges", type=int, default=24000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--star-nodes", type=int, default=60000,
                    help="tree size of the RRT* (config 3) extra; the full config is 200000")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the single-query HBM scan and the fused-planner figures")
    return ap.parse_args()


def config_dict(args, world_size):
    return {
        "workload": "config 4: PRM roadmap build, 7-D Panda-like joint box, %d valid uniform "
                    "samples (seed %d), k=%d, %d capsules (worlds.capsule_world seed %d, len<=3, "
                    "r in [1.0,1.6])" % (args.nodes, SEED, args.k, args.capsules, SEED),
        "nodes": args.nodes,
        "k": args.k,
        "dim": DIM,
        "capsules": args.capsules,
        "sharding": "rows and candidate edges over %d rank(s), replicated store, 2 all-gathers"
                    % world_size,
        "l2": "store (56 MB) is L2-resident by design; a 256 MiB memset flushes L2 between "
              "steps inside the timed region (<0.1% of a step)",
    }


# ------------------------------------------------------------------------------------------
# workload generation (host; the validity filter that discards invalid samples runs on the GPU
# for the product arm and in the oracle for the reference arm — outside every timed region)
# ------------------------------------------------------------------------------------------
def raw_samples(n, seed):
    from oxmpl_b200 import worlds

    rng = np.random.default_rng(seed)
    return worlds.PANDA_LO + (worlds.PANDA_HI - worlds.PANDA_LO) * rng.random((n, DIM))


def resolution():
    from oxmpl_b200 import worlds
    from oxmpl_b200.api import motion_resolution, real_vector_max_extent

    return motion_resolution(real_vector_max_extent(list(zip(worlds.PANDA_LO, worlds.PANDA_HI))))


def valid_samples(n, is_valid_fn):
    """First n valid samples of the seed-4 uniform stream."""
    out = []
    have = 0
    chunk_seed = SEED
    while have < n:
        s = raw_samples(max(n // 2, 4096), chunk_seed)
        ok = is_valid_fn(s).astype(bool)
        out.append(s[ok])
        have += int(ok.sum())
        chunk_seed += 1000
    return np.ascontiguousarray(np.vstack(out)[:n])


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu_index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (the Rust reference cannot be built here), all host threads
# ------------------------------------------------------------------------------------------
def cpu_sample(orc, ow, sp_o, rows, args, res, n_queries, n_edges, threads):
    """Times the oracle on a bounded sample of the same workload.  Returns a dict with the
    per-unit times and the extrapolated whole-build nodes/s."""
    n = rows.shape[0]
    k = args.k
    os.environ["OMP_NUM_THREADS"] = str(threads)
    try:
        lib_omp = C.CDLL("libgomp.so.1")
        lib_omp.omp_set_num_threads(int(threads))
    except OSError:
        pass
    rng = np.random.default_rng(123)
    q_rows = np.sort(rng.choice(n, size=min(n_queries, n), replace=False)).astype(np.uint32)
    t0 = time.perf_counter()
    idx, dist = orc.knn(sp_o, rows, rows[q_rows], k, exclude=q_rows)
    t_knn = time.perf_counter() - t0
    # candidate edges of the sampled rows, checked the way the build checks them
    i = np.repeat(q_rows, k)
    j = idx.reshape(-1)
    ok = j != 0xFFFFFFFF
    i, j = i[ok][:n_edges], j[ok][:n_edges]
    frm = rows[np.maximum(i, j)]
    to = rows[np.minimum(i, j)]
    t0 = time.perf_counter()
    if threads > 1:
        # split the edge sample over the threads with the batch entry point per chunk
        chunks = np.array_split(np.arange(len(i)), threads)
        res_list = [None] * threads

        def work(c):
            res_list[c] = orc.check_motion(ow, sp_o, frm[chunks[c]], to[chunks[c]], res)[0]

        ts = [threading.Thread(target=work, args=(c,)) for c in range(threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    else:
        orc.check_motion(ow, sp_o, frm, to, res)
    t_edge = time.perf_counter() - t0
    s_per_query = t_knn / len(q_rows)
    s_per_edge = t_edge / max(len(i), 1)
    return {"s_per_query": s_per_query, "s_per_edge": s_per_edge, "n_queries": int(len(q_rows)),
            "n_edges": int(len(i)), "t_knn": t_knn, "t_edge": t_edge}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle
    from oxmpl_b200 import worlds

    orc = oracle.get()
    threads = os.cpu_count() or 1
    caps = worlds.capsule_world(SEED, worlds.PANDA_LO, worlds.PANDA_HI, args.capsules)
    sp_o = orc.space(DIM)
    ow, keep = orc.world(DIM, capsules=caps)
    res = resolution()
    rows = valid_samples(args.nodes, lambda s: orc.is_valid(ow, sp_o, s))
    # edges per node of the full build ~ k minus half of the mutual pairs; measured on the sample
    n_q = max(threads * 24, 64)
    n_e = max(threads * 1500, 4000)
    times = []
    per_step = None
    for it in range(args.warmup + args.steps):
        per_step = cpu_sample(orc, ow, sp_o, rows, args, res, n_q, n_e, threads)
        if it >= args.warmup:
            times.append(per_step)
    s_q = float(np.mean([t["s_per_query"] for t in times]))
    s_e = float(np.mean([t["s_per_edge"] for t in times]))
    edges_per_node = 0.75 * args.k  # unique candidate edges per node (mutual pairs counted once)
    s_node = s_q + edges_per_node * s_e
    value = 1.0 / s_node
    ms_per_step = 1e3 * float(np.mean([t["t_knn"] + t["t_edge"] for t in times]))
    sample = ("per step: %d k-NN queries x %d nodes + %d motion checks, OpenMP/threads=%d; nodes/s "
              "extrapolated as 1/(s_per_query + %.2f*s_per_edge)" % (n_q, args.nodes, n_e, threads,
                                                                     edges_per_node))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "knn_queries_per_s": 1.0 / s_q, "edges_checked_per_s": 1.0 / s_e,
        "note": "CPU oracle (C++ restatement of the Rust reference, which cannot be built in this "
                "image: no cargo/rustc); extrapolated from a bounded sample",
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------
# product arm
# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import oxmpl_b200 as oxg
    from oxmpl_b200 import worlds
    from oxmpl_b200._lib import PrmStats, check, lib

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world_size != args.gpus and world_size > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world_size}")
    if not torch.cuda.is_available() or oxg.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: oxmpl_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    check(lib.oxg_set_device(local_rank))
    dev = torch.device("cuda", local_rank)
    if world_size > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n, k = args.nodes, args.k
    sp = oxg.Space.real_vector(DIM)
    caps = worlds.capsule_world(SEED, worlds.PANDA_LO, worlds.PANDA_HI, args.capsules)
    res = resolution()
    gw = oxg.World(DIM, capsules=caps)
    rows_np = valid_samples(n, lambda s: gw.is_valid(sp, s))
    rows_pinned = torch.from_numpy(rows_np).pin_memory()

    st = oxg.StateStore(sp, capacity=n)
    st.append_ptr(rows_pinned.data_ptr(), n)
    ext = torch.cuda.ExternalStream(st.stream(), device=dev)

    # shard: contiguous query rows (uniform work per query in k-NN-all mode)
    from oxmpl_b200 import sharding

    rb, re = sharding.row_range(n, rank, world_size)
    rows_max = sharding.rows_max(n, world_size)
    even = n % world_size == 0

    with torch.cuda.stream(ext):
        d_idx_all = torch.empty((n, k), dtype=torch.int32, device=dev)       # u32 bit patterns
        d_dist_all = torch.empty((n, k), dtype=torch.float64, device=dev)
        d_valid_all = torch.empty((n, k), dtype=torch.uint8, device=dev)
        d_counters = torch.zeros(2, dtype=torch.int64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        if not even:
            pad_idx = torch.empty((world_size, rows_max, k), dtype=torch.int32, device=dev)
            pad_dist = torch.empty((world_size, rows_max, k), dtype=torch.float64, device=dev)
            pad_valid = torch.empty((world_size, rows_max, k), dtype=torch.uint8, device=dev)
    ext.synchronize()

    def gather(all_t, pad_t):
        """all-gather of the row shards, in place on the [n, k] tensor."""
        sharding.all_gather_rows(dist, all_t, rank, world_size, pad_t)

    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def step(timed_events=None):
        with torch.cuda.stream(ext):
            flush.zero_()
            if timed_events:
                timed_events[0].record(ext)
            check(lib.oxg_knn_rows_dev(st.handle, rb, re, k,
                                       C.c_void_p(d_idx_all[rb:re].data_ptr()),
                                       C.c_void_p(d_dist_all[rb:re].data_ptr())))
            if timed_events:
                timed_events[1].record(ext)
            gather(d_idx_all, None if even else pad_idx)
            gather(d_dist_all, None if even else pad_dist)
            check(lib.oxg_prm_edges_dev(gw.handle, st.handle, C.c_void_p(d_idx_all.data_ptr()), k,
                                        float(res), rb, re, C.c_void_p(d_valid_all.data_ptr()),
                                        C.c_void_p(d_counters.data_ptr())))
            gather(d_valid_all, None if even else pad_valid)
            check(lib.oxg_prm_resolve_dev(st.handle, C.c_void_p(d_idx_all.data_ptr()), k,
                                          C.c_void_p(d_valid_all.data_ptr())))
            if timed_events:
                timed_events[2].record(ext)

    def barrier():
        ext.synchronize()
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-store timing -------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    oxg.kernel_launches(reset=True)
    knn_ms, total_ms = [], []
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(ext)
    step_events = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for s_i in range(args.steps):
        step(step_events[s_i])
    stop.record(ext)
    barrier()
    launches = oxg.kernel_launches()
    elapsed_ms = start.elapsed_time(stop)
    for e in step_events:
        knn_ms.append(e[0].elapsed_time(e[1]))
        total_ms.append(e[0].elapsed_time(e[2]))
    counters = d_counters.cpu().numpy()
    n_checked_local, n_valid_local = int(counters[0]), int(counters[1])

    # ---- end-to-end through the host-buffer C ABI ---------------------------------------------
    rows_local = re - rb
    h_idx = torch.empty((rows_local, k), dtype=torch.int32).pin_memory()
    h_dist = torch.empty((rows_local, k), dtype=torch.float64).pin_memory()
    h_valid = torch.empty((rows_local, k), dtype=torch.uint8).pin_memory()
    stats = PrmStats()

    def e2e_step():
        st.clear()
        st.append_ptr(rows_pinned.data_ptr(), n)                       # H2D: the samples
        check(lib.oxg_prm_build_knn(gw.handle, st.handle, k, float(res), rb, re,
                                    C.c_void_p(h_idx.data_ptr()), C.c_void_p(h_dist.data_ptr()),
                                    C.c_void_p(h_valid.data_ptr()), C.byref(stats)))  # D2H: roadmap

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- max over ranks ---------------------------------------------------------------------
    vals = torch.tensor([elapsed_ms, e2e_s, float(np.mean(knn_ms)), float(np.mean(total_ms)),
                         float(n_checked_local)], dtype=torch.float64, device=dev)
    if world_size > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = vals, vals
    elapsed_ms_max, e2e_s_max, knn_ms_max, total_ms_max = (float(mx[0]), float(mx[1]), float(mx[2]),
                                                           float(mx[3]))
    edges_checked_total = float(sm[4])

    if rank == 0:
        ms_per_step = elapsed_ms_max / args.steps
        value = n / (ms_per_step * 1e-3)
        peak64_ops, _ = oxg.fp64_peak()
        peak32_ops, _ = oxg.fp32_peak()
        # algorithmic work of the dominant kernel: 3*D flop per (query, stored) pair
        rows_rank0 = re - rb
        flops_per_launch = 3.0 * DIM * float(rows_rank0) * float(n)
        achieved = flops_per_launch / (knn_ms_max * 1e-3)
        filtered = os.environ.get("OXG_KNN_FILTER", "1") != "0"
        if filtered:
            # The kernel evaluates every pair with the certified FP32 pre-filter (dot form: D
            # packed FP32 FMAs per pair, FFMA2) and recomputes the ~1e-4 survivors in FP64 in
            # the reference's order; the FP32 FMA pipe bounds it.  achieved = the algorithmic
            # 3*D flop per pair of SURVEY §8(d) / time; peak = the measured FP32 FMA rate
            # (lane-instruction rate x 2 flop).
            peak = peak32_ops * 2.0
            executed = 1.0 * DIM * float(rows_rank0) * float(n)   # FMA lane-instructions issued
            roof = {
                "kernel": "oxg::knn_rv_dot_kernel<7,2>", "bound": "fp32",
                "achieved": achieved / 1e12, "peak": peak / 1e12, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": KNN_DRAM_BYTES_NCU,
                "algorithmic_flops_per_launch": flops_per_launch, "launch_ms": knn_ms_max,
                "fma_pipe_utilisation": executed / (knn_ms_max * 1e-3) / peak32_ops,
                "peak_source": "oxg_fp32_peak(): packed FP32 (FADD2/FFMA2) lane-instruction rate "
                               "measured in this run x 2 flop per FMA (MEASURED_PEAKS.json holds HBM "
                               "and bf16-tensor figures only; the path is neither: the store is "
                               "L2-resident and separately-rounded f64 sums have no tensor-core form)",
                "vs_unfused_fp64_peak": {
                    "peak_tflops": peak64_ops / 1e12, "frac": achieved / peak64_ops,
                    "note": "an all-FP64 evaluation (3*D DADD/DMUL per pair, no FMA) cannot exceed "
                            "1.0 here; the certified FP32 pre-filter is how the kernel does"},
                "hbm_peak_gbs_measured": _measured_hbm(),
            }
        else:
            roof = {
                "kernel": "oxg::knn_rv_kernel<7,2>", "bound": "fp64",
                "achieved": achieved / 1e12, "peak": peak64_ops / 1e12, "unit": "TFLOP/s",
                "frac": achieved / peak64_ops, "traffic": 313393408,
                "algorithmic_flops_per_launch": flops_per_launch, "launch_ms": knn_ms_max,
                "peak_source": "oxg_fp64_peak(): unfused DADD/DMUL issue rate measured in this run",
                "hbm_peak_gbs_measured": _measured_hbm(),
            }
        edge_ms = total_ms_max - knn_ms_max
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args, world_size),
            "knn_queries_per_s": n / (knn_ms_max * 1e-3),
            "edges_checked_per_s": edges_checked_total / (max(edge_ms, 1e-6) * 1e-3),
            "edges_checked_per_step": edges_checked_total,
            "phase_ms": {"knn": knn_ms_max, "gather_edges_resolve": edge_ms},
            "e2e": {"value": n / e2e_s_max, "unit": UNIT,
                    "h2d_bytes_per_step": int(n * DIM * 8),
                    "d2h_bytes_per_step": int(rows_local * k * (4 + 8 + 1)),
                    "ms_per_step": e2e_s_max * 1e3,
                    "path": "oxg_store_clear + oxg_store_append(pinned host rows) + "
                            "oxg_prm_build_knn(host outputs); wall clock around synchronous calls"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        if world_size == 1 and not args.no_extras:
            try:
                line["extras"] = extras(oxg, args)
            except Exception as e:  # secondary figures never break the headline line
                line["extras"] = {"error": repr(e)}
        if not args.no_cpu_baseline and world_size == 1:
            try:
                line["cpu_baseline"] = cpu_baseline(args, rows_np, caps, res)
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 1, "kind": "port",
                                        "sample": "failed: " + repr(e)}
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    st.close()
    gw.close()
    return 0


def extras(oxg, args):
    """Secondary figures (N=1 only, a few seconds): the HBM-bound single-query scan and the
    fused sequential planners.  The headline stays the PRM build above."""
    import torch

    out = {}
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.zero_()
        torch.cuda.synchronize()

    # -- single-query 1-NN scan over a store larger than L2: HBM-bound -------------
    n_big, d_big = 10_000_000, 14
    sp = oxg.Space.real_vector(d_big)
    rng = np.random.default_rng(5)
    with oxg.StateStore(sp, capacity=n_big) as st:
        chunk = 1_000_000
        for _ in range(n_big // chunk):
            st.append(rng.random((chunk, d_big)))
        q = rng.random((1, d_big))
        ts, dev = [], []
        for it in range(8):
            flush_l2()
            t0 = time.perf_counter()
            st.nn1(q)
            ts.append(time.perf_counter() - t0)
            dev.append(st.last_scan_ms())
        best, best_dev = min(ts[2:]), min(dev[2:])
        bytes_alg = 8 * d_big * n_big  # every f64 coordinate of the store, read once
        out["nn1_single_query_scan"] = {
            "kernel": "oxg::nn1_stream_kernel<14,1>", "n": n_big, "dim": d_big, "bound": "hbm",
            "algorithmic_bytes": bytes_alg, "device_ms": best_dev, "host_call_ms": best * 1e3,
            "achieved_gbs": bytes_alg / (best_dev * 1e-3) / 1e9,
            "peak_gbs": _measured_hbm(),
            "frac": bytes_alg / (best_dev * 1e-3) / 1e9 / (_measured_hbm() or 6650.0),
            "note": "store (1.12 GB) exceeds the 126 MB L2 and L2 is flushed before every call; "
                    "device_ms = CUDA events around the scan + merge kernels, host_call_ms adds the "
                    "query upload, result download and launch latency of the synchronous C-ABI call",
        }
    torch.cuda.empty_cache()
    # -- fused sequential planners (BASELINE configs 2 and 3): GPU only here; the CPU oracle on the
    #    same pre-drawn streams (trees compared bit for bit) is timed in the cpu_baseline leg -----
    try:
        planner_probe = _planner_probe()
        out["rrt_connect_config2"] = planner_probe.config2(30000, cpu=False)
        out["rrt_star_config3"] = planner_probe.config3(args.star_nodes, 10000, 0)
    except Exception as e:  # extras never break the headline line
        out["planners_error"] = repr(e)
    return out


def _planner_probe():
    sys.path.insert(0, os.path.join(ROOT, "tests", "probes"))
    import planner_probe

    return planner_probe


def _measured_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get("hbm_gbs")
    except Exception:
        return None


def cpu_baseline(args, rows, caps, res):
    """The CPU oracle, one thread (the reference is single-threaded), bounded sample."""
    from oracle import oracle

    orc = oracle.get()
    sp_o = orc.space(DIM)
    ow, keep = orc.world(DIM, capsules=caps)
    t = cpu_sample(orc, ow, sp_o, rows, args, res, args.cpu_queries, args.cpu_edges, 1)
    edges_per_node = 0.75 * args.k
    s_node = t["s_per_query"] + edges_per_node * t["s_per_edge"]
    planners = None
    if not args.no_extras:  # the sequential planners' CPU side (configs 2 and 3), same streams
        try:
            pp = _planner_probe()
            c2 = pp.config2(30000, cpu=True)
            c3 = pp.config3(2000, 10000, 2000)
            planners = {
                "rrt_connect_config2": {k: c2.get(k) for k in ("iters", "cpu_us_per_iter",
                                                               "gpu_us_per_iter", "bit_identical")},
                "rrt_star_config3_first_2000_iters": {
                    "cpu_iters_per_s": c3.get("prefix_cpu_iters_per_s"),
                    "gpu_iters_per_s": c3.get("prefix_gpu_iters_per_s"),
                    "bit_identical": c3.get("prefix_bit_identical")},
            }
        except Exception as e:
            planners = {"error": repr(e)}
    return {
        "planners": planners,
        "value": 1.0 / s_node, "unit": UNIT, "cores": 1, "kind": "port",
        "host_cores_available": os.cpu_count(),
        "knn_queries_per_s": 1.0 / t["s_per_query"], "edges_checked_per_s": 1.0 / t["s_per_edge"],
        "sample": "%d k-NN queries x %d nodes (%.1f s) + %d motion checks (%.1f s), single thread; "
                  "nodes/s extrapolated as 1/(s_per_query + %.2f*s_per_edge)"
                  % (t["n_queries"], rows.shape[0], t["t_knn"], t["n_edges"], t["t_edge"],
                     edges_per_node),
    }


def main():
    args = parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
