#!/usr/bin/env python
"""bench.py — the headline benchmark: PRM roadmap build on a 7-D joint box, 1M samples,
k = 10, capsule-obstacle world (BASELINE.json config 4), at 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one complete roadmap build over the resident store through the library's sharded
entry point (oxg_prm_build_knn_dist_dev): the bounding-box hierarchy of the store (rebuilt every
step: "tree_cache" is switched off), the k-NN of every node, the motion check of every candidate
edge and — at N > 1, where the query rows and the candidate edges are sharded over the ranks
against a replicated store (strong scaling: the 1M-node job is fixed) — the NCCL all-gathers the
library issues itself.  Every rank ends with the complete roadmap.

One JSON line is printed by rank 0 (see the task contract): `value` is nodes/s with the store
resident in HBM; `e2e` is the same build through the host-buffer entry point
(oxg_prm_build_knn_dist) with the H2D upload of the samples and the D2H read of the complete
roadmap inside the timed region; `roofline` describes the dominant kernel on EXECUTED work
(tree_knn_kernel: FP64 lane-operations it issued / its own time, against the unfused FP64 issue
peak measured in the same run), `roofline_all_pairs` the all-pairs scan kernel it replaced
(executed FMA lane-instructions against the measured FP32 lane rate), `roofline_edges` the edge
phase (the sequential reference's algorithmic operation count, from the oracle's counters);
`cpu_baseline` is the CPU oracle (single thread, as the reference is) on a bounded sample.
`--impl reference` times the CPU oracle with all host threads instead and never loads liboxg.
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
# dram__bytes_read.sum + dram__bytes_write.sum of one launch at this config, from the `ncu --set
# full` captures summarised in profiles/ (None until captured for the current kernel)
TREE_DRAM_BYTES_NCU = 904850688  # profiles/r02_ncu_full_hot_kernels.csv (746.5 MB read + 158.4 MB written; 1e6 queries)
# the same capture: share of tree_knn_kernel's L1 sector requests that hit L1 (the rest go to L2),
# the L2 read sectors it caused (x 32 B = 81.6 GB for 1e6 queries), the issue-slot utilisation and
# the executed warp instructions
TREE_L1_HIT_NCU = 0.2256
TREE_L2_READ_BYTES_PER_QUERY_NCU = 2551546522 * 32 / 1000000
TREE_ISSUE_ACTIVE_NCU = 0.712
TREE_WARP_INSTRUCTIONS_NCU = 7987274268
ALLPAIRS_DRAM_BYTES_NCU = 2249420000  # profiles/r01_ncu_full_knn_rv_dot_7_2.csv (2.046 GB read + 0.204 GB written)


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


def config_dict(args):
    """The same dict on both arms and at every N (the driver compares them)."""
    return {
        "workload": "config 4: PRM roadmap build, 7-D Panda-like joint box, %d valid uniform "
                    "samples (seed %d), k=%d, %d capsules (worlds.capsule_world seed %d, len<=3, "
                    "r in [1.0,1.6])" % (args.nodes, SEED, args.k, args.capsules, SEED),
        "nodes": args.nodes,
        "k": args.k,
        "dim": DIM,
        "capsules": args.capsules,
        "l2": "store (56 MB) is L2-resident by design; a 256 MiB memset flushes L2 between "
              "steps inside the timed region (<0.1% of a step)",
    }


# ------------------------------------------------------------------------------------------
# workload generation (host; the validity filter that discards invalid samples runs on the GPU
# for the product arm and in the oracle for the reference arm — outside every timed region)
# ------------------------------------------------------------------------------------------
_WORLDS = None


def load_worlds():
    """oxmpl_b200/worlds.py loaded by path: the generators are plain numpy, and importing the
    package would map liboxg.so into the process — which the reference arm must not do."""
    global _WORLDS
    if _WORLDS is None:
        import importlib.util

        spec = importlib.util.spec_from_file_location("oxg_bench_worlds",
                                                      os.path.join(ROOT, "oxmpl_b200", "worlds.py"))
        _WORLDS = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_WORLDS)
    return _WORLDS


def raw_samples(n, seed):
    worlds = load_worlds()
    rng = np.random.default_rng(seed)
    return worlds.PANDA_LO + (worlds.PANDA_HI - worlds.PANDA_LO) * rng.random((n, DIM))


def resolution():
    """space.get_longest_valid_segment_length() * 0.1 (real_vector_state_space.rs:103-118,251-253,
    rrt.rs:97) for the config-4 box; Python floats are IEEE binary64, math.sqrt is correctly
    rounded (tests/test_abi.py pins the same arithmetic against the config-1 constants)."""
    import math

    worlds = load_worlds()
    s = 0.0
    for lo, hi in zip(worlds.PANDA_LO, worlds.PANDA_HI):
        d = float(hi) - float(lo)
        s = s + d * d
    return math.sqrt(s) * 0.05 * 0.1


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
def _set_omp_threads(threads):
    os.environ["OMP_NUM_THREADS"] = str(threads)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
    except OSError:
        pass


def cpu_sample(orc, ow, sp_o, rows, args, res, n_queries, n_edges, threads, counters=None):
    """Times the oracle on a bounded sample of the same workload.  Returns a dict with the
    per-unit times and the extrapolated whole-build nodes/s."""
    n = rows.shape[0]
    k = args.k
    _set_omp_threads(threads)
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
        orc.check_motion(ow, sp_o, frm, to, res, counters=counters)
    t_edge = time.perf_counter() - t0
    s_per_query = t_knn / len(q_rows)
    s_per_edge = t_edge / max(len(i), 1)
    return {"s_per_query": s_per_query, "s_per_edge": s_per_edge, "n_queries": int(len(q_rows)),
            "n_edges": int(len(i)), "t_knn": t_knn, "t_edge": t_edge}


def edges_per_node_cpu(orc, sp_o, rows, k, n_probe=24):
    """Distinct candidate edges per node of the k-NN build = k minus half of the mutual slots,
    measured with the oracle: the k-NN rows of a few nodes and of all their neighbours."""
    n = rows.shape[0]
    _set_omp_threads(os.cpu_count() or 1)
    rng = np.random.default_rng(321)
    probe = np.sort(rng.choice(n, size=min(n_probe, n), replace=False)).astype(np.uint32)
    idx, _ = orc.knn(sp_o, rows, rows[probe], k, exclude=probe)
    nb = np.unique(idx[idx != 0xFFFFFFFF]).astype(np.uint32)
    nidx, _ = orc.knn(sp_o, rows, rows[nb], k, exclude=nb)
    row_of = {int(v): t for t, v in enumerate(nb)}
    slots = mutual = 0
    for r, i in enumerate(probe):
        for j in idx[r]:
            if j == 0xFFFFFFFF:
                continue
            slots += 1
            mutual += int(i in nidx[row_of[int(j)]])
    if slots == 0:
        return float(k)
    return (slots - 0.5 * mutual) / len(probe)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle  # the only native library this arm loads

    worlds = load_worlds()
    orc = oracle.get()
    threads = os.cpu_count() or 1
    caps = worlds.capsule_world(SEED, worlds.PANDA_LO, worlds.PANDA_HI, args.capsules)
    sp_o = orc.space(DIM)
    ow, keep = orc.world(DIM, capsules=caps)
    res = resolution()
    rows = valid_samples(args.nodes, lambda s: orc.is_valid(ow, sp_o, s))
    edges_per_node = edges_per_node_cpu(orc, sp_o, rows, args.k)
    n_q = max(threads * 24, 64)
    n_e = max(threads * 1500, 4000)
    times = []
    for it in range(args.warmup + args.steps):
        per_step = cpu_sample(orc, ow, sp_o, rows, args, res, n_q, n_e, threads)
        if it >= args.warmup:
            times.append(per_step)
    s_q = float(np.mean([t["s_per_query"] for t in times]))
    s_e = float(np.mean([t["s_per_edge"] for t in times]))
    s_node = s_q + edges_per_node * s_e
    value = 1.0 / s_node
    ms_per_step = 1e3 * float(np.mean([t["t_knn"] + t["t_edge"] for t in times]))
    sample = ("per step: %d k-NN queries x %d nodes + %d motion checks, OpenMP/threads=%d; nodes/s "
              "extrapolated as 1/(s_per_query + %.3f*s_per_edge), %.3f = distinct candidate edges per "
              "node measured with the oracle" % (n_q, args.nodes, n_e, threads, edges_per_node,
                                                 edges_per_node))
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "knn_queries_per_s": 1.0 / s_q, "edges_checked_per_s": 1.0 / s_e,
        "edges_per_node": edges_per_node,
        "note": "CPU oracle (C++ restatement of the Rust reference, which cannot be built in this "
                "image: no cargo/rustc); extrapolated from a bounded sample",
        "liboxg_mapped": any("liboxg.so" in ln for ln in open("/proc/self/maps")),
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
    from oxmpl_b200 import sharding
    from oxmpl_b200._lib import PrmStats, check, lib

    worlds = load_worlds()
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
        dist.init_process_group("nccl", device_id=dev)   # rendezvous only (carries the NCCL id)
    dctx = sharding.Dist.from_torch(dist, rank, world_size)

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
    rb, re = dctx.row_range(n)
    oxg.tuning_set("tree_cache", 0)   # every build pays for its own hierarchy

    with torch.cuda.stream(ext):
        d_idx_all = torch.empty((n, k), dtype=torch.int32, device=dev)       # u32 bit patterns
        d_dist_all = torch.empty((n, k), dtype=torch.float64, device=dev)
        d_valid_all = torch.empty((n, k), dtype=torch.uint8, device=dev)
        d_counters = torch.zeros(2, dtype=torch.int64, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ext.synchronize()
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731

    def step():
        """One sharded roadmap build, device-resident, all exchanges inside the library."""
        with torch.cuda.stream(ext):
            flush.zero_()
        check(lib.oxg_prm_build_knn_dist_dev(dctx.handle, gw.handle, st.handle, k, float(res),
                                             P(d_idx_all), P(d_dist_all), P(d_valid_all), P(d_counters)))

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
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record(ext)
    for _ in range(args.steps):
        step()
    stop.record(ext)
    barrier()
    launches = oxg.kernel_launches()
    elapsed_ms = start.elapsed_time(stop)
    counters = d_counters.cpu().numpy()
    n_checked_local = int(counters[0])

    # ---- phase attribution (untimed for the headline): this rank's k-NN rows, then its edges ----
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    knn_ms, edge_ms = [], []
    tcount = None
    for it in range(4):
        oxg.tuning_set("tree_stats", 2 if it == 0 else 1)   # one counting pass, then timing passes
        with torch.cuda.stream(ext):
            flush.zero_()
            ev[0].record(ext)
        check(lib.oxg_knn_rows_dev(st.handle, rb, re, k, P(d_idx_all[rb:re]), P(d_dist_all[rb:re])))
        ev[1].record(ext)
        ext.synchronize()
        tstats = st.tree_stats()
        if it == 0:
            tcount = tstats
            continue
        if world_size == 1:
            check(lib.oxg_prm_edges_dev(gw.handle, st.handle, P(d_idx_all), k, float(res), rb, re,
                                        P(d_valid_all), P(d_counters)))
            ev[2].record(ext)
            ext.synchronize()
            edge_ms.append(ev[1].elapsed_time(ev[2]))
        knn_ms.append(ev[0].elapsed_time(ev[1]))
    oxg.tuning_set("tree_stats", 0)
    for key in ("leaves_visited", "box_groups_tested", "candidates", "merges", "exact_leaves"):
        tstats[key] = tcount[key]    # counters from the counting pass, times from a timing pass
    knn_phase_ms = float(np.median(knn_ms))
    edge_phase_ms = float(np.median(edge_ms)) if edge_ms else None

    # ---- the all-pairs scan kernel this round's hierarchy replaces: one full wave, N = 1 only ----
    allpairs = None
    if world_size == 1 and n >= 200_000:
        oxg.tuning_set("knn_tree", 0)
        wave_rows = min(n, 148 * 3 * 256)
        t_ms = []
        for it in range(4):
            with torch.cuda.stream(ext):
                flush.zero_()
                ev[0].record(ext)
            check(lib.oxg_knn_rows_dev(st.handle, 0, wave_rows, k, P(d_idx_all[:wave_rows]),
                                       P(d_dist_all[:wave_rows])))
            ev[1].record(ext)
            ext.synchronize()
            if it:
                t_ms.append(ev[0].elapsed_time(ev[1]))
        oxg.tuning_set("knn_tree", -1)
        allpairs = {"rows": wave_rows, "ms": float(np.median(t_ms))}

    # ---- end-to-end through the host-buffer entry points ----------------------------------------
    # Every rank uploads ITS slice of the samples from pinned host memory (oxg_store_append_dist:
    # the other slices arrive over NVLink), all ranks build, rank 0 reads the complete roadmap back.
    want_host = rank == 0
    h_idx = torch.empty((n, k), dtype=torch.int32).pin_memory() if want_host else None
    h_dist = torch.empty((n, k), dtype=torch.float64).pin_memory() if want_host else None
    h_valid = torch.empty((n, k), dtype=torch.uint8).pin_memory() if want_host else None
    HP = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
    stats = PrmStats()
    slice_ptr = rows_pinned[rb:re].data_ptr()
    first = C.c_uint64(0)

    def e2e_step():
        st.clear()
        check(lib.oxg_store_append_dist(dctx.handle, st.handle, C.c_void_p(slice_ptr), re - rb,
                                        C.byref(first)))                              # H2D: the samples
        check(lib.oxg_prm_build_knn_dist(dctx.handle, gw.handle, st.handle, k, float(res),
                                         HP(h_idx), HP(h_dist), HP(h_valid), C.byref(stats)))  # D2H: roadmap

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop() if rank == 0 else None
    # the roadmap the e2e path delivered must be the one the resident path holds
    # (d_idx_all still holds the last complete table of the resident-store loop at N > 1)
    same = bool(torch.equal(h_idx, d_idx_all.cpu())) if (world_size > 1 and want_host) else None

    # ---- max over ranks ---------------------------------------------------------------------
    vals = torch.tensor([elapsed_ms, e2e_s, knn_phase_ms, float(n_checked_local)], dtype=torch.float64,
                        device=dev)
    if world_size > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = vals, vals
    elapsed_ms_max, e2e_s_max, knn_ms_max = float(mx[0]), float(mx[1]), float(mx[2])
    edges_checked_total = float(sm[3])

    if rank == 0:
        ms_per_step = elapsed_ms_max / args.steps
        value = n / (ms_per_step * 1e-3)
        peak64_ops, _ = oxg.fp64_peak()
        peak32_ops, _ = oxg.fp32_peak()
        rows_rank0 = re - rb
        # dominant kernel: tree_knn_kernel, on EXECUTED work.  Per visited leaf every lane runs the
        # FP32 filter (D subtractions + D FMAs); per box group 5D FP32 operations (2 subtractions,
        # a 3-way max, a multiply and an add per coordinate, directed rounding); the leaves the
        # filter cannot discard whole run the reference-order f64 evaluation (3D - 1 separately
        # rounded operations per lane).  The list merges are shuffles: not counted.  frac = the time
        # both pipes would need at their measured peaks / the kernel's time.
        kern_ms = tstats["query_kernel_ms"]
        q_n = max(tstats["queries"], 1.0)
        f32_ops = (tstats["leaves_visited"] * 2.0 * DIM + tstats["box_groups_tested"] * 5.0 * DIM) * 32.0
        f64_ops = tstats["exact_leaves"] * 32.0 * (3.0 * DIM - 1.0)
        t_pipes_ms = (f32_ops / peak32_ops + f64_ops / peak64_ops) * 1e3
        # Memory side.  The store and its hierarchy are L2-resident by design (the leaf copies and
        # boxes of config 4 are ~120 MB of the 126 MB L2, DRAM traffic 0.66 GB per launch), so the
        # memory roofline is the L2, not HBM.  Bytes the warps request: per visited leaf D x 32 f32,
        # per exact leaf D x 32 f64, per box group 2D x 32 f32, per candidate one 32-byte sector of
        # `perm`; the ncu capture in profiles/ gives the share that hits L1.
        req_bytes = (tstats["leaves_visited"] * DIM * 32 * 4.0 + tstats["exact_leaves"] * DIM * 32 * 8.0 +
                     tstats["box_groups_tested"] * 2 * DIM * 32 * 4.0 + tstats["candidates"] * 32.0)
        # achieved: the L2 read traffic ncu measured per query (the counter-based estimate
        # req_bytes x (1 - L1 hit rate) is ~9 % higher; the smaller figure is reported)
        l2_bytes = min(req_bytes * (1.0 - TREE_L1_HIT_NCU), TREE_L2_READ_BYTES_PER_QUERY_NCU * q_n)
        l2_peak_bps, _ = oxg.l2_peak()
        roof = {
            "kernel": "oxg::tree_knn_kernel<7>", "bound": "l2",
            "achieved": l2_bytes / (kern_ms * 1e-3) / 1e9, "peak": l2_peak_bps / 1e9, "unit": "GB/s",
            "frac": l2_bytes / (kern_ms * 1e-3) / l2_peak_bps, "traffic": TREE_DRAM_BYTES_NCU,
            "launch_ms": kern_ms,
            "requested_bytes_per_launch": req_bytes, "l1_hit_rate_ncu": TREE_L1_HIT_NCU,
            "l2_read_bytes_per_query_ncu": TREE_L2_READ_BYTES_PER_QUERY_NCU,
            "requested_bytes_per_query": req_bytes / q_n,
            "pipes": {"what": "time the executed FP32 (leaf filter, box tests) and f64 (exact leaves) "
                              "lane-operations would need at the measured pipe peaks / kernel time",
                      "frac": t_pipes_ms / kern_ms, "executed_f32_lane_ops": f32_ops,
                      "executed_f64_lane_ops": f64_ops, "fp32_peak_tops": peak32_ops / 1e12,
                      "fp64_peak_tops": peak64_ops / 1e12},
            "leaves_visited_per_query": tstats["leaves_visited"] / q_n,
            "exact_leaves_per_query": tstats["exact_leaves"] / q_n,
            "box_groups_per_query": tstats["box_groups_tested"] / q_n,
            "candidates_per_query": tstats["candidates"] / q_n,
            "pairs_evaluated_fraction_of_all_pairs": tstats["leaves_visited"] * 32.0 / (q_n * n),
            "all_pairs_equivalent_tflops": 3.0 * DIM * float(rows_rank0) * float(n) / (kern_ms * 1e-3) / 1e12,
            "hierarchy_build_ms": tstats["build_ms"],
            "peak_source": "oxg_l2_peak(): all SMs reading a 48 MB L2-resident buffer with coalesced 16-byte "
                           "ld.global.cg, measured in this run (MEASURED_PEAKS.json's HBM figure, "
                           "%s GB/s, does not bound an L2-resident working set; the pipe peaks are "
                           "oxg_fp32_peak / oxg_fp64_peak)" % _measured_hbm(),
            "issue_slots": {"what": "smsp__issue_active of the ncu capture in profiles/ (same workload): the "
                                    "share of the schedulers' issue cycles the kernel uses; the executed "
                                    "warp instructions per query are the quantity left to shrink",
                            "frac_ncu": TREE_ISSUE_ACTIVE_NCU,
                            "warp_instructions_per_query_ncu": TREE_WARP_INSTRUCTIONS_NCU / 1e6},
            "note": "an irregular traversal of ~95 dependent steps per query over 1.0 % of the all-pairs "
                    "pair count; the working set is L2-resident, so the memory roofline is the L2's. "
                    "ncu: issue slots 71 %, long-scoreboard stalls 2.5 per issue, L2 hit 98 % -- the "
                    "kernel is bound by instruction issue first and L2 latency second, not by bandwidth "
                    "or arithmetic",
            "hbm_peak_gbs_measured": _measured_hbm(),
        }
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world_size,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(args),
            "sharding": "rows and candidate edges over %d rank(s), replicated store; NCCL inside "
                        "liboxg: one grouped all-gather (idx + dist rows), one all-gather (validity "
                        "rows)" % world_size,
            "knn_queries_per_s": float(rows_rank0) / (knn_ms_max * 1e-3) * world_size,
            "edges_checked_per_step": edges_checked_total,
            "phase_ms": {"hierarchy_build": tstats["build_ms"], "knn_incl_build": knn_ms_max,
                         "edges": edge_phase_ms,
                         "rest_of_step": ms_per_step - knn_ms_max - (edge_phase_ms or 0.0)},
            "e2e": {"value": n / e2e_s_max, "unit": UNIT,
                    "h2d_bytes_per_step": int(n * DIM * 8),
                    "d2h_bytes_per_step": int(n * k * (4 + 8 + 1)),
                    "ms_per_step": e2e_s_max * 1e3,
                    "path": "oxg_store_clear + oxg_store_append_dist(each rank's slice of the pinned host "
                            "rows; the other slices arrive over NVLink) + oxg_prm_build_knn_dist(rank 0 "
                            "reads the complete roadmap into pinned host memory); the byte counts are the "
                            "job's totals; wall clock around synchronous calls, max over ranks",
                    "roadmap_equals_resident_run": same},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
        }
        if edge_phase_ms:
            line["edges_checked_per_s"] = edges_checked_total / (edge_phase_ms * 1e-3)
        if allpairs:
            executed = 1.0 * DIM * float(allpairs["rows"]) * float(n)   # FMA lane-instructions issued
            line["roofline_all_pairs"] = {
                "kernel": "oxg::knn_rv_dot_kernel<7,2>", "bound": "fp32",
                "what": "the all-pairs scan (oxg_tuning_set knn_tree=0) on one full wave of %d query "
                        "rows x %d states; not on the headline path any more" % (allpairs["rows"], n),
                "launch_ms": allpairs["ms"],
                "achieved": executed / (allpairs["ms"] * 1e-3) / 1e12, "peak": peak32_ops / 1e12,
                "unit": "T FMA lane-instructions/s", "frac": executed / (allpairs["ms"] * 1e-3) / peak32_ops,
                "traffic": ALLPAIRS_DRAM_BYTES_NCU,
                "algorithmic_3D_flop_rate_tflops": 3.0 * executed / (allpairs["ms"] * 1e-3) / 1e12,
                "full_table_ms_at_this_rate": allpairs["ms"] * n / allpairs["rows"],
                "peak_source": "oxg_fp32_peak(): packed FP32 lane-instruction rate measured in this run",
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
            fl = (line.get("cpu_baseline") or {}).get("algorithmic_flops_per_edge")
            if fl and edge_phase_ms:
                alg = fl * edges_checked_total
                line["roofline_edges"] = {
                    "kernel": "oxg::motion_cull_kernel<7,4> (+ mark / scatter)", "bound": "fp64",
                    "launch_ms": edge_phase_ms,
                    "algorithmic_flops_per_edge": fl,
                    "achieved": alg / (edge_phase_ms * 1e-3) / 1e12, "peak": peak64_ops / 1e12,
                    "unit": "TFLOP/s (algorithmic: the sequential reference's operation count)",
                    "frac": alg / (edge_phase_ms * 1e-3) / peak64_ops, "traffic": None,
                    "note": "SURVEY §8(d) motion-check cost model counted by the oracle on the "
                            "cpu_baseline edge sample (states x obstacles the sequential loop evaluates "
                            "before its first hit).  The kernel executes far less: one certified FP32 "
                            "midpoint-ball test per (edge, obstacle) rules out ~99% of the obstacles "
                            "for all steps at once, so the algorithmic rate can exceed the FP64 peak",
                }
        print(json.dumps(line), flush=True)
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
    dctx.close()
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
    from oracle.oracle import OrcCounters

    orc = oracle.get()
    sp_o = orc.space(DIM)
    ow, keep = orc.world(DIM, capsules=caps)
    cnt = OrcCounters()
    t = cpu_sample(orc, ow, sp_o, rows, args, res, args.cpu_queries, args.cpu_edges, 1, counters=cnt)
    edges_per_node = edges_per_node_cpu(orc, sp_o, rows, args.k)
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
        "edges_per_node": edges_per_node,
        "algorithmic_flops_per_edge": float(cnt.flops) / max(t["n_edges"], 1),
        "states_per_edge": float(cnt.states) / max(t["n_edges"], 1),
        "sample": "%d k-NN queries x %d nodes (%.1f s) + %d motion checks (%.1f s), single thread; "
                  "nodes/s extrapolated as 1/(s_per_query + %.3f*s_per_edge)"
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
