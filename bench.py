#!/usr/bin/env python
"""Headline benchmark: seconds per interior-point iteration of SMCP's feasible-start solver
(method "M1" of the reference's benchmark tables = ``solve_feas`` with ``kktsolver='chol'``,
doc/source/benchmarks/index.rst:91) on the north-star configuration of BASELINE.json:
rand_SDP n = 2000 on a sparse aggregate pattern, m = 10 000 (configs[2]; it fits one GPU:
H is 800 MB), each A_i with round(0.005 |V|) non-zeros like the reference's own non-chordal
benchmark (doc/source/benchmarks/index.rst, "Nonchordal sparsity patterns").  At N = 1 the
same JSON line carries BASELINE configs[1] (band SDP n = 5000, bandwidth 5, m = 1000) under
``secondary``.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload NAME]

A *step* is one IPM iteration of ``chordalsolver_feas`` through the public API
(``SDP.solve_feas(kktsolver='chol')``): chordal completion of the iterate, assembly of the
Schur complement H, its dense Cholesky, the Newton solves with iterative refinement and the
step-length probes.  W warm-up iterations, then exactly K timed iterations.

* ``e2e``   – wall-clock seconds per iteration through the public API with host buffers
  (every iteration moves its m-vectors and reductions across PCIe; bytes counted).  Lead with
  this number.
* ``value`` – the same K iterations, device time only (sum of the CUDA-event durations of
  every kernel family on the library's streams): the iteration without host gaps.
* ``schur_potrf`` – the part of the step BASELINE.json's metric singles out: device seconds of
  the Schur-complement assembly + the Cholesky of H per iteration (event pairs around the
  C-ABI regions, no synchronisation), with SURVEY 8(d)'s flop model and the flops actually
  executed.  This is the part that is sharded over the GPUs.
* ``roofline`` – the dominant kernel family of the step against the measured peak.
* ``cpu_baseline`` – the CPU oracle (the reference's algorithm restated on NumPy/SciPy,
  ``oracle/``; the sparse-constraint Schur columns through the reference's own compiled
  ``misc.SCMcolumn2`` when ``oracle/_ref`` is built) timed on this box's host cores at the
  workload's real starting iterate; the per-column Schur loop runs on a stratified sample of
  columns and is scaled to m, every other phase runs in full.

With N > 1 (torchrun, one rank per GPU) the columns of H are owned 1-D block-cyclically
(256-column blocks): every rank assembles its own blocks, the factorisation broadcasts
panels with NCCL and every rank updates the blocks it owns; chordal single-matrix
operations are replicas.  The problem is fixed, so scaling is "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "rand_n2000_m10000": dict(kind="rand", n=2000, m=10000, density=0.005),
    "band_n5000_m1000_bw5": dict(kind="band", n=5000, m=1000, bw=5),
    # small stand-ins for tests of this script
    "rand_n300_m400": dict(kind="rand", n=300, m=400, density=0.01),
    "band_n500_m100_bw3": dict(kind="band", n=500, m=100, bw=3),
}
DEFAULT = "rand_n2000_m10000"
SECONDARY = "band_n5000_m1000_bw5"
DIST_BLOCK = 256


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=sorted(WORKLOADS))
    ap.add_argument("--secondary", default=None, choices=sorted(WORKLOADS) + ["none"],
                    help="second workload reported under 'secondary' (default: %s at N = 1)" % SECONDARY)
    ap.add_argument("--no-solve", action="store_true", help="skip the time-to-solve run")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def wait_first(self, timeout=10.0):
        # nvidia-smi needs ~1 s to attach to the driver and perturbs CUDA calls while it does: wait for
        # its first sample so that only steady-state polling overlaps the timed region
        t0 = time.perf_counter()
        while not self.rows and time.perf_counter() - t0 < timeout and self.proc is not None:
            time.sleep(0.05)

    def stop(self):
        if self.proc is None:
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
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_setup():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    pg = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        pg = dist
    return rank, world, local, pg


def measured_fp64_peak():
    """FP64 GEMM peak of this GPU: cuBLAS DGEMM through torch.matmul, 8192^3, best of 5.
    Used ONLY as the roofline denominator (MEASURED_PEAKS.json has no FP64 entry)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return None
        dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        n = 8192
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        torch.matmul(a, b)
        torch.cuda.synchronize(dev)
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize(dev)
            best = min(best, e0.elapsed_time(e1))
        del a, b
        torch.cuda.empty_cache()
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    except Exception:
        return None


# families whose LaunchScope work is algorithmic BYTES per launch; *_dmma / potrf_tile / trsm_slab carry flops;
# the chordal families carry the number of matrices processed (16*|Vp| bytes each)
BYTE_FAMILIES = {"potrs", "amap_dense", "amap", "aadj", "potrf_panel", "gemm_thin", "scm_kstream"}
FLOP_FAMILIES = {"potrf_tile", "trsm_slab", "gemm_smallk", "potrf_dmma", "front_potrf"}
NO_MODEL = {"front_elem", "trsm_cluster", "potrs_trtri", "potrf_panel_t", "thin_up", "thin_down", "thin_compl_tail", "thin_chol", "thin_hinv_local", "thin_hinv_sweep", "thin_trsm", "front_elem_batch", "setup", "scm_position", "scm_sparse", "chordal_trsm", "front_trsm_diag",
            "level1", "reduce", "scatter_cols"}


def roofline_of(name, f, nvp, hbm_peak, hbm_src, fp64_peak):
    """achieved = algorithmic bytes (or flops) per launch / average launch duration (CUDA events on
    the library's stream).  Chordal kernels: 16*|Vp| bytes per matrix (one read + one write of every
    pattern entry); GEMMs / factorisations / triangular solves: their flops;
    potrs: 8*m^2 bytes (the factor is read twice); amap/aadj: the stored entries of Av."""
    if name in NO_MODEL or f["work"] <= 0:
        return {"kernel": name, "bound": None, "achieved": None, "peak": None, "unit": None, "frac": None,
                "traffic": None, "note": "no per-launch work model for this family"}
    per_launch_s = f["ms"] * 1e-3 / max(1, f["launches"])
    work = f["work"] / max(1, f["launches"])
    if name.endswith("_dmma") or name in FLOP_FAMILIES:
        ach = work / per_launch_s / 1e12
        peak = fp64_peak if fp64_peak else 40.0
        return {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": None,
                "peak_source": ("measured in this run: cuBLAS DGEMM 8192^3 via torch.matmul, best of 5 "
                                "(MEASURED_PEAKS.json has no FP64 entry)" if fp64_peak
                                else "nominal FP64 tensor peak (no measurement)")}
    byts = work if name in BYTE_FAMILIES else 16.0 * nvp * work
    ach = byts / per_launch_s / 1e9
    out = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
           "frac": ach / hbm_peak, "traffic": None, "peak_source": hbm_src}
    if name not in BYTE_FAMILIES:
        out["matrices_per_launch"] = work
    return out


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------
def rand_pattern(n):
    """Sparse aggregate pattern of the rand workloads: tridiagonal plus 6n random off-diagonal pairs
    (|V| ~ 8n > m: the A_i live on V)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(0)
    e = rng.integers(0, n, size=(6 * n, 2))
    I = np.concatenate([np.maximum(e[:, 0], e[:, 1]), np.arange(n), np.arange(1, n)])
    J = np.concatenate([np.minimum(e[:, 0], e[:, 1]), np.arange(n), np.arange(0, n - 1)])
    return sp.coo_matrix((np.ones(len(I)), (I, J)), shape=(n, n))


def build_problem(workload):
    """The generators run their cone tests (mk_rand) on the installed backend."""
    import smcp_b200 as S
    w = WORKLOADS[workload]
    if w["kind"] == "band":
        return S.band_SDP(w["n"], w["m"], w["bw"], seed=0)
    return S.rand_SDP(rand_pattern(w["n"]), w["m"], density=w["density"], seed=0)


def config_of(workload, world=None):
    w = WORKLOADS[workload]
    cfg = {"workload": workload, "n": w["n"], "m": w["m"], "solver": "chordalsolver_feas",
           "kktsolver": "chol", "scaling_mode": "primal"}
    if w["kind"] == "rand":
        cfg["nnz_per_constraint"] = "round(%g |V|)" % w["density"]
    else:
        cfg["bandwidth"] = w["bw"]
    if world is not None:
        # identical in both arms (b200 / reference) so that the driver sees the same config
        cfg["l2"] = "working set > L2 per iteration (H alone is %d MB)" % (8 * w["m"] ** 2 // 2 ** 20)
        cfg["parallelism"] = "schur-columns block-cyclic x%d (blocks of %d columns)" % (world, DIST_BLOCK)
    return cfg


def flop_model(prob):
    """Schur + potrf flops per iteration of the REFERENCE's algorithm,
    (m - Ns) F_H + F_ip + F_T2 + m^3/3, and the flops this implementation executes for the same H."""
    symb, m, Ns = prob.symb, prob.m, prob.Ns
    nn = np.diff(symb.snptr).astype(np.float64)
    nj = np.diff(symb.rowptr).astype(np.float64)
    na = nj - nn
    F_H = float(np.sum(4 * nn ** 3 + 6 * na * nn ** 2 + 6 * na ** 2 * nn))
    Av = prob.Av
    nnzc = np.diff(Av.indptr).astype(np.float64)
    md = m - Ns
    tail = np.cumsum(nnzc[::-1])[::-1]                  # sum_{i >= j} nnz(A_i)
    F_ip = 2.0 * float(np.sum(tail[:md]))
    nnzL = float(symb.nvp)
    F_T2 = F_T2x = 0.0
    if Ns:
        Ip, Jp = symb.Ip, symb.Jp
        Kl = np.empty(Ns)
        for j in range(md, m):
            r = Av.indices[Av.indptr[j]:Av.indptr[j + 1]]
            Kl[j - md] = len(np.unique(np.concatenate([Ip[r], Jp[r]])))
        F_T2 = float(np.sum(4.0 * nnzL * Kl + 4.0 * nnzc[md:] * tail[md:]))
        # executed: dense inverse of S once (two chordal solves with n right-hand sides), then the position
        # form: 5 flops per (position, entry of A_j) + 2 per entry of the rows i >= j
        rows_sp = Av.indices[Av.indptr[md]:Av.indptr[m]]
        npos = len(np.unique(rows_sp))
        F_T2x = 4.0 * nnzL * symb.n + float(np.sum(5.0 * npos * nnzc[md:] + 2.0 * tail[md:]))
    F_potrf = m ** 3 / 3.0
    return {"F_H_per_column": F_H, "dense_columns": int(md), "sparse_columns": int(Ns),
            "model_flops": md * F_H + F_ip + F_T2 + F_potrf,
            "executed_flops": md * F_H + F_ip + F_T2x + F_potrf,
            "potrf_flops": F_potrf}


class IterTimer:
    def __init__(self, sync):
        self.sync = sync
        self.t = {}
        self.marks = {}

    def __call__(self, name, it):
        self.sync()
        self.t[it] = time.perf_counter()
        fn = self.marks.get(it)
        if fn:
            fn()


def measure_workload(workload, W, K, ctx, rank, world, pg, local, do_solve, hbm, fp64_peak):
    """Two passes over the same W + K iterations (wall clock, then per-family device time) and, at
    N = 1, the solve to the default tolerances."""
    from smcp_b200 import solvers
    from smcp_b200.device import TRAFFIC
    solvers.options["show_progress"] = False
    solvers.options["maxiters"] = W + K
    P = build_problem(workload)
    n, m = P.n, P.m
    start = {"x": P._X0}

    def barrier():
        if pg is not None:
            pg.barrier()

    # ---- pass 1: end-to-end wall clock through the public API ------------------------
    timer = IterTimer(ctx.sync)
    snap = {}

    def mark_start():
        snap["traffic"] = dict(TRAFFIC)
        snap["launches"] = ctx.launch_count()
        ctx.region_reset()
        barrier()
        ctx.sync()
        timer.t[W] = time.perf_counter()
        ctx.timer_start()

    # CUDA events on the library's stream bracket the K timed iterations: `value` (concurrent lanes make
    # the sum of kernel durations larger than the elapsed device time, so the sum is reported separately)
    timer.marks[W] = mark_start
    snap["span"] = {}
    for it_ in range(W + 1, W + K + 1):      # elapsed device time since the start mark, re-read at every later iteration
        timer.marks[it_] = (lambda i=it_: snap["span"].__setitem__(i, ctx.timer_stop()))
    solvers._iteration_hook = timer
    sampler = ClockSampler(local)
    sampler.start()
    sampler.wait_first()
    barrier()
    sol = P.solve_feas(kktsolver="chol", primalstart=start)
    ctx.sync()
    barrier()
    clocks = sampler.stop()
    iters = sol["iterations"]
    K_req = K
    last = max(timer.t) if timer.t else 0
    if last < W + K:
        # the problem converged before warmup + steps iterations: time the iterations there are and say so
        K = last - W
        if K < 1:
            raise SystemExit("bench: the solver stopped after %d iterations, fewer than warmup + 1 = %d" % (iters, W + 1))
    e2e_s = (timer.t[W + K] - timer.t[W]) / K
    snap["span_ms"] = snap["span"][W + K]
    # counters: maxiters = W + K stops the loop there; after it only the final residual evaluation runs
    h2d = (TRAFFIC["h2d"] - snap["traffic"]["h2d"]) // K
    d2h = (TRAFFIC["d2h"] - snap["traffic"]["d2h"]) // K
    n_launch = ctx.launch_count() - snap["launches"]
    regions = {nm: {"ms_per_step": 0.0, "calls_per_step": 0.0} for nm in ("kkt_assemble", "kkt_factor", "kkt_solve", "kkt_allgather")}
    ops_ms = {}
    for nm in ctx.region_names():
        ms, calls = ctx.region_get(nm)
        if nm.startswith("kkt_"):
            regions[nm] = {"ms_per_step": ms / K, "calls_per_step": calls / K}
        elif calls:
            ops_ms[nm] = {"ms_per_step": ms / K, "calls_per_step": calls / K}
    prob = solvers._last_problem
    fm = flop_model(prob) if prob is not None else None
    nvp = prob.symb.nvp if prob is not None else 0

    # ---- pass 2: same iterations, device time per kernel family (CUDA events) --------
    fam = {}
    timer2 = IterTimer(ctx.sync)
    timer2.marks[W] = lambda: (ctx.prof_reset(), ctx.prof_enable(True))
    timer2.marks[W + K] = lambda: ctx.prof_enable(False)
    solvers._iteration_hook = timer2
    P.solve_feas(kktsolver="chol", primalstart=start)
    ctx.prof_enable(False)
    solvers._iteration_hook = None
    dev_ms = 0.0
    for nm in ctx.prof_names():
        ms, cnt = ctx.prof_get(nm)
        if cnt:
            fam[nm] = {"ms": ms, "launches": cnt, "work": ctx.prof_get_work(nm)}
            dev_ms += ms
    ksum_s = dev_ms * 1e-3 / K
    dev_s = snap["span_ms"] * 1e-3 / K
    sp_s = (regions["kkt_assemble"]["ms_per_step"] + regions["kkt_factor"]["ms_per_step"]
            + regions["kkt_allgather"]["ms_per_step"]) * 1e-3

    # max over ranks
    if pg is not None:
        import torch
        t = torch.tensor([e2e_s, dev_s, sp_s, ksum_s], dtype=torch.float64)
        pg.all_reduce(t, op=pg.ReduceOp.MAX)
        e2e_s, dev_s, sp_s, ksum_s = float(t[0]), float(t[1]), float(t[2]), float(t[3])

    out = {"steps": int(K), "steps_requested": int(K_req), "e2e_s": e2e_s, "dev_s": dev_s, "ksum_s": ksum_s, "h2d": int(h2d), "d2h": int(d2h), "launches": int(n_launch), "ops": ops_ms,
           "clocks": clocks, "status": sol["status"], "iterations": int(iters), "n": n, "m": m, "nvp": int(nvp)}
    if rank != 0:
        return out
    hbm_peak, hbm_src = hbm
    top = max(fam.items(), key=lambda kv: kv[1]["ms"])[0] if fam else None
    out["roofline"] = roofline_of(top, fam[top], nvp, hbm_peak, hbm_src, fp64_peak) if top else None
    if out["roofline"] is not None:
        out["roofline"]["share_of_summed_kernel_time"] = fam[top]["ms"] / dev_ms if dev_ms > 0 else None
        out["roofline"]["how"] = ("family with the largest summed kernel time in the per-launch profile pass (CUDA events around "
                                  "every launch on the library stream; the concurrent lanes of the top set run one after the other "
                                  "in that pass, each launch with the whole GPU to itself)")
    per_family = {}
    for nm, f in fam.items():
        if f["ms"] >= 0.03 * dev_ms:
            r = roofline_of(nm, f, nvp, hbm_peak, hbm_src, fp64_peak)
            per_family[nm] = {"ms_per_step": f["ms"] / K, "launches_per_step": f["launches"] / K,
                              "bound": r["bound"], "achieved": r["achieved"], "unit": r["unit"], "frac": r["frac"]}
    out["family_rooflines"] = per_family
    out["kernel_ms_per_step"] = {k: v["ms"] / K for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if out["roofline"] is not None and os.path.exists(traffic_file):
        with open(traffic_file) as fh:
            ent = json.load(fh).get(out["roofline"]["kernel"])
        if isinstance(ent, dict):
            out["roofline"]["traffic"] = ent.get("bytes")
            out["roofline"]["traffic_unit"] = "bytes/launch (dram__bytes_read.sum + dram__bytes_write.sum)"
            out["roofline"]["traffic_source"] = ent.get("source")
    peak = fp64_peak if fp64_peak else 40.0
    out["schur_potrf"] = {
        "seconds_per_iter": sp_s, "regions_ms_per_step": regions, "flop_model": fm,
        "tflops_model": (fm["model_flops"] / sp_s / 1e12) if fm and sp_s > 0 else None,
        "tflops_executed": (fm["executed_flops"] / sp_s / 1e12) if fm and sp_s > 0 else None,
        "frac_of_fp64_peak_executed": (fm["executed_flops"] / sp_s / 1e12 / peak) if fm and sp_s > 0 else None,
        "potrf_tflops": (fm["potrf_flops"] / (regions["kkt_factor"]["ms_per_step"] * 1e-3) / 1e12)
        if fm and regions["kkt_factor"]["ms_per_step"] > 0 else None,
        "note": "device seconds of smcp_kkt_assemble* + smcp_kkt_factor* per iteration, max over ranks; model = "
                "SURVEY 8(d) flops of the reference's algorithm, executed = flops of the algorithm run here"}

    # time-to-solve at SMCP's default tolerances (the second half of BASELINE.json's metric)
    out["time_to_solve"] = None
    if do_solve and world == 1:
        solvers.options["maxiters"] = 100
        timer3 = IterTimer(ctx.sync)
        solvers._iteration_hook = timer3
        sol3 = P.solve_feas(kktsolver="chol", primalstart=start)
        ctx.sync()
        solvers._iteration_hook = None
        ks = sorted(timer3.t)
        out["time_to_solve"] = {
            "status": sol3["status"], "iterations": int(sol3["iterations"]),
            "seconds": (timer3.t[ks[-1]] - timer3.t[ks[0]]) if len(ks) > 1 else None,
            "primal_objective": sol3["primal objective"], "dual_objective": sol3["dual objective"],
            "primal_infeasibility": sol3["primal infeasibility"], "gap": sol3["gap"]}
    return out


def run_b200(args):
    rank, world, local, pg = dist_setup()
    # NCCL's communicator lines stay visible: NCCL writes its INFO lines to stdout, so for N > 1 file
    # descriptor 1 is pointed at stderr for the whole run and the one JSON line goes to the saved stdout
    json_fd = None
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        os.environ.pop("NCCL_DEBUG_FILE", None)
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
    os.environ["LOCAL_RANK"] = str(local)
    from smcp_b200 import device
    from smcp_b200.device import Context
    ctx = Context.get(local)
    if world > 1:
        import torch
        idt = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            idt = torch.frombuffer(bytearray(device.comm_unique_id()), dtype=torch.uint8).clone()
        pg.broadcast(idt, src=0)
        device.init_comm(rank, world, bytes(idt.numpy().tobytes()), block=DIST_BLOCK, device=local)
        print("smcp_b200: NCCL communicator ready: rank %d nranks %d (column blocks of %d)" % (rank, world, DIST_BLOCK),
              file=sys.stderr, flush=True)

    W, K = args.warmup, args.steps
    hbm = load_peaks()
    fp64_peak = measured_fp64_peak() if rank == 0 else None
    main = measure_workload(args.workload, W, K, ctx, rank, world, pg, local, not args.no_solve, hbm, fp64_peak)
    sec_name = args.secondary if args.secondary is not None else (SECONDARY if world == 1 and args.workload == DEFAULT else "none")
    sec = None
    if sec_name != "none" and sec_name != args.workload:
        sec = measure_workload(sec_name, W, K, ctx, rank, world, pg, local, not args.no_solve, hbm, fp64_peak)
    if rank != 0:
        return
    cpu = None
    if world == 1:                                                  # host baseline: rank 0 at N = 1 only
        try:
            cpu = cpu_baseline(args.workload)
        except Exception as exc:                                    # never lose the GPU line over the CPU arm
            cpu = {"value": None, "unit": "s/iter", "cores": None, "kind": "port", "sample": None,
                   "error": "%s: %s" % (type(exc).__name__, exc)}
    cfg = config_of(args.workload, world)
    out = {
        "metric": "s_per_ipm_iteration", "value": main["dev_s"], "unit": "s/iter", "n_gpus": world,
        "steps": main["steps"], "warmup": W, "ms_per_step": main["dev_s"] * 1e3, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": main["e2e_s"], "unit": "s/iter", "h2d_bytes_per_step": main["h2d"], "d2h_bytes_per_step": main["d2h"]},
        "gpu_launches": main["launches"],
        "steps_requested": main["steps_requested"],
        "kernel_time_sum_s_per_iter": main["ksum_s"],
        "clocks": main["clocks"],
        "roofline": main["roofline"],
        "cpu_baseline": cpu,
        "schur_potrf": main["schur_potrf"],
        "chordal_ops_ms_per_step": main["ops"],
        "kernel_ms_per_step": main["kernel_ms_per_step"],
        "family_rooflines": main["family_rooflines"],
        "time_to_solve": main["time_to_solve"],
        "status": main["status"], "iterations": main["iterations"],
        "fp64_gemm_peak_tflops": fp64_peak,
    }
    if sec is not None:
        out["secondary"] = {
            "config": config_of(sec_name), "e2e": {"value": sec["e2e_s"], "unit": "s/iter",
                                                   "h2d_bytes_per_step": sec["h2d"], "d2h_bytes_per_step": sec["d2h"]},
            "value": sec["dev_s"], "unit": "s/iter", "gpu_launches": sec["launches"], "roofline": sec["roofline"],
            "schur_potrf": sec["schur_potrf"], "chordal_ops_ms_per_step": sec["ops"], "kernel_ms_per_step": sec["kernel_ms_per_step"],
            "family_rooflines": sec["family_rooflines"], "time_to_solve": sec["time_to_solve"], "clocks": sec["clocks"]}
    line = json.dumps(out)
    if json_fd is None:
        print(line)
    else:
        sys.stdout.flush()
        os.write(json_fd, (line + "\n").encode())


# --------------------------------------------------------------------------------------
def cpu_baseline(workload, budget_s=25.0):
    """The reference's algorithm (oracle/) on this box's host cores, ONE iteration of the M1 method at
    the workload's real starting iterate (the generator's strictly feasible X0):

      every phase that does not loop over the constraints runs in full and is timed as it runs
      (completion of X, llt, Hessian factor, one KKT solve with its Hessian / Amap / Aadj / potrs
      calls, one completion probe and one Cholesky probe of the line search, dpotrf of an m x m matrix);
      the per-column loop of the Schur assembly (solvers.py:479-497) runs on a stratified sample of
      columns (every (m/S)-th column of the dense and of the sparse group) and is scaled by m/S.

    The phases are combined with the operation counts of a regular (non-centering) M1 iteration
    (SURVEY.md 3.1): 1 assembly + 1 potrf, 9 KKT solves (3 Newton systems x (1 + 2 refinement
    rounds)), 6 KKT residuals, 8 + 8 line-search probes, 2 completions, 1 Cholesky, 1 llt."""
    os.environ["SMCP_B200_NO_NATIVE_HOST"] = "1"      # pure-NumPy symbolic analysis: no repo .so in this arm
    from smcp_b200 import solvers
    from oracle.backend import OracleBackend, _scm_column2
    from oracle import supernodal as sn
    from oracle import ref as oref
    from oracle import csn
    import scipy.linalg as sl
    c_threads = 0
    use_c = False
    try:
        from threadpoolctl import threadpool_info
        nthreads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        nthreads = 1
    w = WORKLOADS[workload]
    prev = solvers._backend_factory if hasattr(solvers, "_backend_factory") else None
    solvers.set_backend_factory(lambda symb: OracleBackend(symb, batch_columns=32))
    try:
        P = build_problem(workload)
        solvers.options["show_progress"] = False
        opt = solvers._read_options(P.n, feas=True)
        prob = solvers._Problem(P._A, P._b, opt, "chol", None)
        ob, symb, m, Ns = prob.ops, prob.symb, prob.m, prob.Ns
        md = m - Ns
        tm = {}

        def timed(key, fn):
            t0 = time.perf_counter()
            r = fn()
            tm[key] = time.perf_counter() - t0
            return r

        X = prob.from_original(P._X0).buf
        L = X.copy()
        timed("completion", lambda: sn.completion(symb, L))
        Sh = L.copy()
        timed("llt", lambda: sn.llt(symb, Sh))
        hf = timed("hessian_factor", lambda: ob.hessian_factor(L, X))
        Lc = Sh.copy()
        timed("cholesky", lambda: sn.cholesky(symb, Lc))
        # ---- Schur assembly, sampled columns --------------------------------------------
        Av = ob.Av
        t_dense = t_sparse = 0.0
        nd = ns_ = 0
        if md:
            S = max(1, min(md, 48))
            cols = np.unique(np.linspace(0, md - 1, S).astype(int))
            t0 = time.perf_counter()
            U = np.zeros((len(cols), symb.nblk))
            for q, j in enumerate(cols):
                c0, c1 = Av.indptr[j], Av.indptr[j + 1]
                U[q, symb.vec2blk[Av.indices[c0:c1]]] = Av.data[c0:c1]
            sn.hessian(hf, U)
            for q, j in enumerate(cols):
                at = U[q, symb.vec2blk] * ob.halfdiag
                # base.gemv(Av[:, j:m], ..., trans='T') incl. the slice copy of the reference (solvers.py:486)
                _ = Av[:, j:].T @ at
            t_dense = (time.perf_counter() - t0) * md / len(cols)
            nd = len(cols)
        if Ns:
            Ip, Jp = symb.Ip, symb.Jp
            S = max(1, min(Ns, 24))
            cols = md + np.unique(np.linspace(0, Ns - 1, S).astype(int))
            use_ref = oref.available()
            use_c = False
            if csn.available():
                try:
                    csn._load()
                    use_c = True
                except Exception:
                    use_c = False          # e.g. no OpenMP runtime on this host: the NumPy trsm is used
            H = np.zeros((m, m), order="F") if not use_ref else None
            if use_ref:
                # persistent cvxopt-ABI objects like the reference's solver holds them (H, Av, Ip, Jp are
                # created once per solve, solvers.py:346-367, 547): only V and kkl are per-column objects
                refmod, _ = oref._load()
                Hm, Avm = oref.dmatrix(np.zeros((m, m))), oref.spmatrix(Av)
                Ipm, Jpm = oref.imatrix(Ip), oref.imatrix(Jp)
            t0 = time.perf_counter()
            for jj in cols:
                c0, c1 = Av.indptr[jj], Av.indptr[jj + 1]
                rows_j = Av.indices[c0:c1]
                Kj = np.unique(np.concatenate([Ip[rows_j], Jp[rows_j]]))
                V = np.zeros((symb.n, len(Kj)))
                V[symb.iperm[Kj], np.arange(len(Kj))] = 1.0
                if use_c:
                    # compiled restatement of chompack.trsm on all host cores (oracle/csn.c, pinned on sn.trsm)
                    c_threads = max(c_threads, csn.trsm(symb, hf.Lbuf, V, 'N'))
                    csn.trsm(symb, hf.Lbuf, V, 'T')
                else:
                    sn.trsm(symb, hf.Lbuf, V, 'N')
                    sn.trsm(symb, hf.Lbuf, V, 'T')
                kkl = np.zeros(symb.n, dtype=np.int64)
                kkl[Kj] = np.arange(len(Kj))
                if use_ref:
                    # the reference's own compiled SCMcolumn2 (misc.c:620-663); its V is indexed by vertex
                    refmod.SCMcolumn2(Hm, Avm, oref.dmatrix(V[symb.iperm, :]), Ipm, Jpm, oref.imatrix(kkl), int(jj))
                else:
                    _scm_column2(H, Av, V, symb.iperm, Ip, Jp, kkl, int(jj))
            t_sparse = (time.perf_counter() - t0) * Ns / len(cols)
            ns_ = len(cols)
            H = Hm = None
        tm["assembly"] = t_dense + t_sparse
        # ---- dense factorisation and one KKT solve -----------------------------------------
        rng = np.random.default_rng(0)
        G = rng.standard_normal((m, min(m, 64)))
        Hm = G @ G.T + m * np.eye(m)
        ob.Hf = timed("potrf", lambda: sl.cholesky(Hm, lower=True, check_finite=False, overwrite_a=True))
        bx = Sh.copy()
        by = rng.standard_normal(m)

        def kkt_solve():
            r1 = bx.copy()
            sn.hessian(hf, r1.reshape(1, -1))
            y = ob.schur_solve(by + ob.Amap(r1))
            x = ob.Aadj(y) - bx
            sn.hessian(hf, x.reshape(1, -1))
            return x, y
        x, y = timed("kkt_solve", kkt_solve)

        def kkt_res():
            r = x.copy()
            sn.hessian_inv(hf, r.reshape(1, -1))
            return r + ob.Aadj(y) - bx, ob.Amap(x) - by
        timed("kkt_res", kkt_res)
        counts = {"assembly": 1, "potrf": 1, "kkt_solve": 9, "kkt_res": 6, "completion": 2 + 8, "cholesky": 1 + 8,
                  "llt": 1, "hessian_factor": 1}
        per_iter = sum(tm[k] * c for k, c in counts.items())
    finally:
        solvers.set_backend_factory(prev)
        os.environ.pop("SMCP_B200_NO_NATIVE_HOST", None)
    return {"value": per_iter, "unit": "s/iter", "cores": int(max(nthreads, c_threads)), "kind": "port",
            "phase_seconds": {k: round(v, 6) for k, v in tm.items()}, "phase_counts": counts,
            "sample": ("one M1 iteration at the generator's starting iterate X0: every phase run in full and timed "
                       "(%s), the per-column Schur loop on %d of %d dense and %d of %d sparse columns (stratified, scaled "
                       "to m; sparse columns: 2 chordal trsm (%s) + %s), combined with the op counts of a regular M1 "
                       "iteration %s" % (", ".join("%s %.3fs" % (k, v) for k, v in tm.items() if k != "assembly"),
                                         nd, md, ns_, Ns,
                                         ("compiled oracle/csn.c, %d threads" % c_threads) if (Ns and use_c) else "NumPy port",
                                         "the reference's compiled misc.SCMcolumn2 (oracle/_ref)" if (Ns and oref.available())
                                         else "the NumPy restatement of misc.SCMcolumn2", json.dumps(counts)))}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    vals = []
    cpu = None
    for _ in range(max(1, min(args.steps, 2))):
        cpu = cpu_baseline(args.workload)
        vals.append(cpu["value"])
    v = float(np.mean(vals))
    cpu["value"] = v
    cfg = config_of(args.workload, world)
    out = {"impl": "reference", "metric": "s_per_ipm_iteration", "value": v, "unit": "s/iter",
           "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": v * 1e3,
           "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": cfg,
           "cpu_baseline": cpu,
           "e2e": {"value": v, "unit": "s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "note": "cvxopt/chompack are not installable here; this is the oracle port of the reference's "
                   "CPU path (oracle/), each step one sampled iteration as described in cpu_baseline.sample"}
    print(json.dumps(out))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
