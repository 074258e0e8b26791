#!/usr/bin/env python
"""bench.py -- LM bundle-adjustment throughput on synthetic BAL-shaped problems.

A "step" is one LM trial step (one trip of bundle_euclid.m:139-241: residuals + finite-
difference Jacobians + U/V/W, damping, V*^-1, Schur complement / PCG solve, back-substitution,
new residual, accept/reject) over one synthetic problem.  Under N ranks that ONE problem is sharded by
point (shard.shard_points: contiguous point ranges balanced by observation count, every rank holds
all cameras) -- strong scaling; the per-camera sums travel through NCCL all-reduce and the PCG
vector through NVLink peer-memory mailboxes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config venice] [--impl reference]

Prints ONE JSON line (rank 0).  `value` = observations processed per second through whole LM
trial steps, all ranks together, inputs resident in HBM; `e2e` = the same through the
host-buffer C-ABI call (H2D of a, b, observations and D2H of a_new, b_new inside the timing).
`--impl reference` times the CPU arm (oracle/: the reference's algorithm as a sparse OpenMP
port, since the reference's dense n x m arrays cannot hold these sizes).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "LM trial-step throughput (Jacobian + Schur + solve + update), observations/s"
UNIT = "obs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="venice")
    ap.add_argument("--scale", type=float, default=1.0, help="scale points/observations of the problem")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--solver", default="auto", choices=["auto", "chol", "pcg", "pcgx"])
    ap.add_argument("--pcg-rtol", type=float, default=1e-8)
    ap.add_argument("--rtable", default="host", choices=["host", "device"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--model", default="euclid", choices=["euclid", "projective"],
                    help="euclid: bundle_euclid.m with 'fix_calibration' (num_a = 6, the headline); projective: bundle_projective.m (num_a = 12)")
    ap.add_argument("--banded", action="store_true", help="tracks are contiguous camera windows (no loop-closure cameras): S is banded")
    ap.add_argument("--no-verify", action="store_true", help="multi-GPU: skip the check of the first N-rank step against rank 0 alone")
    ap.add_argument("--autotune", type=int, default=0,
                    help="opts.pcg_autotune: re-weight the matvec cut by the measured per-SM rate over the first N solves (0 = off)")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: all-reduce the PCG vector with NCCL instead of peer-memory mailboxes")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([c.strip() for c in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_global(args):
    """The ONE problem of this run (every rank generates the same one: the generator is deterministic)."""
    from bundleadjustmentmatlab_b200 import synth
    P = synth.make_config(args.config, seed=args.seed, scale=args.scale, banded=getattr(args, "banded", False))
    if getattr(args, "model", "euclid") == "projective":
        # a = vec(P_j), P_j = K_j [R(w_j) T_j]  (bundle_projective.m:69-72)
        a = np.zeros((P.m, 12))
        R = synth.rodrigues(P.w)
        for j in range(P.m):
            Kj = np.array([[P.K[0, j], 0, P.K[2, j]], [0, P.K[1, j], P.K[3, j]], [0, 0, 1]])
            a[j] = (Kj @ np.hstack([R[j], P.Te[:, j:j + 1]])).reshape(12, order="F")
    else:
        a = np.ascontiguousarray(np.vstack([P.w, P.Te]).T)
    b = np.ascontiguousarray(P.Xe[:3].T)
    return P, a, b


def make_local(P, b, rank, world):
    """This rank's point shard of the problem: all cameras, a contiguous point range."""
    from types import SimpleNamespace
    from bundleadjustmentmatlab_b200 import shard
    if world == 1:
        return SimpleNamespace(m=P.m, n=P.n, nobs=P.nobs, K=P.K, obs_xy=P.obs_xy, obs_pt=P.obs_pt, obs_cam=P.obs_cam, b=b, lo=0, hi=P.n)
    xy, pt, cam, bl, (lo, hi) = shard.shard_points(P.obs_xy, P.obs_pt, P.obs_cam, b, rank, world)
    return SimpleNamespace(m=P.m, n=hi - lo, nobs=int(pt.shape[0]), K=P.K, obs_xy=xy, obs_pt=pt, obs_cam=cam, b=bl, lo=lo, hi=hi)


def workload_name(args, P, world):
    return (f"{args.config}-shaped synthetic BA: {P.m} cameras, {P.n} points, {P.nobs} observations (ONE problem); "
            f"{world} rank(s), point-sharded" + ("; banded visibility (no loop-closure cameras)" if getattr(args, "banded", False) else ""))


# ------------------------------------------------------------------------------------------
# CPU arm (oracle): bounded sample = whole LM trial steps on one rank's shard
# ------------------------------------------------------------------------------------------
def cpu_trial_steps(args, P, a, b, steps, warmup):
    from oracle import lm
    lib = lm.sparse_lib()
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: do not inherit that)
    lib.orc_set_num_threads(args.cpu_threads if args.cpu_threads > 0 else len(os.sched_getaffinity(0)))
    cores = int(lib.orc_num_threads())
    obs = lm.ObsList(P.m, P.n, P.obs_xy, P.obs_pt, P.obs_cam)
    aa, bb, lam = a.T.copy(), b.T.copy(), 1e-3
    times, iters, first = [], [], None
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        r = lm.trial_step_pcg(P.K, aa, bb, obs, lam, pcg_rtol=args.pcg_rtol, pcg_max_iter=1000)
        dt = time.perf_counter() - t0
        if first is None:
            first = {"old": float(r["old"]), "new": float(r["new"]), "pcg_iters": int(r["pcg_iters"])}
        if s >= warmup:
            times.append(dt); iters.append(r["pcg_iters"])
        if r["old"] - r["new"] > 0:
            rho = (r["old"] - r["new"]) / r["denom"]
            aa, bb = r["a_new"], r["b_new"]
            lam = lam * max(1.0 / 3.0, 1 - (2 * rho - 1) ** 3)
        else:
            lam = lam * 2
    return times, iters, cores, first


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.model == "projective":
        print(json.dumps({"impl": "reference", "unavailable": "the CPU port (oracle_sparse.c) restates the Euclidean model only; "
                          "the projective reference runs dense through oracle/_ref at test sizes"}), flush=True)
        return
    P, a, b = make_global(args)
    times, iters, cores, _ = cpu_trial_steps(args, P, a, b, args.steps, args.warmup)
    total = sum(times)
    val = P.nobs * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, P, 1), "solver": "pcg (block-Jacobi, implicit Schur)",
                   "pcg_rtol": args.pcg_rtol, "pcg_iters_mean": float(np.mean(iters))},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{len(times)} whole LM trial steps of the whole problem ({P.nobs} observations), "
                                   "oracle/oracle_sparse.c orc_trial_step_pcg (OpenMP); the reference's dense n x m "
                                   "arrays cannot hold this size"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lm_iters_per_sec": len(times) / total,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def algorithmic_bytes(name, L, na=6, pcg_iters=0.0, world=1):
    """ALGORITHMIC bytes per launch of each kernel group ON ONE RANK: L is the rank's shard
    (its observations and points, all cameras)."""
    no, n, m = L.nobs, L.n, L.m
    W = 3 * na * 8
    Np = (na * m + 31) // 32 * 32
    # S bytes one matvec of this rank streams: the kept tiles of its column block (vlg_ba_symv_bytes) when the library says so
    sbytes = getattr(L, "symv_bytes", 0) or 4 * Np * (Np + 32) / world
    table = {
        # read (u,v)+point id, write W and the per-observation V/eB terms (read back by the point pass); per point b;
        # per camera table/a/K in, partial U/eA out
        "stage1_cam": no * (16 + 4 + W) + n * 24 + m * (288 + 8 * na + 32 + 8 * (na * (na + 1) // 2 + na)),
        "stage1_pt": no * (16 + 4) + n * (24 + 72 + 24 + 8) + m * (72 + 8 * na + 32),
        "pcg_sweep_pt": no * (W + 4 + 4) + n * (72 + 24 + 4) + m * 8 * na,
        "pcg_sweep_cam": no * (W + 4) + n * 24 + m * 8 * na,
        "schur": no * (W + 4) + n * (72 + 24) + m * 8 * (na * (na + 1) // 2 + na),
        "stage3": no * (W + 4 + 4 + 16 + 4 + 4 + 8) + n * (72 + 24 + 24 + 24 + 24 + 8) + m * (72 + 8 * na + 32),
        "vinv": n * (72 + 72) + m * 2 * 8 * na * na,
        "w_copy": no * (2 * W + 4),
        # explicit-S PCG: this rank's column block of the lower triangle of S (32-column strips incl. the full diagonal
        # blocks; the blocks of the ranks have equal areas) + partial vectors
        "pcg_symv": sbytes + 8 * Np * 4,
        # the persistent PCG kernel: one launch = `pcg_iters` matvecs over this rank's block (+ the vectors)
        "pcg_persistent": pcg_iters * (sbytes + 8 * Np * 8),
    }
    return table.get(name)


# Jacobian+Schur = 320 B per observation + 216 B per point (24 in + 96 out in stage 1, 96 in in stage 2)
# + the dense S written when it is assembled
def jacobian_schur_bytes(L, na, assembled):
    Np = (na * L.m + 31) // 32 * 32
    return 320.0 * L.nobs + 216.0 * L.n + (8.0 * Np * Np if assembled else 0.0)


JS_GROUPS = ("stage1_cam", "stage1_pt", "w_copy", "vinv", "schur", "schur_blocks")
S1_GROUPS = ("stage1_cam", "stage1_pt", "w_copy")


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bundleadjustmentmatlab_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    P, a0, b0g = make_global(args)
    L = make_local(P, b0g, rank, world)
    b0 = L.b
    solver = {"auto": capi.SOLVER_AUTO, "chol": capi.SOLVER_CHOL, "pcg": capi.SOLVER_PCG, "pcgx": capi.SOLVER_PCG_EXPLICIT}[args.solver]
    rtable = capi.RTABLE_HOST_LIBM if args.rtable == "host" else capi.RTABLE_DEVICE
    proj = args.model == "projective"
    na = 12 if proj else 6
    mk = dict(num_variableK=0, solver=solver, pcg_rtol=args.pcg_rtol, rtable=rtable, device=local,
              model=capi.MODEL_PROJECTIVE if proj else capi.MODEL_EUCLID, pcg_autotune=args.autotune)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- multi-GPU correctness inside the bench: the first LM trial step of the WHOLE problem on rank 0 alone
    single = None
    if world > 1 and not args.no_verify:
        if rank == 0:
            c1 = capi.Context(**mk)
            c1.set_problem_sparse(None if proj else P.K.T, a0, b0g, P.obs_xy, P.obs_pt, P.obs_cam)
            single = c1.trial_step()
            c1.close()
        barrier()

    ctx = capi.Context(**mk)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.set_comm(rank, world, bytes(uid.cpu().numpy().tobytes()))
    ctx.set_problem_sparse(None if proj else L.K.T, a0, b0, L.obs_xy, L.obs_pt, L.obs_cam)
    L.symv_bytes = ctx.symv_bytes
    p2p = world > 1 and not args.no_p2p
    if p2p:
        # the per-iteration PCG vector goes through NVLink peer-memory mailboxes (CUDA IPC), not NCCL
        mine = torch.frombuffer(bytearray(ctx.p2p_export()), dtype=torch.uint8).cuda()
        allh = [torch.zeros(capi.P2P_HANDLE_BYTES, dtype=torch.uint8, device="cuda") for _ in range(world)]
        dist.all_gather(allh, mine)
        ctx.p2p_import(b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))

    # ---- device-resident arm.  The steps follow the reference's LM loop (bundle_euclid.m:120-249);
    # when its stop rule ends a solve, the next step starts a new solve from the initial estimate.
    def lm_step():
        if not ctx.lm_continue():
            ctx.lm_reset(a0, b0)
        return ctx.trial_step()

    # the first step from (a0, b0, lambda0) is the one the CPU oracle (N = 1) / the single-GPU context (N > 1) repeats
    first = lm_step()
    for _ in range(max(args.warmup - 1, 0)):
        lm_step()
    barrier()
    # timed region: K steps, CUDA events on the library's stream, no per-kernel instrumentation
    infos, step_ms = [], []
    ctx.reset_timers(False)
    l0 = ctx.kernel_launches
    with ClockSampler(local) as clk:
        ctx.timer_start()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            infos.append(lm_step())              # ends with the host synchronisation of the accept test
            step_ms.append(1e3 * (time.perf_counter() - t0))
        ms = ctx.timer_stop()
        barrier()
    launches = ctx.kernel_launches - l0
    # the same K steps again with a CUDA-event pair around every kernel group (feeds `roofline` and
    # `kernels`; the event records cost a few us per launch, so this pass is not the one `value` comes from)
    ctx.reset_timers(True)
    ctx.timer_start()
    infos2 = [lm_step() for _ in range(args.steps)]
    ms_instr = ctx.timer_stop()
    its2 = float(np.mean([i["pcg_iters"] for i in infos2])) if infos2 else 0.0
    barrier()
    groups = {}
    for g in ("stage1_cam", "stage1_pt", "w_copy", "vinv", "schur", "schur_blocks", "chol", "pcg_sweep_pt", "pcg_sweep_cam",
              "pcg_symv", "pcg_update", "pcg_persistent", "precond", "stage3"):
        avg, cnt = ctx.kernel_time(g)
        groups[g] = {"avg_ms": avg, "count": cnt, "total_ms": avg * cnt}
    ctx.reset_timers(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    nobs_total = float(P.nobs)
    value = nobs_total * args.steps / (ms_max * 1e-3)

    # ---- end-to-end arm: host buffers through the C ABI, LM control on the host
    e2e = None
    if not args.no_e2e:
        pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        hxy = pin(L.obs_xy.shape)
        # current and candidate parameters live in two pinned buffer pairs that swap roles on an accepted step
        # (MATLAB's `a = a_new; b = b_new` is a reference rebind, bundle_euclid.m:225-226, not a 24 MB copy)
        buf = {"cur": (pin(a0.shape), pin(b0.shape)), "new": (pin(a0.shape), pin(b0.shape))}
        buf["cur"][0][:] = a0; buf["cur"][1][:] = b0; hxy[:] = L.obs_xy
        st = {"lam": 1e-3, "nu": 2.0, "it": 1, "it2": 0, "err": []}
        nvis = nobs_total

        def one():
            # host-side LM control, as bundle_euclid.m drives its mex calls (:120-123, :218-241)
            e = st["err"]
            go = st["it"] < 20 and st["it2"] < 10 and (st["it"] < 3 or (
                e[st["it"] - 1] > 1e-20 and e[st["it"] - 2] - e[st["it"] - 1] > 1e-3 * e[st["it"] - 2]))
            if not go:
                buf["cur"][0][:] = a0; buf["cur"][1][:] = b0
                st.update(lam=1e-3, nu=2.0, it=1, it2=0, err=[])
            info = ctx.trial_step_host(buf["cur"][0], buf["cur"][1], hxy, st["lam"], buf["new"][0], buf["new"][1])
            if info["accepted"]:
                buf["cur"], buf["new"] = buf["new"], buf["cur"]
                if proj:
                    st["lam"] /= 10                      # bundle_projective.m:194
                else:
                    st["lam"] *= max(1.0 / 3.0, 1 - (2 * info["rho"] - 1) ** 3)
                    st["nu"] = 2.0
                e = st["err"]
                while len(e) < st["it"] + 1:
                    e.append(0.0)
                e[st["it"] - 1] = info["old_cost"] / nvis
                st["it"] += 1
                e[st["it"] - 1] = info["new_cost"] / nvis
                st["it2"] = 0
            else:
                if proj:
                    st["lam"] *= 10                      # bundle_projective.m:204
                else:
                    st["lam"] *= st["nu"]; st["nu"] *= 2
                st["it2"] += 1
        for _ in range(args.warmup):
            one()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            one()
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        # bytes moved by all ranks together: every rank sends a, its points and its observations, fetches a_new and its points
        hb = torch.tensor([float(buf["cur"][0].nbytes + buf["cur"][1].nbytes + hxy.nbytes), float(buf["new"][0].nbytes + buf["new"][1].nbytes + 64)],
                          dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(hb)
        e2e = {"value": nobs_total * args.steps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": int(hb[0].item()),
               "d2h_bytes_per_step": int(hb[1].item()), "ms_per_step": 1e3 * float(dt.item()) / args.steps}

    rc = 0
    if rank == 0:
        peaks = {}
        pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk_path):
            peaks = json.load(open(pk_path))
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        ab_of = lambda g: algorithmic_bytes(g, L, na, its2, world)
        dom = max((g for g in groups if ab_of(g) and groups[g]["count"] > 0),
                  key=lambda g: groups[g]["total_ms"], default=None)
        roof = None
        if dom:
            ab = ab_of(dom)
            ach = ab / (groups[dom]["avg_ms"] * 1e-3) / 1e9
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tp) and world == 1:
                tj = json.load(open(tp)).get(args.config, {})
                traffic = tj.get(dom)
                if traffic is None and tj.get(dom + "_per_iteration"):
                    traffic = tj[dom + "_per_iteration"] * its2       # per launch = per iteration x mean iterations of a launch
            roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "algorithmic_bytes_per_launch": ab, "avg_launch_ms": groups[dom]["avg_ms"],
                    "launches": groups[dom]["count"], "peak_source": peak_kind, "per_rank": True,
                    "share_of_step": groups[dom]["total_ms"] / (ms_instr * 1.0)}
        per_kernel = {}
        for g, v in groups.items():
            if v["count"]:
                ab = ab_of(g)
                per_kernel[g] = {"avg_ms": round(v["avg_ms"], 5), "count": v["count"],
                                 "GBps": (ab / (v["avg_ms"] * 1e-3) / 1e9) if ab else None}
        # BASELINE.json metric (2): observations / t(residuals + Jacobians + normal equations + V*^-1 + Schur complement) of ONE
        # fresh trial step -- no solve in it.  Stage-1 groups run only on fresh steps (a rejected step re-uses stage 1), the
        # Schur groups on every step.
        js_ms = 0.0
        for g in JS_GROUPS:
            v = groups[g]
            if v["count"]:
                js_ms += v["avg_ms"] if g in S1_GROUPS else v["total_ms"] / max(len(infos2), 1)
        assembled = infos[-1]["solver_used"] in (capi.SOLVER_CHOL, capi.SOLVER_PCG_EXPLICIT)
        js_bytes = jacobian_schur_bytes(L, na, assembled)
        cpu, parity = None, None
        if world > 1 and single is not None:
            parity = {"against": "the same first LM trial step of the whole problem on rank 0 alone (one GPU)",
                      "old_rel": abs(first["old_cost"] - single["old_cost"]) / single["old_cost"],
                      "new_rel": abs(first["new_cost"] - single["new_cost"]) / single["new_cost"],
                      "accept_equal": bool(first["accepted"] == single["accepted"]), "tol": 1e-9}
        if world == 1 and not args.no_cpu_baseline and not proj:     # the CPU port restates the Euclidean model only
            times, iters, cores, cfirst = cpu_trial_steps(args, P, a0, b0g, 1, 0)
            cpu = {"value": P.nobs * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"1 whole LM trial step of the same problem ({P.nobs} observations, {iters[0]} PCG iterations, "
                             f"{sum(times):.1f} s): oracle/oracle_sparse.c orc_trial_step_pcg with OpenMP"}
            parity = {"against": "oracle/oracle_sparse.c orc_trial_step_pcg: the same first LM trial step from (a0, b0, lambda0), same pcg_rtol",
                      "old_rel": abs(first["old_cost"] - cfirst["old"]) / cfirst["old"],
                      "new_rel": abs(first["new_cost"] - cfirst["new"]) / cfirst["new"],
                      "accept_equal": bool(first["accepted"] == (cfirst["old"] - cfirst["new"] > 0)), "tol": 1e-9}
        if parity is not None:
            parity["ok"] = bool(parity["old_rel"] <= 1e-12 and parity["new_rel"] <= parity["tol"] and parity["accept_equal"])
            if not parity["ok"]:
                rc = 3
        acc = [bool(i["accepted"]) for i in infos]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, P, world), "num_a": na, "model": args.model,
                       "solver": {capi.SOLVER_CHOL: "cholesky", capi.SOLVER_PCG: "pcg (cluster-Jacobi, implicit Schur)",
                                  capi.SOLVER_PCG_EXPLICIT: "pcg (two-partition cluster preconditioner, assembled S, symmetric lower-triangle matvec, one persistent kernel)"}[infos[-1]["solver_used"]],
                       "pcg_rtol": args.pcg_rtol, "rtable": args.rtable, "pcg_autotune": args.autotune,
                       "pcg_vector_allreduce": ("nvlink peer-memory mailboxes (k_p2p_allreduce)" if p2p else "nccl") if world > 1 else None,
                       "l2": "inputs larger than L2 (W alone is %.0f MB per rank)" % (L.nobs * 24 * na / 1e6)},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "lm_iters_per_sec": args.steps / (ms_max * 1e-3),
            "ms_per_step_instrumented": ms_instr / args.steps,
            "jacobian_schur_obs_per_sec": (nobs_total / (js_ms * 1e-3)) if js_ms > 0 else None,
            "jacobian_schur_ms": js_ms,
            "jacobian_schur_roofline_frac": (js_bytes / (js_ms * 1e-3) / 1e9 / peak) if js_ms > 0 else None,
            "jacobian_schur_note": "rank 0's kernels: stage 1 (both passes) + damping/V*^-1 + Schur diagonal pass + S assembly of one fresh trial step, "
                                   "no solve; bytes = 320 B/obs + 216 B/pt (+ 8 Np^2 when S is assembled) of rank 0's shard (SURVEY.md 8d)",
            "pcg_iters_mean": float(np.mean([i["pcg_iters"] for i in infos])),
            "accepted_steps": int(sum(acc)), "rejected_steps": int(len(acc) - sum(acc)),
            "ms_per_accepted_step": float(np.mean([t for t, k in zip(step_ms, acc) if k])) if any(acc) else None,
            "ms_per_rejected_step": float(np.mean([t for t, k in zip(step_ms, acc) if not k])) if not all(acc) else None,
            "pcg_iters": [int(i["pcg_iters"]) for i in infos],
            "lambdas": [float("%.3g" % i["lambda_used"]) for i in infos],
            "cost_first_last": [infos[0]["old_cost"], infos[-1]["new_cost"]],
            "kernels": per_kernel,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if rc:
        raise SystemExit(rc)


# ------------------------------------------------------------------------------------------
# --config incremental: the cadence of the reference's heaviest caller
# ------------------------------------------------------------------------------------------
def run_incremental(args):
    """incr_reconstruction.m:223-348 adds one camera at a time and calls bundle_euclid about three times per camera: a
    single-camera motion-only BA (estimate_camera.m:247-253, 'fix_structure'), a BA of all cameras so far (:262/266) and the
    same again after triangulation (:324/328).  test_incremental.m:19-27 runs it on 50 cameras / <= 200 tracked points.
    Here: the same sequence of calls on a synthetic scene of that size, ONE context re-used for every call (set_problem on
    a live context), timed end to end with host buffers (every call uploads its problem and downloads its result), against
    the CPU oracle making the same calls; plus the batched entry (all resections of the sequence side by side)."""
    import torch
    from bundleadjustmentmatlab_b200 import capi, synth
    from oracle import lm
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    P = synth.make_problem(50, 2000, 16000, seed=args.seed)
    a_all = np.ascontiguousarray(np.vstack([P.w, P.Te]).T); b_all = np.ascontiguousarray(P.Xe[:3].T)
    calls = []                      # (kind, camera range, options)
    for mc in range(3, P.m + 1):
        calls.append(("resect", (mc - 1, mc)))
        calls.append(("full", (0, mc)))
        calls.append(("full", (0, mc)))
    def sub(c0, c1):
        keep = (P.obs_cam >= c0) & (P.obs_cam < c1)
        return np.ascontiguousarray(P.obs_xy[keep]), P.obs_pt[keep], (P.obs_cam[keep] - c0).astype(np.int32)
    subs = {rng: sub(*rng) for _, rng in calls}
    ctx = {"resect": capi.Context(num_variableK=0, fix_structure=1), "full": capi.Context(num_variableK=0)}
    def gpu_call(kind, rng):
        xy, pt, cam = subs[rng]
        c = ctx[kind]
        c.set_problem_sparse(P.K.T[rng[0]:rng[1]], a_all[rng[0]:rng[1]], b_all, xy, pt, cam)
        return c.solve()[4]
    for kind, rng in calls[:6]:
        gpu_call(kind, rng)                                   # warm-up (module load, first allocations)
    torch.cuda.synchronize()
    l0 = sum(c.kernel_launches for c in ctx.values())
    t0 = time.perf_counter()
    errs = [gpu_call(kind, rng) for kind, rng in calls]
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    launches = sum(c.kernel_launches for c in ctx.values()) - l0
    # the batched entry: every resection of the sequence in one solve
    cb = capi.Context(num_variableK=0, fix_structure=1)
    cb.set_problem_sparse(P.K.T, a_all, b_all, P.obs_xy, P.obs_pt, P.obs_cam)
    cb.solve_cameras_independent()
    cb.set_problem_sparse(P.K.T, a_all, b_all, P.obs_xy, P.obs_pt, P.obs_cam)
    t0 = time.perf_counter()
    _, berrs, rounds = cb.solve_cameras_independent()
    t_batch = time.perf_counter() - t0
    # CPU oracle: the same calls (bounded sample: every `stride`-th call)
    lib = lm.sparse_lib()
    lib.orc_set_num_threads(args.cpu_threads if args.cpu_threads > 0 else len(os.sched_getaffinity(0)))
    stride = 6
    t_cpu, worst, nsample = 0.0, 0.0, 0
    for k in range(0, len(calls), stride):
        kind, rng = calls[k]
        xy, pt, cam = subs[rng]
        mc = rng[1] - rng[0]
        x = np.zeros((3, P.n, mc), order="F"); vis = np.zeros((P.n, mc), order="F")
        x[0, pt, cam] = xy[:, 0]; x[1, pt, cam] = xy[:, 1]; x[2] = 1.0; vis[pt, cam] = 1.0
        opts = ["fix_calibration", "visibility", vis] + (["fix_structure"] if kind == "resect" else [])
        t0 = time.perf_counter()
        ref = lm.bundle_euclid(P.K[:, rng[0]:rng[1]], P.Te[:, rng[0]:rng[1]], P.w[:, rng[0]:rng[1]], P.Xe, x, *opts, backend="sparse", record=False)
        t_cpu += time.perf_counter() - t0
        nsample += 1
        e = errs[k]
        worst = max(worst, abs(e[0] - ref.error_[0]) / ref.error_[0], abs(e[1] - ref.error_[1]) / ref.error_[1])
    for c in list(ctx.values()) + [cb]:
        c.close()
    val = len(calls) / t_gpu
    cpu_val = nsample / t_cpu
    line = {"metric": "incremental-reconstruction cadence, bundle_euclid calls/s (set-up + solve, host buffers in and out)", "value": val,
            "unit": "calls/s", "n_gpus": 1, "steps": len(calls), "warmup": 6, "ms_per_step": 1e3 * t_gpu / len(calls), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"incremental cadence (incr_reconstruction.m:223-348): cameras 3..{P.m} of a {P.m}-camera / {P.n}-point / {P.nobs}-observation scene, "
                                   "per added camera one single-camera 'fix_structure' BA + two BAs of all cameras so far, one context re-used"},
            "e2e": {"value": val, "unit": "calls/s", "h2d_bytes_per_step": int(np.mean([subs[r][0].nbytes + subs[r][1].nbytes + subs[r][2].nbytes for _, r in calls]) + b_all.nbytes),
                    "d2h_bytes_per_step": int(b_all.nbytes * 4 / 3)},
            "gpu_launches": int(launches), "roofline": None,
            "cpu_baseline": {"value": cpu_val, "unit": "calls/s", "cores": int(lib.orc_num_threads()), "kind": "port",
                             "sample": f"every {stride}th call of the same sequence ({nsample} calls, {t_cpu:.1f} s): oracle/lm.py bundle_euclid over oracle_sparse.c (dense pinv solve)"},
            "parity": {"against": "oracle, first accepted step of the sampled calls", "new_rel": worst, "tol": 1e-9, "ok": bool(worst <= 1e-9)},
            "batched_resections": {"cameras": P.m, "rounds": rounds, "ms": 1e3 * t_batch, "resections_per_sec": P.m / t_batch,
                                   "note": "vlg_ba_solve_cameras_independent: all single-camera 'fix_structure' problems of the sequence side by side"}}
    print(json.dumps(line), flush=True)


def main():
    # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    args = parse()
    if args.config == "incremental":
        if int(os.environ.get("RANK", "0")) == 0:
            run_incremental(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
