#!/usr/bin/env python
"""bench.py -- BSR sampling hot path on B200: MH proposals scored/s (+ tree-node evals/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c1]

A *step* is `sweeps_per_step` sweeps (each sweep = K newProp calls per chain, codes/bsr_class.py:179) of every
chain over the synthetic data set.  Default workload = BASELINE.json configs[1]: K=3, 4096
chains, n=1000 rows, d=2, and the default --steps 10 x 500 sweeps = the 5000 iterations the config names.
For N>1 (torchrun, one rank per GPU) every rank runs its own 4096 chains (global chain ids offset by rank,
no data-path collective): weak scaling.

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` goes through the
C-ABI with host buffers (H2D of X,y and D2H of the results inside the timed region); `roofline` describes the
dominant kernel (k_eval); `cpu_baseline` is the oracle port timed on the host cores on a bounded sample.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (K, chains_per_gpu, n, d, sweeps_per_step, target)
    "c1": dict(K=3, chains=50, n=100, d=2, sweeps=100, target="f1", seed=1001),
    "c2": dict(K=3, chains=4096, n=1000, d=2, sweeps=500, target="sim", seed=2001),
    "c3": dict(K=10, chains=16384, n=10000, d=8, sweeps=16, target="mix8", seed=3001),
    "c4": dict(K=5, chains=8192, n=5000, d=8, sweeps=32, target="mix8", seed=4001),
    # large-data fit (BASELINE configs[4]): rows sharded over the ranks, 12.5 M rows per GPU (1e8 at 8 GPUs), the same 256
    # chains on every rank; data generated on the device; k_wresolve reads the ranks' partial sums over NVLink peer memory
    "c5": dict(K=5, chains=256, n=12500000, d=8, sweeps=13, target="mix8", seed=5001, row_sharded=True),
}


def make_data(w):
    rng = np.random.default_rng(w["seed"])
    X = rng.uniform(-3, 3, (w["n"], w["d"]))           # codes/simulations.py:66-67
    if w["target"] == "f1":
        y = 2.5 * X[:, 0] ** 4 - 1.3 * X[:, 0] ** 3 + 0.5 * X[:, 1] ** 2 - 1.7 * X[:, 1]
    elif w["target"] == "sim":                         # codes/simulations.py:71
        y = 1.35 * X[:, 0] * X[:, 1] + 5.5 * np.sin((X[:, 0] - 1) * (X[:, 1] - 1))
    else:
        y = np.exp(0.5 * X[:, 0]) + 2.0 * np.cos(X[:, 1]) + 0.3 * X[:, 7] * X[:, 2] + np.sin(X[:, 3] * X[:, 4]) + rng.normal(0, 0.1, w["n"])
    return X, y


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------
def _oracle_worker(args):
    """Continue one oracle chain for `sweeps` sweeps (fixed-sweep mode, like the GPU bench)."""
    X, y, K, d, seed, sweeps, init = args
    from oracle import bsr_oracle as O
    cfg = O.Config(n_feature=d)
    dr = O.GeneratorDraws(seed)
    t0 = time.perf_counter()
    r = O.run_chain(X, y, K, cfg, dr, val=0, max_sweeps=sweeps, fixed_sweeps=True, init=init)
    dt = time.perf_counter() - t0
    nxt = dict(sigma=r.sigma, trees=r.final_state, sigma_a=r.sigma_a, sigma_b=r.sigma_b)
    return r.n_proposals, r.node_evals_ref, dt, nxt


class OraclePool:
    def __init__(self, w, procs):
        import multiprocessing as mp
        self.w, self.procs = w, procs
        self.X, self.y = make_data(w)
        self.pool = mp.get_context("fork").Pool(procs)
        self.state = [None] * procs
        self.round = 0

    def step(self, sweeps):
        w = self.w
        args = [(self.X, self.y, w["K"], w["d"], 1000 * self.round + i, sweeps, self.state[i]) for i in range(self.procs)]
        t0 = time.perf_counter()
        res = self.pool.map(_oracle_worker, args)
        wall = time.perf_counter() - t0
        self.round += 1
        self.state = [r[3] for r in res]
        return sum(r[0] for r in res), sum(r[1] for r in res), wall

    def close(self):
        self.pool.terminate()


def cpu_baseline(w, budget_s=12.0):
    procs = os.cpu_count() or 1
    n_full = w["n"]
    if w.get("row_sharded"):
        # the CPU sampler cannot hold 1e7..1e8 rows per chain in reasonable time: time it at n = 1e5 rows (its cost per
        # proposal is linear in n) and label the full-size figure as extrapolated
        w = dict(w, n=100000)
    pool = OraclePool(w, procs)
    try:
        sweeps = max(1, int(40 * 1000 / w["n"]))
        pool.step(max(1, sweeps // 4))               # warm-up (imports, first-touch)
        props = evals = 0
        wall = 0.0
        while wall < budget_s:
            p, e, dt = pool.step(sweeps)
            props += p; evals += e; wall += dt
        extra = {}
        if w["n"] != n_full:
            extra = dict(extrapolated_to_rows=n_full, extrapolated_value=props / wall * w["n"] / n_full,
                         note="timed at n=%d rows, scaled linearly in n to %d rows per GPU" % (w["n"], n_full))
        return dict(value=props / wall, unit="proposals/s", cores=procs, kind="port",
                    node_evals_ref_per_s=evals / wall, **extra,
                    sample="%d oracle chains (1 per core) x %d sweeps x %d proposals on the same X,y (n=%d, d=%d, K=%d), %.1f s wall"
                           % (procs, int(round(props / procs / w["K"])), w["K"], w["n"], w["d"], w["K"], wall))
    finally:
        pool.close()


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    if w.get("row_sharded"):
        w = dict(w, n=100000)                         # bounded sample of the large-data workload (cost is linear in n)
    pool = OraclePool(w, procs)
    sweeps = max(1, int(100 * 1000 / w["n"]))         # bounded sample per step
    for _ in range(args.warmup):
        pool.step(sweeps)
    props = evals = 0
    wall = 0.0
    for _ in range(args.steps):
        p, e, dt = pool.step(sweeps)
        props += p; evals += e; wall += dt
    pool.close()
    v = props / wall
    line = dict(metric="mh_proposals_per_sec", value=v, unit="proposals/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * wall / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                data="synthetic", impl="reference",
                config=dict(workload=args.workload, K=w["K"], n_rows=w["n"], d=w["d"], chains=procs, sweeps_per_step=sweeps,
                            note="oracle port of the reference sampler (numpy, one chain per host core); the Python reference itself cannot travel to the GPU box"),
                node_evals_ref_per_sec=evals / wall,
                cpu_baseline=dict(value=v, unit="proposals/s", cores=procs, kind="port",
                                  sample="%d chains x %d sweeps per step, %d steps" % (procs, sweeps, args.steps)),
                e2e=dict(value=v, unit="proposals/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    emit(line)


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=[])
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons))


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import __graft_entry__ as g
    g.build()
    from mcmc_symreg_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    K, C, n, d, S = w["K"], args.chains or w["chains"], args.rows or w["n"], w["d"], args.sweeps_per_step or w["sweeps"]
    ops, weights = list(range(1, 11)), [0.1] * 10
    row_sharded = bool(w.get("row_sharded"))
    run = None
    if row_sharded:
        # rows [rank * n, (rank + 1) * n) of a global data set of world * n rows, generated on the device (fp32, column-major)
        from mcmc_symreg_b200 import parallel
        ld = (n + 3) // 4 * 4
        gen = torch.Generator(device="cuda").manual_seed(w["seed"] + rank)
        Xd = torch.rand((d, ld), generator=gen, device="cuda", dtype=torch.float32) * 6 - 3
        yd = (torch.exp(0.5 * Xd[0]) + 2.0 * torch.cos(Xd[1]) + 0.3 * Xd[7] * Xd[2] + torch.sin(Xd[3] * Xd[4])
              + 0.1 * torch.randn(ld, generator=gen, device="cuda", dtype=torch.float32)).contiguous()
        X, y = None, None
        eng = capi.Engine(K, C, ops, weights, beta=-1.0, val=0, plateau_rule=False, precision=args.precision, device=local,
                          chain_offset=0, row_sharded=world > 1)
        eng.set_data_device(Xd.data_ptr(), yd.data_ptr(), n, d, ld, n_total=n * world)
        if world > 1:
            rs = parallel.RowShardedEngine(eng, n * world)
            rs.init_chains(w["seed"])
            assert rs.enable_peer_windows()
            run = lambda sweeps, stream=None: rs.run(sweeps)
        else:
            eng.init_chains(w["seed"])
    else:
        X, y = make_data(w)
        eng = capi.Engine(K, C, ops, weights, beta=-1.0, val=0, plateau_rule=False, precision=args.precision, device=local,
                          chain_offset=rank * C)
        eng.set_data(X, y)
        eng.init_chains(w["seed"])
    if run is None:
        run = eng.run
    eng.set_launch_geometry(0, args.groups)
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    for _ in range(args.warmup):
        run(S, stream)
    barrier()
    c0 = eng.get_stats()["counters"].sum(axis=0)
    sampler = ClockSampler(local)
    sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    l0 = eng.launch_count()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                      # L2 flush between timed iterations (outside the event pair)
        evs[i][0].record()
        run(S, stream)
        evs[i][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    n_launches = eng.launch_count() - l0
    clocks = sampler.result()
    ms_steps = [a.elapsed_time(b) for a, b in evs]
    ms_total = float(sum(ms_steps))
    c1 = eng.get_stats()["counters"].sum(axis=0)
    dc = c1 - c0
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    tot = torch.tensor([float(dc[0]), float(dc[5]), float(dc[6]), float(dc[1]), float(dc[2]), float(dc[4]), float(dc[3])], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot)
    ms_max = float(t.item())
    props, ev_ref, ev_exec, accepts, rank_rej, fp64_sw, cap_rej = [float(v) for v in tot.tolist()]
    if row_sharded:      # every rank runs the same chains on its own rows: proposals are not additive, node evaluations are
        props, accepts, rank_rej, fp64_sw, cap_rej = [v / world for v in (props, accepts, rank_rej, fp64_sw, cap_rej)]
    value = props / (ms_max * 1e-3)

    # ---- per-stage / per-kernel device time (CUDA events on the run's stream, separate short run) + roofline ----
    eng.set_profiling(True)
    prof_sweeps = min(S, 128)          # 128 K proposals per chain = 2 K full 64-slot windows
    run(prof_sweeps, stream)
    torch.cuda.synchronize()
    prof = eng.get_profile()
    eng.set_profiling(False)
    tok, pa, pb, nn = eng.get_trees(current=True)
    mean_nodes = float(nn.mean())
    iters = max(1, prof["iterations"])
    stage_ms = dict((k, v / iters) for k, v in prof["ms"].items())          # per window iteration
    k_ms = prof["kernels_ms"]["eval_main"] / iters                           # k_weval alone
    total_ms = sum(stage_ms.values())
    W = max(1, min(64, int(os.environ.get("BSR_WINDOW", "64"))))      # bsr_run's window (library default 64)
    # Dominant kernel: k_weval.  One launch interprets, for every chain, its K live trees and the W proposals of the
    # window on all n rows and reduces K + 4 fp64 sums per proposal.  Algorithmic HBM bytes of a
    # launch: X and y once (shared by every chain, fp32 X + fp64 y), per chain the tokens of K + W trees (20 B per
    # node) and the W records it writes ((K + 4) doubles each).  The kernel is issue-bound, not HBM-bound: the
    # compute figures below and the ncu pipe utilisation in profiles/ are the relevant evidence.
    alg_bytes = 4.0 * d * n + 8.0 * n + C * ((K + W) * mean_nodes * 20.0 + W * (K + 4) * 8.0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    ncu = {}
    try:   # per-launch figures of the same kernel from the committed ncu --set full capture
        ncu = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload) or {}
    except Exception:
        pass
    # executed work of one k_weval launch in steady state, from the device counter of executed node evaluations of the timed
    # region (a tree that repeats an earlier slot of its window is interpreted once and not counted; out-of-range columns
    # count twice): per consumed proposal, times the C * W proposals of a full window (the live columns are in the counter).
    # The fp64 FMAs (K + 3 per row of an interpreted proposal) are scaled by the same interpreted share -- on the low side,
    # the repeated trees being the small ones.
    per_rank_exec = ev_exec / world
    per_rank_props = props if row_sharded else props / world
    node_row_evals = per_rank_exec / max(per_rank_props, 1.0) * C * W
    interpreted_share = min(1.0, node_row_evals / (C * n * (K + W) * mean_nodes))
    fp64_fma = C * n * W * (K + 3) * interpreted_share
    roofline = dict(bound="hbm", achieved=achieved, peak=hbm_peak, unit="GB/s", frac=achieved / hbm_peak,
                    traffic=ncu.get("dram_bytes_per_launch"),
                    kernel="k_weval<float,%d> (K live + %d proposed trees per chain: interpreter + fused Gram sums)" % (K, W),
                    ms_per_launch=k_ms, algorithmic_bytes_per_launch=alg_bytes, peak_source="measured" if peaks else "fallback",
                    bound_actual="issue",
                    note="the data (%.0f KB) is shared by every chain and L1/L2-resident: this path is instruction-issue bound "
                         "(SURVEY 8d), so frac against HBM is small by construction; see `issue` (ncu) and `compute` (live)" % ((4 * d + 8) * n / 1e3),
                    issue=dict(source="ncu --set full, profiles/ (static, not measured in this run)",
                               inst_per_cycle_per_sm=ncu.get("inst_per_cycle_per_sm"), peak_inst_per_cycle_per_sm=4.0,
                               frac=(ncu.get("inst_per_cycle_per_sm") / 4.0) if ncu.get("inst_per_cycle_per_sm") else None,
                               pipes_pct=ncu.get("pipes_pct")),
                    stage_ms_per_window=stage_ms, kernel_ms=dict(k_weval=k_ms, k_weval_fix=prof["kernels_ms"]["eval_second"] / iters),
                    share_of_window=dict((k, v / total_ms) for k, v in stage_ms.items()), windows_profiled=iters,
                    compute=dict(node_row_evals_per_s_in_k_weval=node_row_evals / (k_ms * 1e-3),
                                 fp64_fma_per_s_in_k_weval=fp64_fma / (k_ms * 1e-3), interpreted_share=interpreted_share,
                                 source="BSR_CNT_NODE_EVALS_EXEC of the timed region / k_weval time per window"))

    # ---- end to end through the C-ABI with host buffers ----
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        if X is not None:                          # (the row-sharded workload generates its shard on the device)
            eng.set_data(X, y)                     # H2D of this step's inputs (host float64 row-major, as BSR.fit receives them)
        run(S, stream)
        st = eng.get_stats()                       # D2H of the step's results
        tr = eng.get_trees(current=False, reuse=True)   # lands in the engine's page-locked result buffers
        return sum(v.nbytes for v in st.values()) + sum(v.nbytes for v in tr)

    for _ in range(2):                             # untimed: first-use allocations (page-locked result buffers, staging)
        d2h = e2e_step()
    barrier()
    t0 = time.perf_counter()
    e2e_step_ms = []
    for _ in range(e2e_steps):
        t1 = time.perf_counter()
        d2h = e2e_step()
        e2e_step_ms.append(1e3 * (time.perf_counter() - t1))
    barrier()
    e2e_wall = time.perf_counter() - t0
    t = torch.tensor([e2e_wall], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = (1 if row_sharded else world) * C * K * S * e2e_steps / float(t.item())
    eng.close()

    if rank == 0:
        line = dict(metric="mh_proposals_per_sec", value=value, unit="proposals/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms_max / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f32" if args.precision == "fp32" else "f64", data="synthetic",
                    config=dict(workload=args.workload, K=K, chains_per_gpu=C, n_rows=n * (world if row_sharded else 1), d=d, sweeps_per_step=S,
                                proposals_per_step=(1 if row_sharded else world) * C * K * S, l2_flush_between_steps=True, target=w["target"],
                                precision=args.precision, groups=args.groups, window=max(1, min(64, int(os.environ.get("BSR_WINDOW", "64")))), rng="philox4x32-10", parallelism=("rows x%d (peer-memory windows)" if row_sharded else "chains x%d") % world),
                    node_evals_ref_per_sec=ev_ref / (ms_max * 1e-3), node_evals_exec_per_sec=ev_exec / (ms_max * 1e-3),
                    accept_rate=accepts / max(props, 1), rank_reject_rate=rank_rej / max(props, 1), fp64_sweeps=fp64_sw,
                    capacity_rejects=cap_rej, mean_nodes_per_tree=mean_nodes,
                    gpu_launches=int(n_launches), wall_s=t_wall, clocks=clocks,
                    e2e=dict(value=e2e_value, unit="proposals/s", h2d_bytes_per_step=int(X.nbytes + y.nbytes) if X is not None else 0, d2h_bytes_per_step=int(d2h),
                             steps=e2e_steps, ms_per_step=1e3 * float(t.item()) / e2e_steps, ms_per_step_median=float(np.median(e2e_step_ms)), note="bsr_set_data_host (pageable host X, y) + bsr_run + bsr_get_stats + bsr_get_trees (page-locked result arrays) per step, wall clock"),
                    roofline=roofline)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(w)
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


_JSON_FD = None


def emit(line):
    """The one JSON line of the contract, on the process's real stdout."""
    out = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(out.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, out)


def main():
    # stdout carries exactly one JSON line: whatever libraries print to fd 1 meanwhile (NCCL's version banner when
    # NCCL_DEBUG is set on the box, worker processes) goes to stderr
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=0)
    ap.add_argument("--rows", type=int, default=0, help="rows (per GPU for the row-sharded workload)")
    ap.add_argument("--sweeps-per-step", type=int, default=0)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "fp64"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--groups", type=int, default=0, help="chain groups pipelined on separate streams (0 = library default)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
