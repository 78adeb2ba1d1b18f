#!/usr/bin/env python
"""bench.py -- the VBHEM-H3M E-step on B200, BASELINE.json's metric and config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] -- N = 100 000 base HMMs x K = 64 centre
HMMs, S = 8 states both sides, 2-D full-covariance emissions, tau = 10, Nv = 100, centres from
the reference's `baseem` initialisation; synthetic, seeded (vbhem_b200/synth.py).  One "step" is
one fused E-step over every (base, centre) pair: pre-pass, L_elbo, responsibilities, statistics.
With N GPUs the bases are sharded (100 000 per GPU: weak scaling), centres replicated, and the
per-centre statistics all-reduced with NCCL; the `strong` key also reports the north-star
arrangement (100 000 bases in total, split over the N GPUs).

Prints ONE JSON line (rank 0).  `value` = pairs/s with the base set resident in HBM (timed on
the device with CUDA events, max over ranks); `e2e` = the same metric through the host-buffer
C ABI (vbhem_pack_base + vbhem_estep_hyp from pinned host memory, results read back);
`roofline` = the dominant kernel (first pass, k_lattice_bwd) against the fp64 DMMA peak measured
in the same process; `cpu_baseline` = the reference's own C E-step (oracle/_ref) on the box's
host cores, bounded sample.  `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "E-step (i,j) HMM-pairs/sec"
UNIT = "pairs/s"


def flops_per_pair(b, r, d, T):
    """fp64 operations of the REFERENCE algorithm per pair (FMA = 2, exp/log = 1)."""
    a5 = b * r * (4 * d * d + 3 * d + 2)
    a78 = (T - 1) * r * (8 * b * r + 2 * b * b + b) + (8 * b * r + 3 * b)
    a9 = 2 * b * r + (T - 1) * (2 * r * b * b + 3 * b * r * r + b * r)
    a10 = 2 * r * b * (1 + d + d * d)
    a12 = 14 + 2 * (2 * r + r * r + r * d + r * d * d)
    return dict(total=a5 + a78 + a9 + a10 + a12, first_pass=a5 + a78)


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_worker(nb, K, S, d, tau, seed):
    """One process: the reference's C E-step on nb bases x K centres.  Prints seconds."""
    import numpy as np
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import pack, synth
    from oracle import cref
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util import prepass_natural
    base, _ = synth.make_base_set(nb, S=S, d=d, C=K, seed=seed)
    hyp = synth.init_centres_baseem(base, K, S, seed=1001)
    cen = prepass_natural(hyp, "full")
    pb, pc = pack.pack_base(base), pack.pack_centres(cen)
    kind = "reference" if cref.have_ref() else "port"
    fn = cref.ref_estep_pairs if kind == "reference" else cref.port_estep_pairs
    fn(dict(pb, prior=pb["prior"][:4], Sb=pb["Sb"][:4], A=pb["A"][:4], mu=pb["mu"][:4], cov=pb["cov"][:4]), pc, tau)
    t0 = time.perf_counter()
    out = fn(pb, pc, tau)
    dt = time.perf_counter() - t0
    assert np.isfinite(out["L_elbo"]).all()
    print(json.dumps(dict(seconds=dt, pairs=nb * K, kind=kind)))


def run_cpu_baseline(nb, K, S, d, tau, cores):
    """`cores` concurrent worker processes (the reference's parfor, vbhem_h3m_c.m:28, is process-parallel)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", str(nb), "--centres", str(K), "--states", str(S),
           "--dim", str(d), "--tau", str(tau)]
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd + ["--seed", str(2000 + w)], stdout=subprocess.PIPE, text=True,
                              env=dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES=""))
             for w in range(cores)]
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    wall = time.perf_counter() - t0
    worst = max(o["seconds"] for o in outs)
    pairs = sum(o["pairs"] for o in outs)
    return dict(value=pairs / worst, unit=UNIT, cores=cores, kind=outs[0]["kind"],
                per_core=statistics.median(o["pairs"] / o["seconds"] for o in outs),
                sample=f"{cores} processes x ({nb} bases x {K} centres, S={S}, d={d} full, tau={tau}) = {pairs} pairs "
                       f"through the reference MEX gateway (per-pair part only; the MATLAB-side z_hat and "
                       f"statistics reduction are not included); slowest worker {worst:.2f} s, wall {wall:.1f} s")


def ncu_dram_bytes(path):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) from a profiles/*_ncu.json summary."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    try:
        with open(path) as f:
            nc = json.load(f)
        return sum(float(nc[k]["value"]) * scale[nc[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    except Exception:  # noqa: BLE001
        return None


# ----------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.armed = False      # only samples taken while armed (the timed region) are kept

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                     nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                     nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
            while not self.stop_flag:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                if self.armed:
                    self.samples.append(mhz)
                    for bit, name in names.items():
                        if r & bit:
                            self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"sampler_error:{type(e).__name__}")

    def summary(self):
        return dict(sm_mhz=statistics.median(self.samples) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


# ----------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bases", type=int, default=100_000, help="base HMMs per GPU")
    ap.add_argument("--centres", type=int, default=64)
    ap.add_argument("--states", type=int, default=8)
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--tau", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1001)
    ap.add_argument("--grid-em", type=int, default=0, metavar="MAX_ITER",
                    help="--workload grid: also run every candidate's EM to convergence (at most MAX_ITER passes)")
    ap.add_argument("--cpu-worker", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-bases", type=int, default=0, help="bases per CPU worker (0 = default sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--slices", type=int, default=6, help="slices of the streamed (e2e) E-step")
    ap.add_argument("--workload", default="estep", choices=["estep", "grid", "fb", "exprmt1"],
                    help="estep = BASELINE configs[2] (the metric's workload, default); grid = configs[3], the (K,S) "
                         "model-selection grid K=1..16 x S=1..8 over 20k bases in one batched call; fb = configs[4], "
                         "the VB-HMM forward-backward pass over 1M sequences of length 50; exprmt1 = configs[0], the "
                         "reference's synthetic experiment end to end (learn 40 HMMs, cluster over K=1..6 x S=1..5)")
    ap.add_argument("--trials", type=int, default=20, help="--workload exprmt1: initialisations per (K,S) candidate")
    a = ap.parse_args()
    K, S, d, tau = a.centres, a.states, a.dim, a.tau
    if a.workload == "fb":
        from bench_fb import main_fb
        return main_fb(a)
    if a.workload == "exprmt1":
        from bench_exprmt1 import main_exprmt1
        return main_exprmt1(a)
    if a.cpu_worker:
        return cpu_worker(a.cpu_worker, K, S, d, tau, a.seed)
    grid = a.workload == "grid"
    if grid:
        if a.bases == 100_000:
            a.bases = 20_000
        a.no_cpu, a.no_e2e = True, True

    # The contract is ONE JSON line on stdout: keep the real stdout aside and send everything else that
    # libraries print there (NCCL prints its version banner to stdout) to stderr.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W = max(a.warmup, 3) if a.impl == "b200" else a.warmup
    cores = os.cpu_count() or 1
    config = dict(workload=("configs[3]: (K,S) model-selection grid K=1..16 x S=1..8, every candidate's centres in ONE "
                            "batched fused E-step (128 soft-max groups, ragged state counts)") if grid else
                           "configs[2]: synthetic base HMMs x centre HMMs, full-covariance Gaussian emissions, "
                           "baseem-initialised centres, one fused E-step per step",
                  bases_per_gpu=a.bases, centres=K, states=S, dim=d, tau=tau, Nv=100,
                  pairs_per_step=a.bases * K * world, parallelism=f"bases sharded x{world}, centres replicated",
                  l2="flushed between timed steps (256 MiB write); working set 198 MB per GPU > L2 anyway")

    # ---- reference arm: CPU only, rank 0 only
    if a.impl == "reference":
        if rank != 0:
            return
        nb = a.cpu_bases or 256
        vals = []
        for s in range(a.warmup + a.steps):
            r = run_cpu_baseline(nb, K, S, d, tau, cores)
            if s >= a.warmup:
                vals.append(r)
        v = statistics.median(x["value"] for x in vals)
        ms = 1e3 * (vals[0]["cores"] * nb * K) / v
        line = dict(metric=METRIC, value=v, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=ms,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    impl="reference", config=dict(config, note="each step = a bounded sample of the workload"),
                    cpu_baseline=dict(value=v, unit=UNIT, cores=vals[0]["cores"], kind=vals[0]["kind"],
                                      sample=vals[0]["sample"]),
                    e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        emit(line)
        return

    # ---- CPU baseline first (before CUDA is initialised in this process), rank 0, N = 1 only
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cpu = run_cpu_baseline(a.cpu_bases or 3072, K, S, d, tau, cores)   # ~12 s of CPU work per core

    import numpy as np
    import torch
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import pack, synth
    from vbhem_b200._lib import check, dp, ip, lib
    from vbhem_b200.api import BaseSet

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = lib()

    # ---- workload: centres come from the seed-1001 base set on every rank; shard r uses seed 1001 + r
    base0, _ = synth.make_base_set(a.bases, S=S, d=d, C=K, seed=a.seed)
    grp, ngroups = None, 1
    if grid:
        hyp, grp, models = synth.model_selection_grid(base0, seed=a.seed)
        K, ngroups = int(hyp["Sr"].shape[0]), len(models)
        config.update(centres=K, models=ngroups, pairs_per_step=a.bases * K * world)
    else:
        hyp = synth.init_centres_baseem(base0, K, S, seed=a.seed)
    base = base0 if rank == 0 else synth.make_base_set(a.bases, S=S, d=d, C=K, seed=a.seed + rank)[0]
    pb = pack.pack_base(base)
    c = np.ascontiguousarray
    Sr = c(hyp["Sr"], dtype=np.int32)
    H = dict(alpha=c(hyp["alpha"]), eta=c(hyp["eta"]), eps=c(np.transpose(hyp["epsilon"], (0, 2, 1))),
             m=c(hyp["m"]), W=c(np.transpose(hyp["W"], (0, 1, 3, 2))), v=c(hyp["v"]), lam=c(hyp["lam"]))
    P = lambda x: x.ctypes.data_as(dp)
    ns = int(L.vbhem_stats_size(K, S, d, 1, ngroups))
    gp = None if grp is None else grp.ctypes.data_as(ip)
    stats = torch.zeros(ns, dtype=torch.float64, device="cuda")
    pre = np.zeros(K + S * K * (2 + S))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    peak = (C.c_double * 2)()
    check(L.vbhem_measure_fp64_peak(local, peak))

    def make_handle(packed):
        bs = BaseSet(packed, device=local, packed=True)
        bs.set_stream(torch.cuda.current_stream().cuda_stream)
        return bs

    def step(bs):
        check(L.vbhem_estep_hyp_dev(bs._h, K, Sr.ctypes.data_as(ip), S, gp, ngroups, tau, P(H["alpha"]), P(H["eta"]),
                                    P(H["eps"]), P(H["m"]), P(H["W"]), P(H["v"]), P(H["lam"]),
                                    C.c_void_p(stats.data_ptr()), P(pre)))
        if dist is not None:
            dist.all_reduce(stats)
        check(L.vbhem_finalize_stats_dev(bs._h, K, Sr.ctypes.data_as(ip), S, ngroups, C.c_void_p(stats.data_ptr())))

    def timed_run(bs, steps, warm, profile=False):
        """The timed region.  profile: the library records CUDA events between the phases of every E-step
        (vbhem_last_timings) -- they break the programmatic launch chain, so the headline is timed without."""
        check(L.vbhem_set_profile(bs._h, 1 if profile else 0))
        sampler = ClockSampler(local)
        sampler.start()
        for _ in range(warm):
            step(bs)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        sampler.armed = True
        tot, phases, launches = 0.0, np.zeros(8), 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            e0.record()
            step(bs)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
            if profile:
                t = (C.c_double * 8)()
                check(L.vbhem_last_timings(bs._h, t))
                phases += np.array(list(t))
            launches += bs.counters()["launches"] + 1
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            tt = torch.tensor([tot], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tot = float(tt.item())
        sampler.armed = False
        sampler.stop_flag = True
        sampler.join()
        return tot, phases / steps, launches, sampler.summary(), bs.counters()

    bs = make_handle(pb)
    tot_ms, _, launches, clocks, counters = timed_run(bs, a.steps, W)
    prof_ms, ph, _, _, _ = timed_run(bs, max(3, min(a.steps, 10)), 1, profile=True)    # phase breakdown, untimed
    pairs_step = a.bases * K * world
    value = pairs_step * a.steps / (tot_ms * 1e-3)
    if grid:   # algorithmic flops of the reference per pair depend on the centre's own state count
        per = [flops_per_pair(S, int(r), d, tau) for r in Sr]
        fl = dict(first_pass=sum(f["first_pass"] for f in per) / K, total=sum(f["total"] for f in per) / K)
    else:
        fl = flops_per_pair(S, S, d, tau)
    first_ms = ph[1]
    ach = a.bases * K * fl["first_pass"] / (first_ms * 1e-3) / 1e12
    alg_bytes = a.bases * (S + S * S + S * d + S * d * d) * 8 + a.bases * K * 8
    # DRAM traffic of the same kernel from the committed `ncu --set full` capture (same workload, per launch)
    traffic = None
    if (a.bases, K, S, d, tau) == (100_000, 64, 8, 2, 10) and not grid:
        traffic = ncu_dram_bytes(os.path.join(ROOT, "profiles", "r01c_k_lattice_bwd_ncu.json"))
    roofline = dict(bound="tensor", kernel="k_lattice_bwd (first pass: emission expectation + backward lattice + "
                                           "termination for every pair; fp64 DMMA m8n8k4)",
                    achieved=ach, peak=float(peak[0]), unit="TFLOP/s", frac=ach / float(peak[0]),
                    peak_source="fp64 DMMA chain measured in this process (vbhem_measure_fp64_peak); "
                                "MEASURED_PEAKS.json has no fp64 entry; DFMA chain: %.2f TFLOP/s" % float(peak[1]),
                    flops_per_pair=fl["first_pass"], kernel_ms=first_ms, traffic=traffic,
                    traffic_source="profiles/r01c_k_lattice_bwd_ncu.json (dram__bytes_read.sum + dram__bytes_write.sum, "
                                   "bytes per launch; null when the workload differs from the captured one)",
                    hbm=dict(algorithmic_bytes=alg_bytes, achieved_gbs=alg_bytes / (first_ms * 1e-3) / 1e9),
                    whole_step=dict(flops_per_pair=fl["total"],
                                    achieved=a.bases * K * fl["total"] / (tot_ms / a.steps * 1e-3) / 1e12,
                                    note="reference-algorithm flops of the complete E-step / step time; the second "
                                         "pass runs only on pairs with Z > 1e-8, as the reference's reduction does"))
    phases = dict(stage_ms=ph[0], first_pass_ms=ph[1], zhat_list_ms=ph[2], second_pass_ms=ph[3], reduce_ms=ph[4],
                  estep_ms=ph[5], note="separate pass with phase events recorded inside the E-step", active_pairs=counters["active_pairs"], fallback_pairs=counters["general_pairs"])

    # ---- full VBHEM iterations (E-step + lower bound + M-step + convergence test), N = 1 only
    vbhem_iter = None
    if world == 1 and not grid:
        from vbhem_b200 import h3m
        opt = dict(synth.DEFAULT_OPT, tau=tau, max_iter=6, minDiff=0.0)
        check(L.vbhem_set_profile(bs._h, 0))
        # (a) device-resident run: hyper-parameters stay in HBM, lower bound and M-step on the device (vbhem_em_run)
        bs.em_begin({k: v for k, v in hyp.items()}, opt)
        bs.em_run(tau, 1, 0.0)                                            # warm-up
        bs.em_begin({k: v for k, v in hyp.items()}, opt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dev = bs.em_run(tau, opt["max_iter"], opt["minDiff"])
        t_dev = (time.perf_counter() - t0) / max(1, len(dev["LogLs"]))
        # (b) host mirror: statistics D2H, M-step and lower bound in numpy (h3m.vbhem_h3m_c_step_fc)
        run = h3m.vbhem_h3m_c_step_fc({k: v for k, v in hyp.items()}, bs, opt)
        check(L.vbhem_set_profile(bs._h, 1))
        # (c) SURVEY 8(d): the E-step again with the centres after 5 EM iterations (sharper responsibilities)
        bs.em_begin({k: v for k, v in hyp.items()}, opt)
        bs.em_run(tau, 5, 0.0)
        late = []
        for _ in range(4):
            flush.fill_(1)
            bs.em_estep(tau)
            bs.synchronize()
            t = (C.c_double * 8)()
            check(L.vbhem_last_timings(bs._h, t))
            late.append(list(t))
        late = np.median(np.array(late[1:]), axis=0)
        late_counters = bs.counters()
        vbhem_iter = dict(ms=1e3 * t_dev, iters=int(dev["iters"]), LL_first=float(dev["LogLs"][0]),
                          LL_last=float(dev["LogLs"][-1]), host_mirror_ms=run["ms_per_iter"],
                          LL_last_host_mirror=float(run["LogLs"][-1]),
                          estep_after_5_iters=dict(estep_ms=float(late[5]), first_pass_ms=float(late[1]),
                                                   second_pass_ms=float(late[3]),
                                                   active_pairs=int(late_counters["active_pairs"]),
                                                   fallback_pairs=int(late_counters["general_pairs"])),
                          note="ms = wall time of vbhem_em_run / E-steps taken: fused E-step, lower bound, M-step and "
                               "convergence test through the C ABI with the centre hyper-parameters resident in HBM; "
                               "host_mirror_ms = median iteration of h3m.vbhem_h3m_c_step_fc (statistics D2H, numpy "
                               "M-step and lower bound)")

    # ---- (K,S) grid: every candidate's whole EM run, batched (vbhem_em_run_groups), N = 1 only
    grid_em = None
    if world == 1 and grid and a.grid_em:
        opt = dict(synth.DEFAULT_OPT, tau=tau, max_iter=a.grid_em, minDiff=1e-5)
        check(L.vbhem_set_profile(bs._h, 0))
        bs.em_begin({k: v for k, v in hyp.items()}, opt, group=grp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = bs.em_run_groups(tau, opt["max_iter"], opt["minDiff"])
        t_run = time.perf_counter() - t0
        check(L.vbhem_set_profile(bs._h, 1))
        passes = [len(x) for x in out["LogLs"]]
        centre_passes = int(sum(p * k for p, (k, _) in zip(passes, models)))
        grid_em = dict(seconds=t_run, models=len(models), max_iter=opt["max_iter"], minDiff=opt["minDiff"],
                       e_steps=int(max(passes)), centre_passes=centre_passes,
                       centre_passes_uncompacted=int(max(passes)) * K, compactions=int(bs.counters()["compactions"]),
                       pairs_per_s=centre_passes * a.bases / t_run,
                       note="vbhem_em_run_groups: all (K,S) candidates to their own convergence in one loop; "
                            "centre_passes = sum over candidates of K x E-steps it needed (the useful work); the loop "
                            "compacts the centre set to the running candidates, so its cost follows that sum "
                            "instead of e_steps x all centres")

    # ---- strong scaling of the north-star size (100k bases in total): E-step and whole EM iterations
    strong = None
    if world > 1 and not grid:
        nbs = a.bases // world
        sub = {k: (v[rank * nbs:(rank + 1) * nbs] if isinstance(v, np.ndarray) and v.shape[:1] == (a.bases,) else v)
               for k, v in pack.pack_base(base0).items()}
        sub["N"] = nbs
        bs.close()
        bs = make_handle(sub)
        s_ms, _, _, _, _ = timed_run(bs, a.steps, W)
        strong = dict(bases_total=nbs * world, ms_per_estep=s_ms / a.steps,
                      pairs_per_s=nbs * world * K * a.steps / (s_ms * 1e-3))
        # device-resident EM iterations over the sharded set: E-step, NCCL all-reduce, lower bound + M-step on
        # every rank (dist.ShardedEStep.em_run drives the same vbhem_em_* calls)
        from vbhem_b200.dist import ShardedEStep
        opt = dict(synth.DEFAULT_OPT, tau=tau, max_iter=6, minDiff=0.0)
        check(L.vbhem_set_profile(bs._h, 0))
        sh = ShardedEStep.__new__(ShardedEStep)
        sh.dist, sh.group, sh.world, sh.torch, sh.device, sh.bs = dist, None, world, torch, local, bs
        sh._compute = sh._finalize = None
        sh.em_run({k: v for k, v in hyp.items()}, dict(opt, max_iter=1))      # warm-up
        torch.cuda.synchronize()
        dist.barrier()
        emr = sh.em_run({k: v for k, v in hyp.items()}, opt)
        tt = torch.tensor([emr["ms_per_iter"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        strong["vbhem_iter_ms"] = float(tt.item())
        strong["LL_last"] = float(emr["LogLs"][-1])
        strong["iters"] = int(emr["iters"])
    bs.close()

    # ---- e2e through the host-buffer C ABI: pack (H2D) + E-step + results (D2H), every step
    e2e = None
    if not a.no_e2e:
        pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
        pbp = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in pb.items()}
        Lh = torch.empty((K, a.bases), dtype=torch.float64).pin_memory().numpy()
        zh = torch.empty((K, a.bases), dtype=torch.float64).pin_memory().numpy()
        sh = torch.empty(ns, dtype=torch.float64).pin_memory().numpy()
        h2d = sum(v.nbytes for v in pb.values() if isinstance(v, np.ndarray)) + sum(v.nbytes for v in H.values())
        d2h = Lh.nbytes + zh.nbytes + sh.nbytes + pre.nbytes

        b2 = BaseSet(pbp, device=local, packed=True)

        def e2e_step():
            # base set host -> device in slices, overlapped with the kernels of earlier slices; centres H2D;
            # L_elbo, z_hat rows D2H slice by slice; statistics D2H
            check(L.vbhem_estep_hyp_stream(b2._h, pbp["Sb"].ctypes.data_as(ip), P(pbp["prior"]), P(pbp["A"]),
                                           P(pbp["mu"]), P(pbp["cov"]), P(pbp["Ntilde"]), a.slices, K,
                                           Sr.ctypes.data_as(ip), S, None, 1, tau, P(H["alpha"]), P(H["eta"]),
                                           P(H["eps"]), P(H["m"]), P(H["W"]), P(H["v"]), P(H["lam"]), P(Lh), P(zh),
                                           P(sh), P(pre)))
            return float(sh[0])

        def e2e_step_serial():
            b2.repack(pbp, packed=True)                                   # base set: host -> device
            check(L.vbhem_estep_hyp(b2._h, K, Sr.ctypes.data_as(ip), S, None, 1, tau, P(H["alpha"]), P(H["eta"]),
                                    P(H["eps"]), P(H["m"]), P(H["W"]), P(H["v"]), P(H["lam"]), P(Lh), P(zh),
                                    P(sh), P(pre)))                       # centres H2D; L_elbo, z_hat, stats D2H
            return float(sh[0])

        def wall(fn, n):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if dist is not None:
                tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            return dt

        ne = max(3, min(a.steps, 10))
        dt = wall(e2e_step, ne)
        dts = wall(e2e_step_serial, 3)
        e2e = dict(value=pairs_step * ne / dt, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                   ms_per_step=1e3 * dt / ne, steps=ne, slices=a.slices, serial_ms_per_step=1e3 * dts / 3,
                   note="every step: vbhem_estep_hyp_stream from pinned host memory -- the whole base set H2D in "
                        "slices overlapped with the kernels, centres H2D, L_elbo / z_hat rows D2H slice by slice, "
                        "statistics D2H; wall clock, synchronised both sides; serial_ms_per_step = the same work "
                        "as vbhem_repack_base + vbhem_estep_hyp without overlap; shards are not all-reduced here")
        b2.close()

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=W,
                    ms_per_step=tot_ms / a.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype="f64", data="synthetic", config=config, clocks=clocks, e2e=e2e, gpu_launches=int(launches),
                    roofline=roofline, phases=phases)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if strong is not None:
            line["strong"] = strong
        if vbhem_iter is not None:
            line["vbhem_iter"] = vbhem_iter
        if grid_em is not None:
            line["grid_em"] = grid_em
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
