#!/usr/bin/env python
"""bench.py -- the VBHEM-H3M E-step on B200, BASELINE.json's metric and config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[2] -- N = 100 000 base HMMs x K = 64 centre HMMs, S = 8
states both sides, 2-D full-covariance emissions, tau = 10, Nv = 100, centres from the reference's `baseem`
initialisation; synthetic, seeded (vbhem_b200/synth.py).  One "step" is one fused E-step over every (base,
centre) pair: pre-pass, L_elbo, responsibilities, statistics, all-reduce, normalisation.

Scaling is STRONG (BASELINE north_star: "N = 100k base HMMs ... on 8xB200"): the 100 000 bases are split
contiguously over the N GPUs, centres replicated, and the raw per-centre statistics all-reduced (NCCL) INSIDE the
library's multi-GPU entry (vbhem_multi_*, include/vbhem_b200.h) -- the same entry a MATLAB host reaches through
the MEX gateway.  One process per GPU here (the driver's launch); `single_process` reports the other arrangement
(rank 0 alone driving all N GPUs through the same entry); `weak` keeps the 100 000-bases-per-GPU numbers.

Prints ONE JSON line (rank 0).  `value` = pairs/s with the base set resident in HBM (CUDA events on the launching
stream around E-step + all-reduce + normalisation, L2 flushed between steps, max over ranks); `e2e` = the same
metric through the host-buffer C ABI (every rank streams its shard from pinned host memory, results read back,
all-reduce included); `roofline` = the dominant kernel (first pass, k_lattice_bwd) against the fp64 DMMA peak
measured in the same process; `cpu_baseline` = the reference's own C E-step (oracle/_ref) on the box's host
cores, bounded sample; `parity` = observed errors against that reference at the full shape.
`--impl reference` times the CPU path alone.
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = os.path.join(ROOT, "clustering-hidden-markov-models-with-variational-bayesian-hierarchical-em_b200")

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


def make_config(a, world, grid=False, K=None, models=None):
    """The workload description -- identical for the b200 arm and the reference arm (same_config)."""
    K = a.centres if K is None else K
    cfg = dict(workload=("configs[3]: (K,S) model-selection grid K=1..16 x S=1..8, every candidate's centres in ONE "
                         "batched fused E-step (128 soft-max groups, ragged state counts)") if grid else
                        "configs[2]: synthetic base HMMs x centre HMMs, full-covariance Gaussian emissions, "
                        "baseem-initialised centres, one fused E-step per step",
               bases_total=a.bases, centres=K, states=a.states, dim=a.dim, tau=a.tau, Nv=100,
               pairs_per_step=a.bases * K,
               parallelism=f"strong scaling: bases sharded x{world} ({a.bases // world} per GPU), centres replicated, "
                           "one all-reduce of the per-centre statistics inside the C ABI (vbhem_multi_*: over NVLink "
                           "peer memory, NCCL where peers cannot be mapped)",
               l2="flushed between timed steps (256 MiB write)",
               start="every timed step starts after a device-side rendezvous of the ranks (vbhem_multi_align, outside the "
                     "timed region): an EM loop is aligned the same way by its previous iteration's all-reduce")
    if models is not None:
        cfg["models"] = models
    return cfg


# ----------------------------------------------------------------------------- CPU reference arm
def cpu_worker(a):
    """One process: the reference's C E-step on bases [lo, lo + nb) of the sample file.  Prints seconds; with --out
    also leaves L_elbo of its pairs and the per-pair statistics of its first --keep bases (parity leg)."""
    import numpy as np
    import vbhem_b200  # noqa: F401
    from oracle import cref
    z = np.load(a.sample, allow_pickle=False)
    nb, lo = a.cpu_worker, a.lo
    pb = dict(N=nb, maxSb=int(z["maxSb"]), d=int(z["d"]), covmode=int(z["covmode"]))
    for k in ("Sb", "prior", "A", "mu", "cov", "Ntilde"):
        pb[k] = np.ascontiguousarray(z["b_" + k][lo:lo + nb])
    pc = dict(K=int(z["K"]), maxSr=int(z["maxSr"]))
    for k in ("Sr", "logPi", "logA", "m", "vW", "cterm", "logOmega", "W", "v"):
        pc[k] = np.ascontiguousarray(z["c_" + k])
    kind = "reference" if cref.have_ref() else "port"
    fn = cref.ref_estep_pairs if kind == "reference" else cref.port_estep_pairs
    small = dict(pb, N=4, **{k: pb[k][:4] for k in ("Sb", "prior", "A", "mu", "cov", "Ntilde")})
    fn(small, pc, a.tau)                                           # touch the library once
    t0 = time.perf_counter()
    out = fn(pb, pc, a.tau)
    dt = time.perf_counter() - t0
    assert np.isfinite(out["L_elbo"]).all()
    if a.out:
        keep = min(a.keep, nb)
        np.savez(a.out, L_elbo=out["L_elbo"], **{k: np.ascontiguousarray(out[k][:, :keep]) for k in
                                                 ("nu1", "xi", "pr", "mu", "Mu")})
    print(json.dumps(dict(seconds=dt, pairs=nb * pc["K"], kind=kind)))


def write_sample(path, pb, pc, n):
    """Packed arrays of the first n bases + the packed centre set, for the CPU workers."""
    import numpy as np
    np.savez(path, maxSb=pb["maxSb"], d=pb["d"], covmode=pb["covmode"], K=pc["K"], maxSr=pc["maxSr"],
             **{"b_" + k: pb[k][:n] for k in ("Sb", "prior", "A", "mu", "cov", "Ntilde")},
             **{"c_" + k: pc[k] for k in ("Sr", "logPi", "logA", "m", "vW", "cterm", "logOmega", "W", "v")})


def run_cpu_workers(sample, nb, cores, tau, outdir=None, keep=0):
    """`cores` concurrent worker processes (the reference's parfor, vbhem_h3m_c.m:28, is process-parallel)."""
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", str(nb), "--sample", sample, "--tau", str(tau),
           "--keep", str(keep)]
    t0 = time.perf_counter()
    procs = []
    for w in range(cores):
        extra = ["--lo", str(w * nb)]
        if outdir:
            extra += ["--out", os.path.join(outdir, f"w{w}.npz")]
        procs.append(subprocess.Popen(cmd + extra, stdout=subprocess.PIPE, text=True,
                                      env=dict(os.environ, OMP_NUM_THREADS="1", CUDA_VISIBLE_DEVICES="")))
    outs = [json.loads(p.communicate()[0].strip().splitlines()[-1]) for p in procs]
    wall = time.perf_counter() - t0
    worst = max(o["seconds"] for o in outs)
    pairs = sum(o["pairs"] for o in outs)
    return dict(value=pairs / worst, unit=UNIT, cores=cores, kind=outs[0]["kind"],
                per_core=statistics.median(o["pairs"] / o["seconds"] for o in outs),
                sample=f"{cores} processes x {nb} bases of the workload x all centres = {pairs} pairs through the "
                       f"reference MEX gateway (all six per-pair outputs for every pair; the MATLAB-side z_hat and "
                       f"statistics reduction are not included); slowest worker {worst:.2f} s, wall {wall:.1f} s")


LATTICE_SOURCES = ("vbhem_lattice.cu", "vbhem_math.cuh", "vbhem_internal.cuh")


def source_sha1(files=LATTICE_SOURCES):
    """Identity of the kernel sources a committed ncu summary was captured from."""
    h = hashlib.sha1()
    for f in files:
        with open(os.path.join(PKG, "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def ncu_dram_bytes(path, files=LATTICE_SOURCES):
    """(dram__bytes_read.sum + dram__bytes_write.sum per launch, note) from a profiles/*_ncu.json summary; the value
    is used only when the summary says it was captured from the kernel sources as they are now."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    try:
        with open(path) as f:
            nc = json.load(f)
        if nc.get("source_sha1") != source_sha1(files):
            return None, f"{os.path.basename(path)} was captured from other kernel sources: not quoted"
        v = sum(float(nc[k]["value"]) * scale[nc[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        return v, f"profiles/{os.path.basename(path)} (ncu --set full of this kernel source, bytes per launch)"
    except Exception as e:  # noqa: BLE001
        return None, f"no ncu capture of the current kernel ({type(e).__name__})"


# ----------------------------------------------------------------------------- clocks, placement
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


def pin_to_gpu_numa_node(index):
    """Run this rank (and allocate its pinned buffers, first touch) on the CPUs of its GPU's NUMA node."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read())
        if node < 0:
            return "single NUMA node (numa_node = -1): no pinning needed"
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, cpus & os.sched_getaffinity(0) or cpus)
        return f"pinned to NUMA node {node} ({len(cpus)} CPUs)"
    except Exception as e:  # noqa: BLE001
        return f"not pinned ({type(e).__name__})"


# ----------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--bases", type=int, default=100_000, help="base HMMs in total (split over the GPUs)")
    ap.add_argument("--centres", type=int, default=64)
    ap.add_argument("--states", type=int, default=8)
    ap.add_argument("--dim", type=int, default=2)
    ap.add_argument("--tau", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1001)
    ap.add_argument("--grid-em", type=int, default=0, metavar="MAX_ITER",
                    help="--workload grid: also run every candidate's EM to convergence (at most MAX_ITER passes)")
    ap.add_argument("--cpu-worker", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--sample", default="", help=argparse.SUPPRESS)
    ap.add_argument("--lo", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--out", default="", help=argparse.SUPPRESS)
    ap.add_argument("--keep", type=int, default=0, help=argparse.SUPPRESS)
    ap.add_argument("--cpu-bases", type=int, default=0, help="bases per CPU worker (0 = default sample)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip weak scaling, EM iterations, single-process leg")
    ap.add_argument("--slices", type=int, default=6, help="slices of the streamed (e2e) E-step")
    ap.add_argument("--workload", default="estep", choices=["estep", "grid", "fb", "exprmt1"],
                    help="estep = BASELINE configs[2] (the metric's workload, default); grid = configs[3], the (K,S) "
                         "model-selection grid K=1..16 x S=1..8 over 20k bases in one batched call; fb = configs[4], "
                         "the VB-HMM forward-backward pass over 1M sequences of length 50; exprmt1 = configs[0], the "
                         "reference's synthetic experiment end to end (learn 40 HMMs, cluster over K=1..6 x S=1..5)")
    ap.add_argument("--trials", type=int, default=20, help="--workload exprmt1: initialisations per (K,S) candidate")
    a = ap.parse_args()
    S, d, tau = a.states, a.dim, a.tau
    if a.workload == "fb":
        from bench_fb import main_fb
        return main_fb(a)
    if a.workload == "exprmt1":
        from bench_exprmt1 import main_exprmt1
        return main_exprmt1(a)
    if a.cpu_worker:
        return cpu_worker(a)
    grid = a.workload == "grid"
    if grid and a.bases == 100_000:
        a.bases = 20_000

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

    import numpy as np
    import vbhem_b200  # noqa: F401
    from vbhem_b200 import pack, synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))

    # ---- workload (numpy only: no CUDA yet).  Every rank builds the same 100 000 bases and takes its shard.
    base0, _ = synth.make_base_set(a.bases, S=S, d=d, C=a.centres, seed=a.seed)
    grp, ngroups, models = None, 1, None
    if grid:
        hyp, grp, models = synth.model_selection_grid(base0, seed=a.seed)
        ngroups = len(models)
    else:
        hyp = synth.init_centres_baseem(base0, a.centres, S, seed=a.seed)
    K = int(np.asarray(hyp["Sr"]).shape[0])
    config = make_config(a, world, grid, K, ngroups if grid else None)
    pb0 = pack.pack_base(base0)

    def centre_pack():
        from util import prepass_natural
        cen = prepass_natural(hyp, "full")
        if grp is not None:
            cen["group"] = grp
        return cen, pack.pack_centres(cen)

    # ---- reference arm: CPU only, rank 0 only; each step = a bounded sample of the workload
    if a.impl == "reference":
        if rank != 0:
            return
        _, pc = centre_pack()
        per_base_s = sum(float(r) ** 2 for r in np.asarray(hyp["Sr"])) / 64.0 * 60e-6   # ~60 us per 8 x 8 pair and core
        nb = a.cpu_bases or max(256, min(3072, int(240.0 / max(1, a.steps) / per_base_s)))
        nb = min(nb, a.bases // cores)
        with tempfile.TemporaryDirectory() as td:
            sample = os.path.join(td, "sample.npz")
            write_sample(sample, pb0, pc, cores * nb)
            for _ in range(a.warmup):                                # warm-ups on a sliver of the sample
                run_cpu_workers(sample, min(64, nb), cores, tau)
            vals = [run_cpu_workers(sample, nb, cores, tau) for _ in range(a.steps)]
        v = statistics.median(x["value"] for x in vals)
        ms = 1e3 * (cores * nb * K) / v
        emit(dict(metric=METRIC, value=v, unit=UNIT, n_gpus=a.gpus, steps=a.steps, warmup=a.warmup, ms_per_step=ms,
                  higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f64", data="synthetic",
                  impl="reference", config=config,
                  cpu_baseline=dict(value=v, unit=UNIT, cores=vals[0]["cores"], kind=vals[0]["kind"],
                                    sample=vals[0]["sample"]),
                  sample_note=f"each step = {cores * nb} bases x {K} centres of the workload (cost is linear in pairs); "
                              "warm-up steps use 64 bases per process",
                  e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0))
        return

    # ---- CPU baseline + parity oracle first (before CUDA is initialised in this process), rank 0, N = 1 only
    cpu, oracle = None, None
    if rank == 0 and world == 1 and not a.no_cpu:
        from oracle import cref
        cen, pc = centre_pack()
        # ~12 s of CPU work per core (the reference's C takes ~60 us per pair of 8-state HMMs, less for smaller centres)
        per_base_s = sum(float(r) ** 2 for r in np.asarray(hyp["Sr"])) / 64.0 * 60e-6
        nb = a.cpu_bases or max(32, min(3072, a.bases // cores, int(12.0 / per_base_s)))
        keep = min(128, nb)
        with tempfile.TemporaryDirectory() as td:
            sample = os.path.join(td, "sample.npz")
            write_sample(sample, pb0, pc, cores * nb)
            cpu = run_cpu_workers(sample, nb, cores, tau, outdir=td, keep=keep)
            outs = [np.load(os.path.join(td, f"w{w}.npz")) for w in range(cores)]
            L_ref = np.concatenate([o["L_elbo"] for o in outs], axis=1)            # [K, cores * nb]
            sub = np.concatenate([np.arange(w * nb, w * nb + keep) for w in range(cores)])
            pairs_sub = {k: np.concatenate([o[k] for o in outs], axis=1) for k in ("nu1", "xi", "pr", "mu", "Mu")}
        zz = cref.port_zhat(L_ref, pc["logOmega"], pb0["Ntilde"][:cores * nb], pc.get("group"))
        zs = cref.port_zhat(np.ascontiguousarray(L_ref[:, sub]), pc["logOmega"], pb0["Ntilde"][sub], pc.get("group"))
        red = cref.port_reduce(zs["Z"], dict(pairs_sub, L_elbo=np.ascontiguousarray(L_ref[:, sub])), pc["maxSr"],
                               pb0["d"], pb0["covmode"])
        oracle = dict(n=cores * nb, L=L_ref, zhat=zz["zhat"], sub=sub, red=red, Nj=zs["Nj"], lb=zs["lb"], cen=cen)
        pc_full = pc

    import torch
    from vbhem_b200._lib import check, lib
    from vbhem_b200.api import BaseSet, MultiBaseSet
    from vbhem_b200.dist import shard_bounds

    torch.cuda.set_device(local)
    placement = pin_to_gpu_numa_node(local)
    dist = None
    uid = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")    # host-side waits: an NCCL barrier would spin on the waiting GPUs
        box = [MultiBaseSet.unique_id() if rank == 0 else None, MultiBaseSet.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)                       # ids of the strong / weak communicators
        uid = box
    L = lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = (C.c_double * 2)()
    check(L.vbhem_measure_fp64_peak(local, peak))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def rank_max(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rank_sum(x):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def shard_of(pb, r):
        lo, hi = shard_bounds(pb["N"], world, r)
        out = {k: (v[lo:hi] if isinstance(v, np.ndarray) and v.shape[:1] == (pb["N"],) else v) for k, v in pb.items()}
        out["N"] = hi - lo
        return out

    def make_set(pb_shard, which):
        ms_ = MultiBaseSet(pb_shard, devices=[local], packed=True, world=world, rank_first=rank,
                           unique_id=None if uid is None else uid[which])
        if dist is not None:      # all-reduce over peer memory: the ranks exchange the IPC handles of their exchange areas
            try:
                mine = ms_.peer_export()
            except Exception:     # noqa: BLE001  (no peer access on this box: NCCL stays in charge)
                mine = None
            handles = [None] * world
            dist.all_gather_object(handles, mine)
            if all(h is not None for h in handles):
                ms_.peer_import(handles)
        return ms_

    def timed_run(ms, P, steps, warm, profile=False):
        """The timed region: `steps` E-steps (+ all-reduce + normalisation), each timed on the device with the
        events the library records on the launching stream, L2 flushed in between.  profile: CUDA events between
        the phases of every E-step as well (they break the programmatic launch chain: separate, untimed pass)."""
        h = ms.shard_handle(0)
        check(L.vbhem_set_profile(h, 1 if profile else 0))
        sampler = ClockSampler(local)
        sampler.start()
        for _ in range(warm):
            ms.estep_hyp_dev(P, tau)
        ms.synchronize()
        barrier()
        sampler.armed = True
        tot, phases, launches = 0.0, np.zeros(8), 0
        for _ in range(steps):
            flush.zero_()
            barrier()
            ms.align()                 # device-side rendezvous: the ranks' E-steps start together (the hosts' launch skew
            ms.estep_hyp_dev(P, tau)   # after the NCCL barrier would otherwise be waited out inside the all-reduce)
            ms.synchronize()
            tot += ms.last_ms()
            if profile:
                t = (C.c_double * 8)()
                check(L.vbhem_last_timings(h, t))
                phases += np.array(list(t))
            launches += ms.counters()["launches"] + 1 + (1 if world > 1 else 0)
        barrier()
        tot = rank_max(tot)
        sampler.armed = False
        sampler.stop_flag = True
        sampler.join()
        check(L.vbhem_set_profile(h, 0))
        return tot, phases / steps, launches, sampler.summary(), ms.counters()

    # ---- headline: strong arrangement, device-resident
    pbs = shard_of(pb0, rank)
    ms = make_set(pbs, 0)
    P = ms.prepare_hyp(hyp, grp)
    allreduce_mode = ms.allreduce_mode()
    tot_ms, _, launches, clocks, counters = timed_run(ms, P, a.steps, W)
    nprof = max(3, min(a.steps, 10))
    prof_ms, ph, _, _, _ = timed_run(ms, P, nprof, 1, profile=True)
    pairs_step = a.bases * K
    value = pairs_step * a.steps / (tot_ms * 1e-3)
    nloc = pbs["N"]
    if grid:   # algorithmic flops of the reference per pair depend on the centre's own state count
        per = [flops_per_pair(S, int(r), d, tau) for r in P["Sr"]]
        fl = dict(first_pass=sum(f["first_pass"] for f in per) / K, total=sum(f["total"] for f in per) / K)
    else:
        fl = flops_per_pair(S, S, d, tau)
    first_ms = ph[1]
    ach = nloc * K * fl["first_pass"] / (first_ms * 1e-3) / 1e12
    alg_bytes = nloc * (S + S * S + S * d + S * d * d) * 8 + nloc * K * 8
    traffic, traffic_src = None, "the committed capture is of the 100 000-base single-GPU launch only"
    if (nloc, K, S, d, tau) == (100_000, 64, 8, 2, 10) and not grid:
        traffic, traffic_src = ncu_dram_bytes(os.path.join(ROOT, "profiles", "r02d_k_lattice_bwd_ncu.json"))
    roofline = dict(bound="tensor", kernel="k_lattice_bwd (first pass: emission expectation + backward lattice + "
                                           "termination for every pair; fp64 DMMA m8n8k4)",
                    achieved=ach, peak=float(peak[0]), unit="TFLOP/s", frac=ach / float(peak[0]),
                    peak_source="fp64 DMMA chain measured in this process (vbhem_measure_fp64_peak); "
                                "MEASURED_PEAKS.json has no fp64 entry; DFMA chain: %.2f TFLOP/s" % float(peak[1]),
                    flops_per_pair=fl["first_pass"], kernel_ms=first_ms, pairs_per_launch=nloc * K,
                    traffic=traffic, traffic_source=traffic_src,
                    hbm=dict(algorithmic_bytes=alg_bytes, achieved_gbs=alg_bytes / (first_ms * 1e-3) / 1e9),
                    whole_step=dict(flops_per_pair=fl["total"],
                                    achieved=pairs_step * fl["total"] / (tot_ms / a.steps * 1e-3) / 1e12 / world,
                                    note="per GPU: reference-algorithm flops of the complete E-step / step time; the "
                                         "second pass runs only on pairs with Z > 1e-8, as the reference's reduction "
                                         "does"))
    phases = dict(stage_ms=ph[0], first_pass_ms=ph[1], zhat_list_ms=ph[2], second_pass_ms=ph[3], reduce_ms=ph[4],
                  estep_ms=ph[5], step_ms_with_phase_events=prof_ms / nprof,
                  active_pairs=int(rank_sum(counters["active_pairs"])),
                  fallback_pairs=int(rank_sum(counters["general_pairs"])),
                  note="rank 0's shard, separate pass with phase events recorded inside the E-step")

    # ---- parity at the full shape (N = 1): observed errors against the reference's C, not just pass / fail
    parity = None
    if oracle is not None:
        ms.estep_hyp_dev(P, tau, finalize=False)
        ms.synchronize()
        Lg, zg = ms.fetch(K)
        n = oracle["n"]
        rel = lambda g, w: float(np.max(np.abs(g - w) / np.maximum(np.abs(w), 1e-300)))
        parity = dict(bases_compared=n, pairs_compared=n * K,
                      L_elbo_max_rel=rel(Lg[:, :n], oracle["L"]),
                      zhat_max_abs=float(np.max(np.abs(zg[:, :n] - oracle["zhat"]))),
                      zhat_max_rel=float(np.max(np.abs(zg[:, :n] - oracle["zhat"]) /
                                                np.maximum(oracle["zhat"], 1e-12))))
        # reduced statistics, N-sized lower-bound terms and the bound itself on the sub-sample whose per-pair
        # statistics the oracle kept (the oracle's reduction over 100 000 x 64 cell entries is not affordable)
        sub = oracle["sub"]
        pbsub = {k: (np.ascontiguousarray(v[sub]) if isinstance(v, np.ndarray) and v.shape[:1] == (a.bases,) else v)
                 for k, v in pb0.items()}
        pbsub["N"] = len(sub)
        with BaseSet(pbsub, device=local, packed=True) as bsub:
            got = bsub.estep_hyp(hyp, tau, group=grp)
            want = pack.stats_to_natural(dict(oracle["red"], Nj=oracle["Nj"], lb=oracle["lb"]), d, 1)
            serr = {}
            for k in ("Nj", "nu1", "xi", "pr", "mu", "Mu"):
                g, w = got[k], want[k]
                # relative to the largest entry of the centre's block (Nj: of the vector -- a centre that attracts no
                # base has Nj ~ 1e-46, where an element-wise ratio means nothing)
                scale = np.maximum(np.max(np.abs(w), axis=tuple(range(1, w.ndim)), keepdims=True), 1e-300) \
                    if w.ndim > 1 else np.max(np.abs(w))
                serr[k] = float(np.max(np.abs(g - w) / scale))
            parity.update(stats_bases=len(sub), stats_max_rel=max(serr.values()), stats_rel_by_block=serr,
                          Lt1_rel=float(abs(got["lb"][0] - want["lb"][0]) / abs(want["lb"][0])),
                          Lt7_abs=float(abs(got["lb"][1] - want["lb"][1])))
            if not grid:
                from oracle import numpy_ref as nr
                opt = dict(synth.DEFAULT_OPT, tau=tau)
                bsub.em_begin(hyp, opt)
                bsub.em_estep(tau)
                lb_dev = float(bsub.em_lower_bound()[0, 10])
                centres = [dict(eta=hyp["eta"][j], epsilon=hyp["epsilon"][j], m=hyp["m"][j], W=hyp["W"][j],
                                v=hyp["v"][j], lam=hyp["lam"][j]) for j in range(K)]
                pre = [nr.prepass(c, "full") for c in centres]
                lb_ref, _ = nr.lower_bound(want["lb"][0], want["lb"][1], want["Nj"], oracle["cen"]["logOmega"],
                                           hyp["alpha"], centres, pre, opt, "full")
                parity.update(LB_rel=float(abs(lb_dev - lb_ref) / abs(lb_ref)), LB_device=lb_dev, LB_oracle=float(lb_ref))
            # the reference arm's OWN work on the GPU: the literal MEX boundary, all six per-pair outputs
            # of EVERY pair (no skip rule, no z_hat, no reduction), host buffers in and out -- what cpu_baseline times
            try:
                pcs, ns, r_ = pc_full, len(sub), int(pc_full["maxSr"])
                dI = lambda x: np.ascontiguousarray(x, dtype=np.int32).ctypes.data_as(C.POINTER(C.c_int32))
                dD = lambda x: x.ctypes.data_as(C.POINTER(C.c_double))
                o = dict(L=np.zeros((K, ns)), nu1=np.zeros((K, ns, r_)), xi=np.zeros((K, ns, r_, r_)), pr=np.zeros((K, ns, r_)),
                         mu=np.zeros((K, ns, d, r_)), Mu=np.zeros((K, ns, d * d, r_)))      # full covariance (the bench workloads)
                Srs = np.ascontiguousarray(pcs["Sr"], dtype=np.int32)
                call = lambda: check(L.vbhem_estep_pairs(bsub._h, K, dI(Srs), r_, tau, dD(pcs["logPi"]), dD(pcs["logA"]),
                                                         dD(pcs["m"]), dD(pcs["vW"]), dD(pcs["cterm"]), dD(o["L"]),
                                                         dD(o["nu1"]), dD(o["xi"]), dD(o["pr"]), dD(o["mu"]), dD(o["Mu"])))
                call()                                                                # warm-up
                reps, t0 = 3, time.perf_counter()
                for _ in range(reps):
                    call()
                dt = (time.perf_counter() - t0) / reps
                lit = dict(L_elbo=o["L"].T)
                npair = len(sub) * K
                same_work = dict(pairs=npair, gpu_pairs_per_s=npair / dt, cpu_pairs_per_s=cpu["value"],
                                 ratio=npair / dt / cpu["value"],
                                 L_elbo_max_rel=rel(lit["L_elbo"].T, Lg[:, sub]),
                                 note="vbhem_estep_pairs (Boundary-1: L_elbo, nu_1, xi, and the three emission "
                                      "statistics of every pair, no skip rule) over stats_bases bases x all centres, wall "
                                      "clock of the C-ABI call including the device-to-host copy of the six outputs "
                                      "(1 KB per pair, pageable host arrays) -- the same outputs the reference arm "
                                      "produces; the headline compares the fused E-step instead")
                cpu["same_work"] = same_work
            except Exception as e:  # noqa: BLE001
                cpu["same_work"] = dict(error=f"{type(e).__name__}: {e}")
        parity["note"] = ("L_elbo, z_hat: the rows of the first bases_compared bases of the FULL run, element-wise "
                          "against the reference's C (oracle/_ref); statistics / Lt1 / Lt7 / LB: fused E-step over "
                          "stats_bases of them against the oracle's skip-rule reduction and numpy lower bound; "
                          "tolerances of the tests: 1e-9 relative (z_hat: + 1e-12 absolute)")
        oracle = None

    # ---- whole VBHEM iterations over the sharded set (E-step, all-reduce, lower bound, M-step, convergence test)
    vbhem_iter = None
    if not a.no_extras and not grid:
        opt = dict(synth.DEFAULT_OPT, tau=tau, max_iter=6, minDiff=0.0)
        ms.em_begin(hyp, opt)
        ms.em_run(tau, 1, 0.0)                                            # warm-up
        ms.em_begin(hyp, opt)
        barrier()
        t0 = time.perf_counter()
        run = ms.em_run(tau, opt["max_iter"], opt["minDiff"])
        t_it = rank_max((time.perf_counter() - t0) / max(1, len(run["LogLs"])))
        # the E-step again with the centres after 5 EM iterations (sharper responsibilities)
        ms.em_begin(hyp, opt)
        ms.em_run(tau, 5, 0.0)
        P5 = ms.prepare_hyp(ms.em_fetch())
        late = []
        for _ in range(4):
            flush.zero_()
            barrier()
            ms.align()
            ms.estep_hyp_dev(P5, tau)
            ms.synchronize()
            late.append(ms.last_ms())
        vbhem_iter = dict(ms=1e3 * t_it, iters=int(run["iters"]), LL_first=float(run["LogLs"][0]),
                          LL_last=float(run["LogLs"][-1]),
                          estep_after_5_iters=dict(estep_ms=rank_max(statistics.median(late[1:])),
                                                   active_pairs=int(rank_sum(ms.counters()["active_pairs"]))),
                          note="ms = wall time of vbhem_multi_em_run / E-steps taken: fused E-step, all-reduce, lower "
                               "bound, M-step and convergence test through the C ABI with the centre hyper-parameters "
                               "resident in HBM on every GPU")
        if world > 1:   # the sharded run must reproduce the single-GPU lower bounds (rank 0 holds all bases once)
            barrier()
            if rank == 0:
                with BaseSet(pb0, device=local, packed=True) as b1:
                    b1.em_begin(hyp, opt)
                    one = b1.em_run(tau, opt["max_iter"], opt["minDiff"])
                relL = float(np.max(np.abs(one["LogLs"] - run["LogLs"]) / np.abs(one["LogLs"])))
                vbhem_iter["single_gpu_check"] = dict(LL_last_single_gpu=float(one["LogLs"][-1]),
                                                      LogLs_max_rel=relL, tolerance=1e-10)
                assert relL <= 1e-10, f"sharded EM run differs from the single-GPU run: {relL:.3e}"
            barrier()

    # ---- (K,S) grid: every candidate's whole EM run, batched (vbhem_multi_em_run_groups)
    grid_em = None
    if grid and a.grid_em:
        opt = dict(synth.DEFAULT_OPT, tau=tau, max_iter=a.grid_em, minDiff=1e-5)
        ms.em_begin(hyp, opt, group=grp)
        barrier()
        t0 = time.perf_counter()
        out = ms.em_run_groups(tau, opt["max_iter"], opt["minDiff"])
        t_run = rank_max(time.perf_counter() - t0)
        passes = [len(x) for x in out["LogLs"]]
        centre_passes = int(sum(p * k for p, (k, _) in zip(passes, models)))
        grid_em = dict(seconds=t_run, models=len(models), max_iter=opt["max_iter"], minDiff=opt["minDiff"],
                       e_steps=int(max(passes)), centre_passes=centre_passes,
                       centre_passes_uncompacted=int(max(passes)) * K, compactions=int(ms.counters()["compactions"]),
                       pairs_per_s=centre_passes * a.bases / t_run,
                       note="vbhem_multi_em_run_groups: all (K,S) candidates to their own convergence in one loop; "
                            "centre_passes = sum over candidates of K x E-steps it needed (the useful work); the loop "
                            "compacts the centre set to the running candidates")

    # ---- e2e through the host-buffer C ABI: every rank streams its shard H2D, E-step, all-reduce, results D2H
    e2e = None
    if not a.no_e2e:
        pin = lambda x: torch.from_numpy(x).pin_memory().numpy()
        pbp = {k: (pin(v) if isinstance(v, np.ndarray) else v) for k, v in pbs.items()}
        ns = ms.stats_size(K, P["r"], ngroups)
        Lh = torch.empty((K, nloc), dtype=torch.float64).pin_memory().numpy()
        zh = torch.empty((K, nloc), dtype=torch.float64).pin_memory().numpy()
        sh = torch.empty(ns, dtype=torch.float64).pin_memory().numpy()
        pre = np.zeros(K + P["r"] * K * (2 + P["r"]))
        h2d = rank_sum(sum(v.nbytes for v in pbs.values() if isinstance(v, np.ndarray)) + sum(h.nbytes for h in P["H"]))
        d2h = rank_sum(Lh.nbytes + zh.nbytes + sh.nbytes + pre.nbytes)

        def wall(fn, n):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return rank_max(time.perf_counter() - t0)

        ne = max(3, min(a.steps, 10))
        dt = wall(lambda: ms.estep_hyp_stream(pbp, P, tau, a.slices, Lh, zh, sh, pre), ne)
        dts = wall(lambda: ms.estep_hyp_stream(pbp, P, tau, 1, Lh, zh, sh, pre), 3)
        e2e = dict(value=pairs_step * ne / dt, unit=UNIT, h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                   ms_per_step=1e3 * dt / ne, steps=ne, slices=a.slices, unsliced_ms_per_step=1e3 * dts / 3,
                   note="every step, every rank: vbhem_multi_estep_hyp_stream from pinned host memory -- its shard of "
                        "the base set H2D in slices overlapped with the kernels, centres H2D, all-reduce of the "
                        "statistics, L_elbo / z_hat rows D2H slice by slice, statistics D2H; wall clock, barrier + "
                        "synchronise both sides, max over ranks; bytes are summed over the ranks; "
                        "unsliced_ms_per_step = the same call with one slice (no overlap)")
    ms.close()

    # ---- weak scaling (the driver's round-1 curve): 100 000 bases per GPU
    weak = None
    if world > 1 and not a.no_extras and not grid:
        basew = base0 if rank == 0 else synth.make_base_set(a.bases, S=S, d=d, C=a.centres, seed=a.seed + rank)[0]
        pbw = pack.pack_base(basew)
        mw = make_set(pbw, 1)
        w_ms, _, _, _, _ = timed_run(mw, P, a.steps, W)
        mw.close()
        weak = dict(bases_per_gpu=a.bases, ms_per_estep=w_ms / a.steps,
                    pairs_per_s=a.bases * world * K * a.steps / (w_ms * 1e-3))

    # ---- the same strong arrangement driven by ONE host process (the MATLAB host's situation): rank 0 takes every
    #      GPU of the job through vbhem_multi_* while the other ranks wait
    single = None
    if world > 1 and not a.no_extras and not grid:
        barrier()
        dist.barrier(group=cpu_group)
        if rank == 0:
            with MultiBaseSet(pb0, devices=list(range(world)), packed=True) as m1:
                P1 = m1.prepare_hyp(hyp)
                for _ in range(W):
                    m1.estep_hyp_dev(P1, tau)
                m1.synchronize()
                dev = []
                for _ in range(a.steps):
                    flush.zero_()
                    torch.cuda.synchronize()
                    m1.align()
                    m1.estep_hyp_dev(P1, tau)
                    m1.synchronize()
                    dev.append(m1.last_ms())
                opt = dict(synth.DEFAULT_OPT, tau=tau, max_iter=6, minDiff=0.0)
                m1.em_begin(hyp, opt)
                m1.em_run(tau, 1, 0.0)
                m1.em_begin(hyp, opt)
                t0 = time.perf_counter()
                r1 = m1.em_run(tau, opt["max_iter"], opt["minDiff"])
                t1 = (time.perf_counter() - t0) / max(1, len(r1["LogLs"]))
            single = dict(devices=world, ms_per_estep=statistics.mean(dev),
                          pairs_per_s=pairs_step / (statistics.mean(dev) * 1e-3), vbhem_iter_ms=1e3 * t1,
                          LL_last=float(r1["LogLs"][-1]),
                          note="one host process, one thread per device inside the library (ncclCommInitAll): device "
                               "time of the slowest shard, E-step + all-reduce + normalisation")
        dist.barrier(group=cpu_group)                 # the other ranks wait on the host, their GPUs stay idle

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=a.steps, warmup=W,
                    ms_per_step=tot_ms / a.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f64", data="synthetic", config=config, clocks=clocks, e2e=e2e, gpu_launches=int(launches),
                    roofline=roofline, phases=phases, placement=placement, allreduce=allreduce_mode)
        for k, v in (("cpu_baseline", cpu), ("parity", parity), ("vbhem_iter", vbhem_iter), ("grid_em", grid_em),
                     ("weak", weak), ("single_process", single)):
            if v is not None:
                line[k] = v
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
