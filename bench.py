#!/usr/bin/env python
"""bench.py - Fit-Hi-C contact pairs/sec on B200 (BASELINE.json's metric), one JSON line on stdout.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU arm (oracle port on the host cores)

Workload (config.workload): BASELINE config 2 - chr1 at 5 kb (49,851 bins), every pair within 10 Mb
(97,750,851 records, zeros kept), ICE-style biases, synthetic counts generated on the device.
At N > 1 every rank holds one such chromosome-sized shard (weak scaling): the per-distance table and
totals are all-reduced over NCCL (S, the bins and the spline are genome-wide, identical on all ranks),
p-values are per shard; q-values are ranked genome-wide across the ranks (all-reduce of the coarse p histogram
+ all-gather of the few candidate keys; `--q-scope shard` ranks per chromosome instead).
A "step" is one whole pass over the resident records: K1 histogram -> [allreduce] -> K2/K3 fit ->
K4 p-values (+ coarse p histogram) -> K5 Benjamini-Hochberg q-values.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RESOLUTION = 5000
CHR1_BINS = 49851            # ceil(249,250,621 / 5000)
MAX_DIST = 10_000_000
N_BINS = 100
DEPTH = 600.0                # Poisson mean of a d=0 pair with unit bias -> S ~ 2e8 per shard (8 shards stay < 2^31)
DECAY = 1.08
SEED = 20161108
BYTES_PER_PAIR = 48          # 12 (K1 read) + 20 (K4 read+write) + 16 (K5 read+write), BASELINE.md section 2
K4_BYTES_PER_PAIR = 20


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons; started before the warm-up (nvidia-smi takes ~100 ms to come up),
    reduced over the samples whose timestamp falls inside the timed region."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=self.tmp, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), [nm for nm, v in zip(names, f[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.tmp.name)
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.005 <= r[0] <= self.t1 + 0.005]
        use = inside if inside else rows[-20:]
        if use:
            out["sm_mhz"] = statistics.median(r[1] for r in use)
            out["sm_max_mhz"] = max(r[2] for r in use)
            out["reasons"] = sorted(set(x for r in use for x in r[3]))
            out["samples"] = len(use)
            out["window"] = "timed region" if inside else "last samples under load (timed region shorter than the sampling period)"
        return out


# =================================================================================================
# CPU arm: the oracle port of the reference path on the host cores
# =================================================================================================
def _cpu_block(args):
    """One bounded sample: a chromosome-shaped block (nb bins, all pairs within max_dist) through the whole
    reference path restated in oracle/fithic_oracle.py: histogram, binning, spline, scoring, BH."""
    nb, seed, repeat = args
    import numpy as np
    from blueberry_b200 import synth
    from oracle import fithic_oracle as fo
    bias = synth.make_bias([nb], seed)
    fc, fm = synth.make_fragments([nb], RESOLUTION)
    c = synth.make_contacts([nb], RESOLUTION, MAX_DIST, DEPTH, seed, bias)
    bd, _ = fo.read_bias_arrays(np.zeros(nb, dtype=np.int64), fm, bias[0])
    t0 = time.perf_counter()
    for _ in range(repeat):
        res = fo.fithic_arrays(fc, fm, None, c["mid1"], None, c["mid2"], c["count"], RESOLUTION, N_BINS, 0, MAX_DIST, bias=bd)
        keep = res.keep
        fo.benjamini_hochberg_correction(res.p[keep], int(keep.sum()))
    return len(c["count"]) * repeat, time.perf_counter() - t0


def cpu_baseline(cores, nb=3000, repeat=1):
    """pairs/s of the oracle port on `cores` host processes (each gets its own block)."""
    import multiprocessing as mp
    jobs = [(nb, SEED + 101 * i, repeat) for i in range(cores)]
    t0 = time.perf_counter()
    if cores == 1:
        res = [_cpu_block(jobs[0])]
        wall = res[0][1]
    else:
        with mp.get_context("spawn").Pool(cores) as pool:
            res = pool.map(_cpu_block, jobs)
        wall = max(r[1] for r in res)              # slowest worker's compute time (generation excluded)
    pairs = sum(r[0] for r in res)
    return pairs / wall, pairs, wall, time.perf_counter() - t0


def run_reference_arm(args, out_fd):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    cores = min(cores, 64)
    nb = 2600
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline(cores, nb=nb)
    vals, per_step = [], []
    for _ in range(args.steps):
        v, pairs, wall, _ = cpu_baseline(cores, nb=nb)
        vals.append(v); per_step.append(wall)
    value = statistics.mean(vals)
    sample = "%d blocks/step of %d bins (all pairs within 10 Mb at 5 kb, ~%.1fM records each), one per core" % (
        cores, nb, (2001 * nb - 2000 * 2001 // 2) / 1e6)
    line = {
        "impl": "reference", "metric": "fithic_contact_pairs_per_sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(per_step),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg2 chr1@5kb pairs within 10Mb (bounded sample per step)", "resolution": RESOLUTION,
                   "n_bins": N_BINS, "max_dist": MAX_DIST},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(out_fd, line)


# =================================================================================================
# GPU arm
# =================================================================================================
def run_ours(args, out_fd):
    import numpy as np
    import torch
    import torch.distributed as dist
    from blueberry_b200 import _lib
    from blueberry_b200.engine import BiasTables, HostPipeline, PassEngine, Shard

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    # --workload cfg4: BASELINE config 4's shape (chr1 @ 1 kb, 249,251 bins, 2 Mb cap, 496,750,251 records, sparse
    # counts, second pass after outlier removal); the default (and what the driver measures) is config 2
    global RESOLUTION, MAX_DIST, DEPTH
    two_pass = False
    if args.workload == "cfg4":
        RESOLUTION, MAX_DIST, DEPTH = 1000, 2_000_000, 60.0
        if args.bins == CHR1_BINS:
            args.bins = 249251
        two_pass = True
    # --workload cfg5: one GPU's share of BASELINE config 5 (genome-wide 1 kb, 2 Mb cap, 3,036,315 bins over 8 GPUs):
    # 379,540 bins per GPU as one shard, every pair within 2 Mb (759 M records), single pass with q-values
    if args.workload == "cfg5":
        RESOLUTION, MAX_DIST, DEPTH = 1000, 2_000_000, 25.0
        if args.bins == CHR1_BINS:
            args.bins = 379540
    R, nb, K = RESOLUTION, args.bins, MAX_DIST // RESOLUTION
    P = int(lib.bbk_synth_n_pairs(nb, K))
    # ---- synthetic shard, generated on the device (not timed)
    rng = np.random.default_rng(SEED + 7919 * rank)
    bias_host = np.exp(rng.normal(0.0, 0.25, size=nb))
    bias_dev = torch.from_numpy(bias_host).to(dev)
    mid1 = torch.empty(P, dtype=torch.int32, device=dev)
    mid2 = torch.empty(P, dtype=torch.int32, device=dev)
    count = torch.empty(P, dtype=torch.int32, device=dev)
    _lib.check(lib.bbk_synth_contacts(nb, K, R, DEPTH, DECAY, SEED + rank, _lib.ptr(bias_dev), _lib.ptr(mid1), _lib.ptr(mid2),
                                      _lib.ptr(count), _lib.stream_ptr()), "bbk_synth_contacts")
    shard = Shard(mid1, mid2, count, chrom=rank)
    # ---- the genome: `world` chromosomes of nb bins each (fragment mid = i*R + R/2)
    nkeys = (nb - 1) * R // R + 1
    eng = PassEngine(R, N_BINS, 0, MAX_DIST, nkeys, dev)
    eng.set_fragments([nb] * world, [(nb - 1) * R] * world)
    tab = np.where((bias_host < 0.5) | (bias_host > 2), -1.0, bias_host)      # read_bias_file, fithic.py:147-149
    values = [np.zeros(0)] * world
    values[rank] = tab
    eng.set_bias(BiasTables(values, [R // 2] * world, dev))
    p = torch.empty((P + 1) & ~1, dtype=torch.float64, device=dev)[:P]
    q = torch.empty((P + 1) & ~1, dtype=torch.float64, device=dev)[:P]
    group = None

    genome_q = world > 1 and args.q_scope == "genome"

    def bh():
        if genome_q:      # all-reduce of the p histogram + all-gather of the candidate keys (2 host syncs)
            eng.qvalues_global(p, q, n_tests=-1, group=group, hist=eng.p_hist, prepared=True)
        else:             # K4 pre-filled q and listed the small p (bbk_pvalues_bh): no second pass over p
            eng.qvalues(p, q, n_tests=-1, use_hist=True, prepared=True)

    p_first = torch.empty((P + 1) & ~1, dtype=torch.float64, device=dev)[:P] if two_pass else None
    p_outlier = 1.0 / float(world * (nb * (K + 1) - K * (K + 1) // 2 - nb))   # 1 / possibleIntraInRangeCount (d = R..K*R)

    def bh_on(pp, qq):
        if genome_q:
            eng.qvalues_global(pp, qq, n_tests=-1, group=group, hist=eng.p_hist, prepared=True)
        else:
            eng.qvalues(pp, qq, n_tests=-1, use_hist=True, prepared=True)

    def step_on(sh, pp, qq):
        eng.hist([sh])
        eng.allreduce_stats(group)
        eng.fit()
        eng.p_hist.zero_()
        if two_pass:
            eng.pvalues(sh, p_first, with_hist=False)
            eng.hist_excluding([sh], [p_first], p_outlier)
            eng.allreduce_stats(group)
            eng.fit()
        eng.pvalues(sh, pp, with_hist=True, q_out=qq)
        bh_on(pp, qq)

    def step():
        step_on(shard, p, q)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    fit = eng.read_fit()
    eng.launches = 0
    step()
    launches_per_step = eng.launches + 1            # + the p_hist memset
    barrier()

    # ---- timed region: exactly K steps, device events, max over ranks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.mark_start()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    if sampler:
        sampler.mark_end()
    ms = e0.elapsed_time(e1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms = float(tms.item())
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    value = world * P / (ms_per_step * 1e-3)

    # ---- per-stage device times (separate instrumented steps; same stream, CUDA events)
    names = ["hist", "allreduce", "fit", "pvalues", "bh"]
    acc = dict((n, 0.0) for n in names)
    reps = min(args.steps, 5)
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        ev[0].record(); eng.hist([shard])
        ev[1].record(); eng.allreduce_stats(group)
        ev[2].record(); eng.fit()
        ev[3].record(); eng.p_hist.zero_(); eng.pvalues(shard, p, with_hist=True, q_out=q)
        ev[4].record(); bh()
        ev[5].record()
        torch.cuda.synchronize()
        for i, n in enumerate(names):
            acc[n] += ev[i].elapsed_time(ev[i + 1]) / reps
    fit_diag = eng.read_fit()          # phase cycles of the last instrumented fit (before the e2e section)
    peak, peak_src = _peaks()
    k4_gbs = K4_BYTES_PER_PAIR * P / (acc["pvalues"] * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one K4 launch on this exact workload, from the ncu --set full
    # capture summarised in profiles/r01_ncu_full_final.txt (1.1738 GB read + 1.5203 GB written: with the K4 -> K5
    # hand-over K4 also writes the 8 B/pair of q that the SURVEY's byte table books under K5)
    k4_traffic = 2.6941e9 if (nb == CHR1_BINS and not two_pass) else None
    roofline = {"bound": "hbm", "kernel": "pvalues_kernel (K4)", "achieved": k4_gbs, "peak": peak, "unit": "GB/s",
                "frac": k4_gbs / peak, "traffic": k4_traffic, "peak_source": peak_src,
                "traffic_note": "K4's DRAM bytes include the 8 B/pair of q it writes for K5 (hand-over); its own algorithmic bytes are 20 B/pair, read once and written once",
                "algorithmic_bytes_per_launch": K4_BYTES_PER_PAIR * P,
                "whole_pass_frac": (104 if two_pass else BYTES_PER_PAIR) * P / (ms_per_step * 1e-3) / 1e9 / peak,
                "stage_gbs": {"hist": 12 * P / (acc["hist"] * 1e-3) / 1e9, "pvalues": k4_gbs,
                              "bh": 16 * P / (acc["bh"] * 1e-3) / 1e9}}

    # ---- end to end: host (pinned) buffers in, p and q back to the host, every step, through engine.HostPipeline
    # (the public streamed call: two device slots, so step k+1's records arrive while step k's p/q leave)
    h_in = [torch.empty(P, dtype=torch.int32).pin_memory() for _ in range(3)]
    for h, d in zip(h_in, (mid1, mid2, count)):
        h.copy_(d)
    h_out = [(torch.empty(P, dtype=torch.float64).pin_memory(), torch.empty(P, dtype=torch.float64).pin_memory()) for _ in range(2)]
    torch.cuda.synchronize()
    pipe, e2e_mode = HostPipeline(eng, P, chrom=rank, slots=2), "pipelined across steps (2 device slots)"

    def e2e_step(k):
        h_p, h_q = h_out[k % 2]
        pipe.submit(h_in[0], h_in[1], h_in[2], h_p, h_q, run=step_on)

    def e2e_drain():
        if pipe is not None:
            pipe.drain()
        torch.cuda.synchronize()

    e2e_steps = max(2, min(args.steps, 10 if P < 200_000_000 else 3))     # the first inbound copy has nothing to overlap with: more steps amortise it
    e2e_step(0)
    e2e_step(1)
    e2e_drain()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    e2e_drain()
    # what a caller does with a pass whose outputs have arrived (here: the two the pipeline still holds): the fit status
    # (raises the reference's exception for a failed fit) and whether the reference's own s = min(y)**2 would have differed
    # from the kernel's (then that pass is submitted again with that s)
    e2e_resubmit = 0
    for k in range(pipe.submitted - 2, pipe.submitted):
        if PassEngine.reference_smoothing(pipe.fit_of(k)) is not None:
            e2e_resubmit += 1
    wall = time.perf_counter() - t0       # host clock around enqueue + drain: the copies run on three streams
    barrier()
    tms = torch.tensor([wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    e2e_value = world * P / (float(tms.item()) / e2e_steps * 1e-3)
    h_p, h_q = h_out[(e2e_steps - 1) % 2]
    kept = int((h_p <= 1).sum().item())
    sig = int((h_q <= 0.01).sum().item())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, pairs, wall_c, _ = cpu_baseline(1, nb=4500)
        cpu = {"value": v, "unit": "pairs/s", "cores": 1, "kind": "port",
               "sample": "one %d-bin block of the same workload (%.1fM records): histogram, binning, spline, bdtrc scoring, BH; "
                         "oracle/fithic_oracle.py on 1 of %d host cores, %.1f s" % (4500, pairs / 1e6, os.cpu_count() or 1, wall_c)}

    if rank == 0:
        t = eng.totals.cpu().numpy()
        line = {
            "metric": "fithic_contact_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("cfg4: chr1@1kb, all pairs within 2 Mb, two passes (refit after outlier removal)" if two_pass else
                                    "cfg5 share: 1/8 of the genome at 1 kb per GPU, all pairs within 2 Mb, single pass with q-values" if args.workload == "cfg5" else
                                    "cfg2: chr1@5kb, all pairs within 10 Mb, one chromosome-sized shard per GPU"),
                       "pairs_per_gpu": P, "bins_per_gpu": nb, "resolution": R, "max_dist": MAX_DIST, "n_bins": N_BINS,
                       "biases": True, "q_values": "genome-wide (histogram all-reduce + candidate all-gather)" if genome_q else "per shard", "l2": "inputs (%.2f GB/GPU) exceed the 126 MB L2" % (12 * P / 1e9),
                       "bytes_per_pair": 104 if two_pass else BYTES_PER_PAIR, "S": int(t[0]), "spline_knots": int(fit.n_knots),
                       "emitted_rows_rank0": kept, "q_le_0.01_rank0": sig},
            "stages_ms": acc,
            "fit_phase_cycles": dict(zip(["stage+boundaries", "bin_stats", "spline_search", "grid_eval", "pava+residual", "total"],
                                         [int(v) for v in fit_diag.phase_cycles])),
            "spline_diag": [int(v) for v in fit_diag.spline_diag],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": 12 * P, "d2h_bytes_per_step": 16 * P + int(eng.fit_result.numel()),
                    "steps": e2e_steps, "ms_per_step": float(tms.item()) / e2e_steps, "mode": e2e_mode,
                    "fit_checked_passes": 2, "smoothing_resubmits": e2e_resubmit},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        _emit(out_fd, line)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner there) get stderr instead."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(saved_fd, line):
    sys.stdout.flush()
    os.write(saved_fd, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bins", type=int, default=CHR1_BINS, help="bins of the per-GPU chromosome (default chr1 @ 5 kb)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg4", "cfg5"],
                    help="cfg2 (default, the measured config), cfg4 (chr1 at 1 kb, two passes) or cfg5 (one GPU's share of the genome at 1 kb)")
    ap.add_argument("--q-scope", default="genome", choices=["genome", "shard"],
                    help="N > 1: rank p-values across all ranks (default) or per shard")
    args = ap.parse_args()
    out_fd = _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args, out_fd)
    else:
        run_ours(args, out_fd)


if __name__ == "__main__":
    main()
