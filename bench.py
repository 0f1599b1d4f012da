#!/usr/bin/env python
"""Benchmark of the fairness-guidance path (BASELINE.json metric: guided images/sec, fwd+bwd loss path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C5|C1|C2|C3|C4]
                    [--scaling weak|strong] [--surface step|closures]

Workload (default C5 = BASELINE.json configs[4], the one the 1/2/4/8-GPU sweep is quoted on): 1024 synthetic 512x512 bf16
images PER GPU (weak scaling, the reference's own data parallelism: every rank generates its own images and the balanced
assignment runs over the gathered N = 1024 * gpus rows, E3:1978-2016), 1 face box each (5 % without a face), gender + race
+ age heads (K = 16 assignment classes, 75/25 age target), 100 Monte-Carlo draws per rank, threshold 0.2, full backward
into the image gradient.  With several ranks the line ALSO carries `strong_scaling`: the same 1024 images in total,
sharded (`--scaling strong` makes that the headline value instead).  A "step" is one pass of pipeline.GuidancePath.step
over the batch.  The classifier backbone (torchvision MobileNetV3 `features`, cuDNN) and CLIP/DINO are not part of the path
(SURVEY.md section 8d): their outputs / gradients enter as stand-in tensors of the right shape; `--backbone` additionally
times the step with the real torchvision backbone forward+backward in the loop (`with_backbone`).

One JSON line on stdout (rank 0):
  value        images/s with inputs resident in HBM (CUDA-graph replay of the step; K steps, CUDA events, max over ranks)
  e2e          the same step through the REFERENCE-FACING call surface -- the closures of fairguide.bind(), called in the
               order of the reference's own loop (tools/closure_loop.py, E3:1956-2147) -- with HOST (pinned) buffers: H2D of
               every input and D2H of loss + targets inside the timed region
  roofline     the dominant owned kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline the oracle's CPU restatement of the reference on a bounded sample of the workload, timed in the same run
  closures     latency of the reference's micro-batch loop at its real sizes (40 images per GPU, micro-batches of 4)
               on the closure surface, next to the same loop of the oracle on the host cores
  multi_gpu_parity (N > 1) "ok" after the ranks' targets were compared bit for bit with each other and with the oracle

`--impl reference` times the reference's CPU implementation of the path (the oracle port: per-image loops, 0-d tensor
histogram loop, per-element cost loop; ot.emd -> scipy linear_sum_assignment as BASELINE.md section 3 states) on the
host cores ON THE SAME CONFIG (all 1024 images of the workload per step, fp32) within a time budget; if a full step cannot
fit the budget it falls back to a sample and says so (`cap`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, global images, dtype, max_faces, description)
    "C1": ("gender", 64, "float32", 1, "exp-1 gender 2-class: 64 synthetic 512^2 images, 1 face each"),
    "C2": ("gender_race", 512, "float32", 1, "exp-3 gender x race joint balanced assignment, 512 images"),
    "C3": ("gender_race_age", 1024, "float32", 1, "exp-4 gender+race+age, 75/25 age target, 1024 images"),
    "C4": ("gender_race", 4096, "float32", 3, "exp-5 multi-concept prompts, up to 3 faces per image, 4096 images"),
    "C5": ("gender_race_age", 1024, "bfloat16", 1, "full backward into image gradient, 1024 images bf16"),
}
METRIC = "guided images/sec (fwd+bwd loss path)"


def bytes_per_image(e, H=512, W=512, s=224):
    """fwd reads the image once and writes chip + small; bwd reads both
    gradients and writes the dense image gradient."""
    fwd = (3 * H * W + 2 * 3 * s * s) * e
    bwd = (2 * 3 * s * s + 3 * H * W) * e
    return fwd, bwd


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


class EventProbe:
    """CUDA events on the launching stream around named stages of the step."""

    def __init__(self, torch):
        self.torch, self.pairs, self.open = torch, {}, {}
        self.enabled = False

    def begin(self, name):
        if self.enabled:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self.open[name] = e

    def end(self, name):
        if self.enabled:
            e = self.torch.cuda.Event(enable_timing=True)
            e.record()
            self.pairs.setdefault(name, []).append((self.open.pop(name), e))

    def mean_ms(self, name):
        p = self.pairs.get(name, [])
        return sum(a.elapsed_time(b) for a, b in p) / len(p) if p else None

    def reset(self):
        self.pairs, self.open = {}, {}


def ncu_traffic(kernel, n_local, dtype_name):
    """dram read + write bytes of one launch of `kernel` from the committed `ncu --set full` capture (the newest
    profiles/*_traffic.json, taken on workload C5 at 1 GPU); None for any other workload."""
    if n_local != 1024 or dtype_name != "bfloat16":
        return None
    here = os.path.dirname(os.path.abspath(__file__))
    files = sorted(f for f in os.listdir(os.path.join(here, "profiles")) if f.endswith("_traffic.json"))
    if not files:
        return None
    d = json.load(open(os.path.join(here, "profiles", files[-1])))
    for name, launches in d["kernels"].items():
        if kernel in name:
            l = launches[-1]
            return l["dram_read_bytes"] + l["dram_write_bytes"]
    return None


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def pin_to_gpu_numa(torch, local_rank):
    """Bind this rank (and therefore the pinned host buffers it allocates next: first touch) to the CPU cores of the NUMA
    node its GPU hangs off.  With every rank on node 0 the end-to-end number stopped scaling at 4 GPUs (all H2D copies
    crossed one memory controller / one socket link).  Best effort: returns a short description or the reason it was skipped."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{dev}/numa_node").read().strip())
        if node < 0:
            return f"gpu {dev}: no NUMA affinity reported"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return f"gpu {dev}: node {node} has no allowed cpu"
        os.sched_setaffinity(0, allowed)
        return f"gpu {dev} -> numa node {node} ({len(allowed)} cpus)"
    except Exception as ex:                                            # noqa: BLE001
        return f"skipped: {type(ex).__name__}: {str(ex)[:80]}"


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(kind, n_images, solver, seed=5991):
    """One step of the oracle's CPU restatement of the reference path on `n_images` images of the workload (fp32)."""
    import torch
    from fairguide import pipeline
    from oracle import emd as oemd, pipeline as opipe
    cfg = pipeline.GuidanceConfig(kind=kind)
    batch = pipeline.synth_batch(n_images, cfg, torch.float32, "cpu", seed=seed, host=True)
    head = pipeline.make_head_weights(cfg, torch.float32, "cpu")
    emd = {"lsa": oemd.emd_lsa, "c": oemd.emd_c}[solver]

    def run():
        tm = {}
        t0 = time.perf_counter()
        opipe.step(batch, cfg, head, literal=True, timings=tm, emd=emd)
        return time.perf_counter() - t0, tm
    return run


def cpu_reference_run(kind, n_workload, steps, budget_s, solver="lsa", sample=None):
    """The reference arm / cpu_baseline: the oracle port on all host threads.  `sample` = images per step (None: the whole
    workload if a calibration run predicts that `steps` (at least one) full steps fit `budget_s`, else the largest sample
    that does, with the cap stated)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    note = {"lsa": "ot.emd -> scipy.optimize.linear_sum_assignment on the column-replicated cost matrix (BASELINE.md section 3)",
            "c": "ot.emd -> the oracle's exact C row-insertion solver"}[solver]
    t_cal, _ = cpu_reference_step(kind, 32, solver)()                    # warm-up + calibration: 32 images
    cap = None
    if sample is None:
        # The literal per-image loops make the backward SUPER-linear in the batch (every `images[i]` selected inside the loop
        # back-propagates a full [n,3,512,512] zero tensor: measured t ~ n^1.8, 2.4 / 14.8 / 78.8 s for 32 / 96 / 256 images on
        # 8 cores), and so is the Monte-Carlo assignment.  A second calibration point fixes the exponent; the step is sized so
        # that its PREDICTED duration uses 60 % of the budget (the exponent still grows slowly with n).
        import math
        t_cal2, _ = cpu_reference_step(kind, 96, solver)()
        expo = min(2.5, max(1.0, math.log(max(t_cal2, 1e-6) / max(t_cal, 1e-6)) / math.log(3.0)))
        predict = lambda n: t_cal2 * (n / 96.0) ** expo
        room = 0.6 * budget_s
        est_full = predict(n_workload)
        if est_full <= room:
            sample = n_workload
            steps = max(1, min(steps, int(room // est_full)))
        else:
            sample = int(96.0 * (room / t_cal2) ** (1.0 / expo)) // 32 * 32
            sample = max(32, min(sample, n_workload))
            steps = 1
            cap = (f"a full {n_workload}-image step was predicted at {est_full:.0f} s (t ~ n^{expo:.2f} from 32- and 96-image runs) > 60 % of the "
                   f"{budget_s:.0f} s budget: {sample}-image sample")
    run = cpu_reference_step(kind, sample, solver)
    times, stages = [], {}
    for _ in range(steps):
        dt, tm = run()
        times.append(dt)
        for k, v in tm.items():
            stages[k] = stages.get(k, 0.0) + v / steps
    t = sum(times) / len(times)
    out = {"value": sample / t, "unit": "images/s", "cores": cores, "kind": "port",
           "sample": f"{sample} of the workload's {n_workload} images per step, {steps} timed step(s) after a 32-image warm-up step "
                     f"(oracle/pipeline.py: the reference's literal loops, fp32, {note})",
           "images_per_step": sample, "same_config": sample == n_workload, "steps": steps,
           "seconds_per_step": t, "stage_seconds": {k: round(v, 4) for k, v in stages.items()}}
    if cap:
        out["cap"] = cap
    return out


# ----------------------------------------------------------------------------------------------- main
_JSON_FD = None


def own_stdout():
    """Rank 0 prints ONE JSON line: everything else that writes to file descriptor 1 during the run (NCCL's version banner,
    library chatter from C code that Python's sys.stdout redirection cannot see) is sent to stderr, and the line goes to the
    saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--backbone", action="store_true", help="also time the step with the torchvision backbone in the loop")
    ap.add_argument("--cpu-sample", type=int, default=96, help="images per step of the cpu_baseline leg of the default run")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: seconds available for the timed CPU steps")
    ap.add_argument("--ref-solver", default="lsa", choices=["lsa", "c"], help="--impl reference: exact solver standing in for ot.emd")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every launch of a step from Python instead of replaying a CUDA graph")
    ap.add_argument("--exchange", default="auto", choices=["auto", "nccl"],
                    help="N > 1: auto = one-shot NVLink stores between symmetric buffers (dist.PeerExchange) when the GPUs can map each "
                         "other's memory, nccl = one NCCL call per exchange between three graph replays")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's image count per GPU (default); strong: that count in total, sharded")
    ap.add_argument("--profile-only", action="store_true",
                    help="for ncu: time only the eager (launch by launch) steps between cudaProfilerStart/Stop, print a short line, exit")
    ap.add_argument("--surface", default="step", choices=["step", "closures"],
                    help="closures: `value` is measured through the fairguide.bind() closure surface too (eager, no graph)")
    a = ap.parse_args()
    kind, n_workload, dtype_name, max_faces, desc = WORKLOADS[a.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_local = n_workload if a.scaling == "weak" else n_workload // world
    n_global = n_local * world

    if a.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(kind, n_workload, max(1, a.steps), a.ref_budget_s, solver=a.ref_solver)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": a.gpus,
                "steps": r["steps"], "warmup": 1, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True,
                "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{a.workload}: {desc}", "kind": kind, "global_batch": n_workload, "images_per_step": r["images_per_step"],
                           "same_config": r["same_config"],
                           "note": ("CPU arm: " + ("the whole workload" if r["same_config"] else f"{r['images_per_step']} of the workload's {n_workload} images")
                                    + " on the host cores in fp32 (the GPU arm computes C5 in bf16); the literal per-image loops are super-linear in "
                                    f"the batch, so requested --steps {a.steps} --warmup {a.warmup} are clamped to ONE step sized for the "
                                    f"{a.ref_budget_s:.0f} s budget" + (": " + r["cap"] if r.get("cap") else ""))},
                "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    own_stdout()                                   # NCCL still prints its version banner to fd 1 at WARN: keep stdout for the ONE JSON line
    import torch
    import torch.distributed as dist
    import fairguide
    from fairguide import pipeline, _lib
    from tools import closure_loop

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_note = pin_to_gpu_numa(torch, local_rank) if world > 1 else "single rank: not pinned"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus or world == 1, (world, a.gpus)
    dtype = getattr(torch, dtype_name)
    esize = torch.empty((), dtype=dtype).element_size()
    cfg = pipeline.GuidanceConfig(kind=kind)
    head = pipeline.make_head_weights(cfg, dtype, dev)
    probe = EventProbe(torch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        probe.enabled = True
        c0 = dict(_lib.CALLS)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.profiler.start()          # `ncu --profile-from-start off` captures the timed steps only
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.profiler.stop()
        barrier()
        probe.enabled = False
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        calls = {k: _lib.CALLS[k] - c0.get(k, 0) for k in _lib.CALLS}
        return ms / steps, calls

    def global_faces(batch):
        # N (faces in the global batch) is host knowledge in the reference (the detector runs on the host)
        counts_all = batch["counts"].clone()
        if world > 1:
            gathered = [torch.empty_like(counts_all) for _ in range(world)]
            dist.all_gather(gathered, counts_all)
            counts_all = torch.cat(gathered)
        return int((counts_all > 0).sum().item())

    # ---- correctness of the multi-rank step on this hardware, before anything is timed
    multi_gpu_parity = None
    if world > 1:
        from tests import stepcheck                       # the oracle as the checker, outside every timed region
        rep = stepcheck.check_multi_rank(kind, dtype, dev, rank, world, n_side_global=256, S=cfg.num_samples_per_device, captured=not a.no_graph)
        multi_gpu_parity = {"status": "ok", "checked": "targets_all and plan counts bit-identical on all ranks; gathered rows in rank order; "
                            "targets equal to the oracle's simulated world (same draws per rank); graph replay == eager", **rep}

    def measure(n_loc, steps, warmup, want_stages=True):
        """Device-resident timing of one configuration -> dict (ms_per_step = graph replay, eager stages, launches)."""
        batch = pipeline.synth_batch_device(n_loc, cfg, dtype, dev, seed=5991 + rank, max_faces=max_faces)
        path = pipeline.GuidancePath(cfg, head, peer_exchange=("auto" if a.exchange == "auto" else False))
        nv = global_faces(batch)
        res = {}
        captured, note = None, "off (--no-graph)"
        if not a.no_graph and not a.profile_only:
            try:
                captured = pipeline.CapturedStep(path, batch, nv)
                note = "on"
            except Exception as ex:                                    # noqa: BLE001  (keep measuring, say why)
                captured, note = None, f"capture failed, eager launches: {type(ex).__name__}: {str(ex)[:120]}"
                torch.cuda.synchronize()
        probe.reset()
        ms_eager, calls = timed(lambda: res.__setitem__("out", path.step(batch, num_valid=nv, probe=probe)), max(3, steps // 2), warmup)
        launches_per_step = sum(_lib.KERNELS_PER_CALL.get(k, 1) * v for k, v in calls.items()) // max(3, steps // 2)
        stage_ms = {k: probe.mean_ms(k) for k in ("sample_fwd", "assign", "image_grad")}
        if world > 1:      # where the assignment stage's time goes on this rank (eager launches; waits include the slowest peer)
            stage_ms["assign_parts"] = {k: probe.mean_ms(k) for k in ("exchange_rows", "plan_counts", "exchange_counts")}
        if a.profile_only:
            if rank == 0:
                emit({"profile_only": True, "workload": a.workload, "ms_per_step_eager": ms_eager, "stage_ms": stage_ms,
                      "gpu_launches_per_step": launches_per_step})
            if world > 1:
                dist.destroy_process_group()
            sys.exit(0)
        ms_step, _ = timed(lambda: res.__setitem__("out", captured.replay() if captured is not None else path.step(batch, num_valid=nv)),
                           steps, warmup)
        pipeline.validate_step(res["out"])                            # status words of the assignment kernels (frozen face count, exact plans)
        return dict(batch=batch, path=path, nv=nv, out=res["out"], captured=captured, graph=note, ms_eager=ms_eager, ms_step=ms_step,
                    stage_ms=stage_ms, launches_per_step=launches_per_step)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    m = measure(n_local, a.steps, a.warmup)
    clocks = sampler.summary() if sampler else None
    batch, path, num_valid, captured = m["batch"], m["path"], m["nv"], m["captured"]
    ms_step, ms_eager, stage_ms, launches_per_step = m["ms_step"], m["ms_eager"], m["stage_ms"], m["launches_per_step"]
    value = n_global / (ms_step * 1e-3)
    if world > 1:
        # every rank must hold the same global targets after the timed steps as well
        for t in m["out"]["targets_all"]:
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t.contiguous())
            assert all(torch.equal(p, t) for p in parts), "targets_all differs between ranks after the timed steps"

    # ---- the reference-facing closure surface (eager): device-resident and end to end with host buffers
    step_closure = closure_loop.ClosureStep(fairguide, cfg, head, dtype, micro_batch=None, literal=True)
    surface = None
    if a.surface == "closures":
        ms_cl, _ = timed(lambda: step_closure(batch, num_valid=num_valid), max(3, a.steps // 2), 3)
        step_fused = closure_loop.ClosureStep(fairguide, cfg, head, dtype, micro_batch=None, literal=False)
        ms_cf, _ = timed(lambda: step_fused(batch, num_valid=num_valid), max(3, a.steps // 2), 3)
        surface = {"literal_closures": {"ms_per_step": ms_cl, "value": n_global / (ms_cl * 1e-3)},
                   "fused_entry_points": {"ms_per_step": ms_cf, "value": n_global / (ms_cf * 1e-3)},
                   "note": "device-resident, eager (no CUDA graph): tools/closure_loop.py, the reference's call order on fairguide.bind()"}

    e2e = None
    if not a.no_e2e:
        keys = [k for k, v in batch.items() if torch.is_tensor(v) or isinstance(v, list)]
        pin = lambda v: v.cpu().pin_memory()
        host = {k: (pin(batch[k]) if torch.is_tensor(batch[k]) else [pin(x) for x in batch[k]]) for k in keys}
        stage = {k: (torch.empty_like(batch[k]) if torch.is_tensor(batch[k]) else [torch.empty_like(x) for x in batch[k]]) for k in keys}
        stage["k_head"] = batch["k_head"]
        h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v)) + \
            sum(x.numel() * x.element_size() for v in host.values() if isinstance(v, list) for x in v)
        out_host = {"loss": torch.empty((), dtype=torch.float32).pin_memory(),
                    "targets": [torch.empty(n_local, dtype=torch.int64).pin_memory() for _ in range(cfg.n_attr)]}
        d2h = 4 + 8 * n_local * cfg.n_attr

        def upload():
            for k in keys:
                if torch.is_tensor(host[k]):
                    stage[k].copy_(host[k], non_blocking=True)
                else:
                    for d, s in zip(stage[k], host[k]):
                        d.copy_(s, non_blocking=True)

        def download(out):
            out_host["loss"].copy_(out["loss_mean"], non_blocking=True)
            for d, s in zip(out_host["targets"], out["targets"]):
                d.copy_(s, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        def step_e2e_closures():
            upload()
            download(step_closure(stage, num_valid=num_valid))

        ms_e2e, _ = timed(step_e2e_closures, max(3, a.steps // 2), 3)
        e2e = {"value": n_global / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "ms_per_step": ms_e2e,
               "surface": "fairguide.bind() closures called in the reference's order (tools/closure_loop.py; E3:1956-2147), pinned host "
                          "buffers -> H2D -> closures -> D2H of loss and targets, every step"}
        # the same with the product's batched step (one CUDA-graph replay) between the copies
        upload()
        torch.cuda.synchronize()
        cap_e2e = None
        if captured is not None:
            try:
                cap_e2e = pipeline.CapturedStep(path, stage, num_valid)      # the staging buffers are the graph's inputs
            except Exception:                                              # noqa: BLE001
                cap_e2e = None
                torch.cuda.synchronize()

        def step_e2e_fused():
            upload()
            download(cap_e2e.replay() if cap_e2e is not None else path.step(stage, num_valid=num_valid))

        ms_e2f, _ = timed(step_e2e_fused, max(3, a.steps // 2), 3)
        e2e["batched_step"] = {"value": n_global / (ms_e2f * 1e-3), "ms_per_step": ms_e2f,
                               "note": "pipeline.GuidancePath.step (graph replay) instead of the closure sequence, same copies"}
        del cap_e2e, host, stage

    # ---- the reference's micro-batch regime on the closure surface: 40 images per GPU, micro-batches of 4 (E3:2088)
    closures = None
    if world == 1:
        try:
            n_mb, mb = 40, 4
            small_batch = pipeline.synth_batch_device(n_mb, cfg, dtype, dev, seed=77, max_faces=max_faces)
            nv_mb = int((small_batch["counts"] > 0).sum().item())
            loop = closure_loop.ClosureStep(fairguide, cfg, head, dtype, micro_batch=mb, literal=True)
            ms_loop, _ = timed(lambda: loop(small_batch, num_valid=nv_mb), 10, 3)
            loop_f = closure_loop.ClosureStep(fairguide, cfg, head, dtype, micro_batch=mb, literal=False)
            ms_loop_f, _ = timed(lambda: loop_f(small_batch, num_valid=nv_mb), 10, 3)
            closures = {"images_per_gpu": n_mb, "micro_batch": mb, "ms_per_step": ms_loop, "images_per_s": n_mb / (ms_loop * 1e-3),
                        "ms_per_step_fused_entry_points": ms_loop_f,
                        "note": "fairguide.bind() closures in the reference's micro-batch loop (E3:2088-2147), eager, device-resident"}
            if not a.no_cpu_baseline:
                run = cpu_reference_step(kind, n_mb, "lsa", seed=77)
                run()
                t_cpu, _ = run()
                closures["cpu_port_ms_per_step"] = t_cpu * 1e3
                closures["cpu_port_note"] = "oracle/pipeline.py on the same 40 images (per-image loops; one backward for the batch), all host cores"
        except Exception as ex:                                        # noqa: BLE001
            closures = {"error": f"{type(ex).__name__}: {str(ex)[:160]}"}
            torch.cuda.synchronize()

    # ---- fairness-only sub-path: crop without the resize, chip gradient without the semantics branch
    fairness_only = None
    try:
        if world > 1:
            raise NotImplementedError
        from fairguide import ops as _ops
        o = m["out"]
        boxes_f, ind_f = o["boxes"], o["indicators"]

        def step_fair():
            _ops.crop_resize_fwd(batch["images"], boxes_f, ind_f, (cfg.size_face,) * 2, None, cfg.fill_value)
            _ops.image_grad(batch["g_chips"], None, boxes_f, ind_f, None, None, tuple(batch["images"].shape), dtype, dev)

        ms_fair, _ = timed(step_fair, max(3, a.steps // 2), 3)
        side = (boxes_f[:, 2] - boxes_f[:, 0]).clamp(min=0).double()
        box_px = float((torch.where(ind_f, side * side, torch.zeros_like(side))).sum().item())       # sum of s_i^2 over the faces
        alg_fair = (3 * box_px + n_local * (3 * cfg.size_face ** 2 * 2 + 3 * 512 * 512)) * esize
        fairness_only = {"ms": ms_fair, "GBps": alg_fair / (ms_fair * 1e-3) / 1e9,
                         "algorithmic_bytes": alg_fair, "note": "crop forward + chip-gradient backward only (no resize / semantics branch)"}
    except NotImplementedError:
        fairness_only = {"note": "reported at 1 GPU only (rank-local kernels)"}
    except Exception as ex:                                        # noqa: BLE001
        fairness_only = {"error": f"{type(ex).__name__}: {str(ex)[:100]}"}
        torch.cuda.synchronize()

    with_backbone = None
    if a.backbone:
        import torchvision
        net = torchvision.models.mobilenet_v3_large(weights=None).to(dev, dtype).eval().to(memory_format=torch.channels_last)
        for p in net.parameters():
            p.requires_grad_(False)

        def step_backbone():
            # forward of the backbone on the chips, backward from the head's g_pooled to the chip gradient
            out = path.step(batch, num_valid=num_valid)
            chips = out["chips"].detach().requires_grad_(True)
            with torch.enable_grad():
                pooled = torch.flatten(net.avgpool(net.features(chips.contiguous(memory_format=torch.channels_last))), 1)
                pooled.backward(out["g_pooled"])
            return chips.grad

        ms_bb, _ = timed(step_backbone, max(3, a.steps // 4), 2)
        with_backbone = {"value": n_global / (ms_bb * 1e-3), "unit": "images/s", "ms_per_step": ms_bb,
                         "note": "owned step + torchvision mobilenet_v3_large features fwd+bwd (cuDNN) on the chips"}

    # ---- the other scaling mode, so that both are on record for every N > 1 (the driver computes the efficiencies)
    other_scaling = None
    if world > 1:
        other = "strong" if a.scaling == "weak" else "weak"
        n_other = n_workload // world if other == "strong" else n_workload
        mo = measure(n_other, a.steps, a.warmup)
        slow = max((k for k in mo["stage_ms"] if k != "assign_parts"), key=lambda k: mo["stage_ms"][k] or 0.0)
        other_scaling = {"scaling": other, "global_batch": n_other * world, "images_per_gpu": n_other, "ms_per_step": mo["ms_step"],
                         "value": n_other * world / (mo["ms_step"] * 1e-3), "ms_per_step_eager": mo["ms_eager"], "stage_ms": mo["stage_ms"],
                         "largest_stage": slow, "cuda_graph": mo["graph"],
                         "note": "strong = BASELINE C5's sweep (the workload's images IN TOTAL, sharded over the GPUs); "
                                 "weak = the workload's images per GPU"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_src = peaks()
    fwd_b, bwd_b = bytes_per_image(esize)
    dom = "image_grad" if (stage_ms["image_grad"] or 0) >= (stage_ms["sample_fwd"] or 0) else "sample_fwd"
    alg = (bwd_b if dom == "image_grad" else fwd_b) * n_local
    ach = alg / (stage_ms[dom] * 1e-3) / 1e9
    kname = {"image_grad": "image_grad_staged_kernel" if esize == 2 else "image_grad_tiled_kernel", "sample_fwd": "sample_fwd_tiled_kernel"}[dom]
    roofline = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": ncu_traffic(kname, n_local, dtype_name), "peak_source": pk_src,
                "algorithmic_bytes_per_launch": alg,
                "kernels": {k: {"ms": stage_ms[k], "GBps": ((fwd_b if k == "sample_fwd" else bwd_b) * n_local / (stage_ms[k] * 1e-3) / 1e9)}
                            for k in ("sample_fwd", "image_grad") if stage_ms[k]},
                "whole_step_frac": (fwd_b + bwd_b) * n_local / (ms_step * 1e-3) / 1e9 / pk["hbm_gbs"]}
    cpu = None
    if not a.no_cpu_baseline and world == 1:
        cpu = cpu_reference_run(kind, n_workload, 2, 60.0, solver="lsa", sample=min(a.cpu_sample, n_workload))
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": {"bfloat16": "bf16", "float32": "f32", "float16": "f16"}[dtype_name], "data": "synthetic",
            "config": {"workload": f"{a.workload}: {desc}", "kind": kind, "global_batch": n_global, "images_per_gpu": n_local,
                       "image": "3x512x512", "chip": "3x224x224", "mc_draws_per_rank": cfg.num_samples_per_device,
                       "faces_in_batch": num_valid, "parallelism": f"dp{world}", "numa": numa_note,
                       "backbone": "excluded (stand-in pooled features and chip gradient; SURVEY.md 8d)",
                       "l2": f"inputs larger than L2 ({(fwd_b + bwd_b) * n_local / 1e6:.0f} MB touched per step per GPU vs 126 MB)"},
            "stage_ms": stage_ms, "ms_per_step_eager": ms_eager, "cuda_graph": m["graph"], "gpu_launches": launches_per_step * a.steps,
            "gpu_launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "fairness_only": fairness_only, "e2e": e2e, "closures": closures, "cpu_baseline": cpu}
    if surface:
        line["surface"] = surface
    if world > 1:
        line["exchange"] = ("peer: one-shot NVLink stores / loads between symmetric buffers with epoch flags (csrc/fg_peer.cu), the step is ONE CUDA graph"
                            if path.peer else "nccl: all_gather_into_tensor + all_reduce between three graph replays")
    if multi_gpu_parity:
        line["multi_gpu_parity"] = multi_gpu_parity["status"]
        line["multi_gpu_parity_detail"] = multi_gpu_parity
    if other_scaling:
        line["strong_scaling" if other_scaling["scaling"] == "strong" else "weak_scaling"] = other_scaling
    if with_backbone:
        line["with_backbone"] = with_backbone
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
