#!/usr/bin/env python
"""Benchmark of the fairness-guidance path (BASELINE.json metric: guided images/sec, fwd+bwd loss path).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C5|C2|C3|C4|C1]

Workload (default C5 = BASELINE.json configs[4], the one the 1/2/4/8-GPU sweep is quoted on):
1024 synthetic 512x512 bf16 images PER GPU (weak scaling, the reference's own data parallelism: every
rank generates its own `train_images_per_prompt_GPU` images and the balanced assignment runs over the
gathered N = 1024 * gpus rows, E3:1978-2016; `--scaling strong` shards 1024 images in total instead), 1 face
box each (5 % without a face), gender+race+age heads (K = 16 assignment classes, 75/25 age target),
100 Monte-Carlo draws per rank, threshold 0.2, full backward into the image gradient.  A "step" is one
pass of pipeline.GuidancePath.step over that batch.  The classifier backbone (torchvision MobileNetV3
`features`, cuDNN) and CLIP/DINO are not part of the path: their outputs /
gradients enter as stand-in tensors of the right shape; `--backbone` additionally times the step with
the real torchvision backbone forward+backward in the loop and reports it as `with_backbone`.

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
step called with HOST (pinned) buffers, H2D of the inputs and D2H of loss + targets inside the timed
region; `roofline` = the dominant owned kernel against the measured HBM copy bandwidth;
`cpu_baseline` = the oracle's CPU restatement of the reference on a bounded sample of the workload.
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


def ncu_traffic(kernel, n_local, dtype_name):
    """dram read + write bytes of one launch of `kernel` from the committed `ncu --set full` capture (profiles/*_traffic.json,
    taken on workload C5 at 1 GPU); None for any other workload."""
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


def cpu_reference_run(kind, n_sample, steps, warmup, seed=5991):
    """The oracle's CPU restatement of the reference path (per-image loops, 0-d tensor histogram loop,
    per-element cost loop; ot.emd replaced by the oracle's exact C solver), all host threads."""
    import torch
    from fairguide import pipeline
    from oracle import pipeline as opipe
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = pipeline.GuidanceConfig(kind=kind)
    batch = pipeline.synth_batch(n_sample, cfg, torch.float32, "cpu", seed=seed, host=True)
    head = pipeline.make_head_weights(cfg, torch.float32, "cpu")
    times, stages = [], {}
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        tm = {}
        opipe.step(batch, cfg, head, literal=True, timings=tm)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
            for k, v in tm.items():
                stages[k] = stages.get(k, 0.0) + v / steps
    t = sum(times) / len(times)
    return {"value": n_sample / t, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{n_sample} images of the same workload per step, {steps} steps after {warmup} warm-up "
                      f"(oracle/pipeline.py: literal reference loops, fp32, ot.emd -> oracle C solver)",
            "seconds_per_step": t, "stage_seconds": {k: round(v, 4) for k, v in stages.items()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", choices=sorted(WORKLOADS))
    ap.add_argument("--backbone", action="store_true", help="also time the step with the torchvision backbone in the loop")
    ap.add_argument("--cpu-sample", type=int, default=96, help="images per step of the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every launch of a step from Python instead of replaying a CUDA graph")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's image count per GPU (default); strong: that count in total, sharded")
    a = ap.parse_args()
    kind, n_workload, dtype_name, max_faces, desc = WORKLOADS[a.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_ranks = world
    n_local = n_workload if a.scaling == "weak" else n_workload // n_ranks
    n_global = n_local * n_ranks

    if a.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
        r = cpu_reference_run(kind, a.cpu_sample, steps, warmup)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": a.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True,
                "scaling": a.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{a.workload}: {desc}", "kind": kind, "global_batch": n_global, "images_per_gpu": n_local,
                           "sample_images_per_step": a.cpu_sample},
                "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    import torch
    import torch.distributed as dist
    import fairguide
    from fairguide import pipeline, _lib

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus or world == 1, (world, a.gpus)
    dtype = getattr(torch, dtype_name)
    esize = torch.empty((), dtype=dtype).element_size()
    cfg = pipeline.GuidanceConfig(kind=kind)
    head = pipeline.make_head_weights(cfg, dtype, dev)
    batch = pipeline.synth_batch_device(n_local, cfg, dtype, dev, seed=5991 + rank, max_faces=max_faces)
    path = pipeline.GuidancePath(cfg, head)
    # N (faces in the global batch) is host knowledge in the reference (the detector runs on the host)
    counts_all = batch["counts"].clone()
    if world > 1:
        gathered = [torch.empty_like(counts_all) for _ in range(world)]
        dist.all_gather(gathered, counts_all)
        counts_all = torch.cat(gathered)
    num_valid = int((counts_all > 0).sum().item())
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

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    result = {}

    def step_eager():
        result["out"] = path.step(batch, num_valid=num_valid, probe=probe)

    # the timed step: one CUDA-graph replay of GuidancePath.step (the same launches, enqueued by one driver call); the eager
    # loop below it supplies the per-stage CUDA events and the launch count (events cannot be read back from a graph)
    captured, graph_note = None, "off (--no-graph)"
    if not a.no_graph:
        try:
            captured = pipeline.CapturedStep(path, batch, num_valid)
            graph_note = "on"
        except Exception as ex:                                    # noqa: BLE001  (keep measuring, say why)
            captured, graph_note = None, f"capture failed, eager launches: {type(ex).__name__}: {str(ex)[:120]}"
            torch.cuda.synchronize()

    def step_resident():
        result["out"] = captured.replay() if captured is not None else path.step(batch, num_valid=num_valid)

    ms_eager, calls = timed(step_eager, max(3, a.steps // 2), a.warmup)
    launches_per_step = (sum(_lib.KERNELS_PER_CALL.get(k, 1) * v for k, v in calls.items())
                         + calls.get("fg_ot_plan_counts", 0) * _lib.ot_levels(num_valid)) // max(3, a.steps // 2)
    stage_ms = {k: probe.mean_ms(k) for k in ("sample_fwd", "assign", "image_grad")}
    ms_step, _ = timed(step_resident, a.steps, a.warmup)
    clocks = sampler.summary() if sampler else None
    launches = launches_per_step * a.steps
    value = n_global / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers
    e2e = None
    if not a.no_e2e:
        host = {k: (v.cpu().pin_memory() if torch.is_tensor(v) else [x.cpu().pin_memory() for x in v] if isinstance(v, list) else v)
                for k, v in batch.items() if k != "k_head"}
        host["k_head"] = batch["k_head"]
        stage = {k: (torch.empty_like(batch[k]) if torch.is_tensor(batch[k]) else [torch.empty_like(x) for x in batch[k]])
                 for k in host if k != "k_head"}
        stage["k_head"] = batch["k_head"]
        h2d = sum(v.numel() * v.element_size() for v in host.values() if torch.is_tensor(v)) + \
            sum(x.numel() * x.element_size() for x in host["preds_ori"])
        out_host = {"loss": torch.empty((), dtype=torch.float32).pin_memory(),
                    "targets": [torch.empty(n_local, dtype=torch.int64).pin_memory() for _ in range(cfg.n_attr)]}
        d2h = 4 + 8 * n_local * cfg.n_attr

        captured_e2e = None
        if captured is not None:
            for k, v in host.items():                              # real values in the staging buffers before they are captured
                if torch.is_tensor(v):
                    stage[k].copy_(v)
                elif isinstance(v, list):
                    for d, s_ in zip(stage[k], v):
                        d.copy_(s_)
            torch.cuda.synchronize()
            try:
                captured_e2e = pipeline.CapturedStep(path, stage, num_valid)      # the staging buffers are the graph's inputs
            except Exception:                                      # noqa: BLE001
                captured_e2e = None
                torch.cuda.synchronize()

        def step_e2e():
            for k, v in host.items():
                if torch.is_tensor(v):
                    stage[k].copy_(v, non_blocking=True)
                elif isinstance(v, list):
                    for d, s in zip(stage[k], v):
                        d.copy_(s, non_blocking=True)
            out = captured_e2e.replay() if captured_e2e is not None else path.step(stage, num_valid=num_valid)
            out_host["loss"].copy_(out["loss_mean"], non_blocking=True)
            for d, s in zip(out_host["targets"], out["targets"]):
                d.copy_(s, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        ms_e2e, _ = timed(step_e2e, max(3, a.steps // 2), 3)
        e2e = {"value": n_global / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "ms_per_step": ms_e2e}

    # ---- fairness-only sub-path: crop without the resize, chip gradient without the semantics branch
    fairness_only = None
    try:
        if world > 1:
            raise NotImplementedError
        from fairguide import ops as _ops
        o = result["out"]
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
        cpu = cpu_reference_run(kind, a.cpu_sample, 2, 1)
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": {"bfloat16": "bf16", "float32": "f32", "float16": "f16"}[dtype_name], "data": "synthetic",
            "config": {"workload": f"{a.workload}: {desc}", "kind": kind, "global_batch": n_global, "images_per_gpu": n_local,
                       "image": "3x512x512", "chip": "3x224x224", "mc_draws_per_rank": cfg.num_samples_per_device,
                       "faces_in_batch": num_valid, "parallelism": f"dp{world}",
                       "backbone": "excluded (stand-in pooled features and chip gradient; SURVEY.md 8d)",
                       "l2": f"inputs larger than L2 ({(fwd_b + bwd_b) * n_local / 1e6:.0f} MB touched per step per GPU vs 126 MB)"},
            "stage_ms": stage_ms, "ms_per_step_eager": ms_eager, "cuda_graph": graph_note, "gpu_launches": launches, "gpu_launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "fairness_only": fairness_only, "e2e": e2e, "cpu_baseline": cpu}
    if with_backbone:
        line["with_backbone"] = with_backbone
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
