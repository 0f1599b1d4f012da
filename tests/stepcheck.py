"""Checker shared by ``__graft_entry__.smoke()`` and the ``-m gpu`` tests: runs ``GuidancePath.step`` on the GPU and
compares it, stage by stage, with the oracle (``oracle/``).  Test infrastructure -- nothing in the product imports it.

Two checks:

* ``check_small_steps``   -- the three experiment kinds end to end on 12 fp32 images of 256x256 (every stage, incl. the
  head, from the same fp32 inputs on both sides).
* ``check_step_vs_oracle`` -- one step at a BASELINE shape (512x512 images, 224x224 crops, any dtype / batch), compared
  piecewise so that each bar of BASELINE.json applies to the stage it names:
    boxes                       bit-exact
    head logits                 within ``rtol`` of the fp32 torch reference of the same weights
    probabilities / argmax      the oracle's softmax of the DEVICE logits (<= 1 unit in the last place; preds exact)
    target classes              bit-exact against the oracle's assignment of the DEVICE probabilities and the same draws
                                (``literal=False``: same arithmetic, vectorised), before and after thresholding
    losses                      within ``rtol``
    crops, resized images and the image gradient, on ``n_grad`` images, within ``rtol`` (gradient: of its largest entry)
"""
import numpy as np
import torch

RTOL = {torch.float32: 1e-3, torch.bfloat16: 2e-2, torch.float16: 5e-3}      # BASELINE.json north_star: 1e-3 fp32, 2e-2 bf16


def _to(v, device):
    if torch.is_tensor(v):
        return v.to(device)
    if isinstance(v, list):
        return [_to(x, device) for x in v]
    return v


def tied_rows(probs_valid):
    """Rows whose cost row holds two exactly equal class costs (e.g. two race probabilities that round to the same 16-bit
    value).  Swapping such a row between the tied classes leaves the transport cost unchanged, so the optimal plan is not
    unique there -- POT's own choice is arbitrary -- and an exact solver may legitimately differ from
    another exact solver on THESE rows only.  ``probs_valid``: the probability tensors of the rows with a face."""
    from oracle import assign as oassign
    pa = probs_valid[2] if len(probs_valid) == 3 else None
    M = oassign.cost_matrix(probs_valid[0], probs_valid[1], pa)
    srt = np.sort(M, axis=1)
    return torch.tensor((np.diff(srt, axis=1) == 0).any(axis=1)), M


def assert_equal_up_to_ties(got, ref, probs_all, what):
    """``got`` / ``ref``: per-row tensors over ALL rows.  Bit-exact on every row whose costs are tie-free; rows with tied
    costs may differ (see tied_rows)."""
    if torch.equal(got, ref):
        return 0
    valid = (probs_all[0] != -1).all(-1) & (probs_all[1] != -1).all(-1)
    tied_v, _ = tied_rows([p[valid] for p in probs_all])
    tied = torch.zeros(valid.shape[0], dtype=torch.bool)
    tied[valid] = tied_v
    diff = got != ref
    assert not bool((diff & ~tied).any()), (what, "differs on", int((diff & ~tied).sum()), "rows with tie-free costs")
    return int(diff.sum())


def check_small_steps(device="cuda:0"):
    from fairguide import pipeline
    from oracle import pipeline as opipe
    torch.cuda.set_device(device)
    for kind in ("gender", "gender_race", "gender_race_age"):
        cfg = pipeline.GuidanceConfig(kind=kind, num_samples_per_device=20)
        host = pipeline.synth_batch(12, cfg, torch.float32, "cpu", seed=7, H=256, W=256, max_faces=2, host=True)
        head = pipeline.make_head_weights(cfg, torch.float32, "cpu")
        dev_batch = {k: _to(v, device) for k, v in host.items()}
        nv = int((host["counts"] > 0).sum())
        rands = None
        if kind != "gender":
            g = torch.Generator().manual_seed(3)
            rands = tuple(torch.rand(cfg.num_samples_per_device, nv, generator=g) for _ in range(cfg.n_attr))
        out = pipeline.GuidancePath(cfg, tuple(t.to(device) for t in head)).step(
            dev_batch, rand_tensors=None if rands is None else tuple(r.to(device) for r in rands), num_valid=nv)
        ref = opipe.step(host, cfg, head, rand_tensors=rands)
        torch.cuda.synchronize()
        assert torch.equal(out["boxes"].cpu(), ref["boxes"]), kind
        for a in range(cfg.n_attr):
            assert torch.equal(out["targets"][a].cpu(), ref["targets"][a]), (kind, a)
        np.testing.assert_allclose(out["chips"].cpu().numpy(), ref["chips"].numpy(), rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(out["small"].cpu().numpy(), ref["small"].numpy(), rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(out["loss"].cpu().numpy(), ref["loss"].numpy(), rtol=1e-3, atol=1e-4)
        gref = ref["g_images"].numpy()
        np.testing.assert_allclose(out["g_images"].cpu().numpy(), gref, rtol=1e-3, atol=1e-3 * float(np.abs(gref).max()))
    return True


def _ulps16(a, b):
    return (a.contiguous().view(torch.int16).long() - b.contiguous().view(torch.int16).long()).abs()


def check_step_vs_oracle(kind="gender_race_age", n=1024, dtype=torch.bfloat16, n_grad=32, S=100, seed=5991, device="cuda:0",
                         max_faces=1, captured=False):
    """One ``GuidancePath.step`` (or its CUDA-graph replay) on ``n`` synthetic 512x512 images against the oracle; see the
    module docstring for what is compared with what.  Returns a dict of the measured errors."""
    from fairguide import pipeline
    from oracle import assign as oassign, boxes as oboxes, crop as ocrop, head as ohead, hooks as ohooks, loss as oloss
    dev = torch.device(device)
    torch.cuda.set_device(dev)
    rtol = RTOL[dtype]
    cfg = pipeline.GuidanceConfig(kind=kind, num_samples_per_device=S)
    widths, col_start, k_head, K, e1_rule = pipeline.KINDS[kind]
    n_attr = len(widths)
    head = pipeline.make_head_weights(cfg, dtype, dev)
    batch = pipeline.synth_batch_device(n, cfg, dtype, dev, seed=seed, max_faces=max_faces)
    nv = int((batch["counts"] > 0).sum())
    rands = None
    if kind != "gender":
        g = torch.Generator().manual_seed(seed + 1)
        rands = tuple(torch.rand(S, nv, generator=g).to(dtype) for _ in range(n_attr))
    rands_dev = None if rands is None else tuple(r.to(dev) for r in rands)
    path = pipeline.GuidancePath(cfg, head)
    out = path.step(batch, rand_tensors=rands_dev, num_valid=nv)
    torch.cuda.synchronize()
    cpu = lambda t: t.detach().cpu()
    report = {}

    # ---- boxes: bit-exact
    ind_ref, boxes_ref = oboxes.select_and_expand(cpu(batch["cand_boxes"]).numpy(), cpu(batch["counts"]).numpy(), 512, cfg.expand_coef, 1, -1)
    assert np.array_equal(cpu(out["indicators"]).numpy(), ind_ref) and np.array_equal(cpu(out["boxes"]).numpy(), boxes_ref)
    ind = torch.tensor(ind_ref)
    boxes = torch.tensor(boxes_ref)

    # ---- head: logits vs the fp32 torch reference of the same (dtype-rounded) weights
    w1, b1, w2, b2 = [cpu(t).float() for t in head]
    ref_logits = ohead.mobilenet_head_reference(cpu(batch["pooled"]).float(), w1, b1, w2, b2)
    err = (cpu(out["logits"]).float() - ref_logits).abs().max().item()
    report["logits_max_err"] = err
    assert err <= rtol * float(ref_logits.abs().max()), ("head logits", err)

    # ---- probabilities / predictions: the oracle's 16-bit (or fp32) softmax of the DEVICE logits
    dev_logits = cpu(out["logits"]).to(dtype)
    fn = {"gender": ohead.get_face_gender, "gender_race": ohead.get_face_gender_race, "gender_race_age": ohead.get_face_gender_race_age}[kind]
    houts = fn(lambda x: dev_logits[ind], torch.zeros(n, 1, dtype=dtype), selector=ind, fill_value=-1)
    probs_dev = [cpu(p) for p in out["probs"]]
    for a in range(n_attr):
        assert torch.equal(cpu(out["preds"][a]), houts[3 * a]), ("preds", a)
        if dtype == torch.float32:
            assert torch.allclose(probs_dev[a], houts[3 * a + 1], rtol=1e-5, atol=1e-7), ("probs", a)
        else:
            assert int(_ulps16(probs_dev[a], houts[3 * a + 1]).max()) <= 1, ("probs", a)

    # ---- assignment: bit-exact on the device probabilities and the same draws
    if kind == "gender":
        t_all, u_all = oassign.generate_dynamic_targets(probs_dev[0], cfg.target_ratio, True)
        res = (t_all, u_all)
    elif kind == "gender_race":
        res = oassign.generate_dynamic_targets_gender_race(probs_dev[0], probs_dev[1], True, S, rand_tensors=rands, literal=False)
    else:
        res = oassign.generate_dynamic_targets_gender_race_age(probs_dev[0], probs_dev[1], probs_dev[2], True, S, rand_tensors=rands,
                                                               literal=False)
    targets = []
    for a in range(n_attr):
        t = res[2 * a].clone()
        t[res[2 * a + 1] > cfg.uncertainty_threshold] = -1                                   # E3:2022-2023
        if kind == "gender":
            assert torch.equal(cpu(out["targets"][a]), t), ("targets", a, int((cpu(out["targets"][a]) != t).sum()))
        else:
            report[f"rows_differing_by_cost_ties_{a}"] = assert_equal_up_to_ties(cpu(out["targets"][a]), t, probs_dev, f"targets[{a}]")
        targets.append(cpu(out["targets"][a]))          # the losses / gradients below are checked against the device's own targets
    report["rows_with_target"] = [int((t != -1).sum()) for t in targets]

    # ---- losses
    logits_attr = [cpu(out["logits"])[:, c:c + w] for c, w in zip(col_start, widths)]
    logits_attr = [l.to(dtype).float() for l in logits_attr]
    loss_fair = [oloss.fairness_ce(logits_attr[a], targets[a], ind, torch.float32) for a in range(n_attr)]
    preds_ori = [cpu(p) for p in batch["preds_ori"]]
    f2, f1 = list(cfg.factors2[:n_attr]), list(cfg.factors1[:n_attr])
    dyn_w = ohooks.gen_dynamic_weights(ind, targets, preds_ori, f1, torch.float32, e1_rule)
    assert torch.equal(cpu(out["dyn_weights"]), dyn_w)
    loss = oloss.assemble(loss_fair, dyn_w, cpu(batch["loss_clip"]).float(), cpu(batch["loss_dino"]).float(), cpu(batch["loss_face"]).float(),
                          cfg.weight_loss_img, cfg.weight_loss_face)
    np.testing.assert_allclose(cpu(out["loss"]).float().numpy(), loss.numpy(), rtol=rtol, atol=rtol)
    for a in range(n_attr):
        np.testing.assert_allclose(cpu(out["loss_fair"][a]).float().numpy(), loss_fair[a].numpy(), rtol=rtol, atol=rtol)

    # ---- crops, resized images, image gradient on a subset (first n_grad images + every no-face image among the first 4*n_grad)
    idx = list(range(min(n_grad, n))) + [i for i in range(min(n_grad, n), min(4 * n_grad, n)) if not bool(ind[i])][:2]
    idx_t = torch.tensor(idx)
    x = cpu(batch["images"][idx_t.to(dev)]).float().requires_grad_(True)
    sub = lambda t: t[idx_t]
    bbox_ori = cpu(out["bbox_ori"])
    chips_ref = ocrop.crop_faces(x, sub(boxes), sub(ind), cfg.size_face, cfg.fill_value)
    hooked = ohooks.apply_grad_hook_face(x, sub(boxes), sub(bbox_ori), [sub(t) for t in targets], [sub(p) for p in preds_ori], f2, e1_rule)
    small_ref = ocrop.resize_small(hooked, cfg.img_size_small)
    gc, gs = cpu(batch["g_chips"][idx_t.to(dev)]).float(), cpu(batch["g_small"][idx_t.to(dev)]).float()
    ((chips_ref * gc).sum() + (small_ref * gs).sum()).backward()
    atol16 = rtol if dtype != torch.float32 else 1e-5
    np.testing.assert_allclose(cpu(out["chips"][idx_t.to(dev)]).float().numpy(), chips_ref.detach().numpy(), rtol=rtol, atol=atol16)
    np.testing.assert_allclose(cpu(out["small"][idx_t.to(dev)]).float().numpy(), small_ref.detach().numpy(), rtol=rtol, atol=atol16)
    gref = x.grad.numpy()
    got = cpu(out["g_images"][idx_t.to(dev)]).float().numpy()
    report["g_images_max_err_rel"] = float(np.abs(got - gref).max() / np.abs(gref).max())
    np.testing.assert_allclose(got, gref, rtol=rtol, atol=rtol * float(np.abs(gref).max()))

    # ---- the CUDA-graph replay of the same step (what bench.py times) returns the same integers
    if captured:
        cap = pipeline.CapturedStep(path, batch, nv)
        o2 = cap.replay()
        torch.cuda.synchronize()
        assert torch.equal(o2["boxes"], out["boxes"]) and torch.equal(o2["chips"], out["chips"]) and torch.equal(o2["small"], out["small"])
        if kind == "gender":
            assert torch.equal(o2["g_images"], out["g_images"])
    return report


def check_multi_rank(kind, dtype, dev, rank, world, n_side_global=256, S=100, captured=True, seed=4242):
    """Correctness of the multi-rank step on hardware (NCCL): every rank runs the step on its shard of a side batch of
    ``n_side_global`` images with draws from a per-rank seeded CPU generator; then
      * ``targets_all`` and the all-reduced plan ``counts`` are BIT-IDENTICAL on all ranks (gathered and compared),
      * the gathered rows are the ranks' own rows in rank order (a6: customized_all_gather E1:222-235),
      * rank 0 recomputes the assignment with the oracle, simulating every rank's draws (``world_rand``, E3:1491-1535),
        and the targets agree bit for bit, before and after the local slice (E3:2024-2025),
      * the three-graph CapturedStep replay (NCCL between the graph replays) returns the same integers as the eager step.
    Raises on any mismatch; returns a short report dict."""
    import torch.distributed as dist
    from fairguide import pipeline
    from oracle import assign as oassign
    cfg = pipeline.GuidanceConfig(kind=kind, num_samples_per_device=S)
    widths = pipeline.KINDS[kind][0]
    n_attr = len(widths)
    n = n_side_global // world
    head = pipeline.make_head_weights(cfg, dtype, dev)
    batch = pipeline.synth_batch_device(n, cfg, dtype, dev, seed=seed + rank)
    counts_local = batch["counts"].clone()
    gathered = [torch.empty_like(counts_local) for _ in range(world)]
    dist.all_gather(gathered, counts_local)
    nv = int((torch.cat(gathered) > 0).sum().item())
    draws = lambda r: tuple(torch.rand(S, nv, generator=torch.Generator().manual_seed(seed + 100 + 7 * r + a)).to(dtype) for a in range(n_attr))
    rands = None if kind == "gender" else tuple(t.to(dev) for t in draws(rank))
    path = pipeline.GuidancePath(cfg, head)
    out = path.step(batch, rand_tensors=rands, num_valid=nv)
    pipeline.validate_step(out)
    torch.cuda.synchronize()

    def same_on_all_ranks(t, what):
        t = t.contiguous()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        for r, p in enumerate(parts):
            assert torch.equal(p, t), f"{what} differs between rank {rank} and rank {r}"

    for a in range(n_attr):
        same_on_all_ranks(out["targets_all"][a], f"targets_all[{a}]")
    if out["counts"] is not None:
        same_on_all_ranks(out["counts"], "plan counts")
    # gathered rows = every rank's own rows, in rank order
    for a in range(n_attr):
        parts = [torch.empty_like(out["probs"][a]) for _ in range(world)]
        dist.all_gather(parts, out["probs"][a].contiguous())
        assert torch.equal(torch.cat(parts), out["probs_all"][a]), "gathered probabilities are not the ranks' rows in rank order"
    parts = [torch.empty_like(out["indicators"]) for _ in range(world)]
    dist.all_gather(parts, out["indicators"].contiguous())
    assert torch.equal(torch.cat(parts), out["indicators_all"])
    # local slice (E3:2024-2025)
    for a in range(n_attr):
        assert torch.equal(out["targets"][a], out["targets_all"][a][n * rank:n * (rank + 1)])
    report = {"rows": n * world, "faces": nv}
    if rank == 0:
        probs_all = [p.detach().cpu() for p in out["probs_all"]]
        if kind == "gender":
            res = oassign.generate_dynamic_targets(probs_all[0], cfg.target_ratio, True)
        else:
            wr = [draws(r) for r in range(world)]
            fn = oassign.generate_dynamic_targets_gender_race if n_attr == 2 else oassign.generate_dynamic_targets_gender_race_age
            res = fn(*probs_all, True, S, world_rand=wr, literal=False)
        for a in range(n_attr):
            t = res[2 * a].clone()
            t[res[2 * a + 1] > cfg.uncertainty_threshold] = -1
            got = out["targets_all"][a].cpu()
            if kind == "gender":
                assert torch.equal(got, t), f"targets_all[{a}] differs from the oracle in {int((got != t).sum())} rows"
            else:
                assert_equal_up_to_ties(got, t, probs_all, f"targets_all[{a}]")
        report["rows_with_target"] = [int((out["targets_all"][a] != -1).sum()) for a in range(n_attr)]
    if captured:
        cap = pipeline.CapturedStep(path, batch, nv, rand_tensors=rands)
        o2 = cap.replay(validate=True)
        torch.cuda.synchronize()
        for a in range(n_attr):
            assert torch.equal(o2["targets_all"][a], out["targets_all"][a]), "graph replay differs from the eager step"
        if o2["counts"] is not None:
            assert torch.equal(o2["counts"], out["counts"])
        assert torch.equal(o2["g_images"], out["g_images"]) and torch.equal(o2["loss"], out["loss"])
        report["graphs_per_step"] = len(cap.graphs)
    # which transport carried the two exchanges; when it was the one-shot NVLink one (dist.PeerExchange), the same step over
    # NCCL must return the same integers
    report["transport"] = "peer" if path.peer else "nccl"
    if path.peer:
        path.peer.check()
        path_nccl = pipeline.GuidancePath(cfg, head, peer_exchange=False)
        o3 = path_nccl.step(batch, rand_tensors=rands, num_valid=nv)
        torch.cuda.synchronize()
        for a in range(n_attr):
            assert torch.equal(o3["targets_all"][a], out["targets_all"][a]), "PeerExchange and NCCL disagree on targets_all"
            assert torch.equal(o3["probs_all"][a], out["probs_all"][a]), "PeerExchange and NCCL disagree on the gathered rows"
        if o3["counts"] is not None:
            assert torch.equal(o3["counts"], out["counts"]), "PeerExchange and NCCL disagree on the plan counts"
    dist.barrier()
    return report
