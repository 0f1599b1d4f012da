"""The guidance path as ONE forward+backward pass over a batch of images (what bench.py times).

Stage order (reference lines in parentheses; a1..a15 = SURVEY.md section 8a rows):

  1  select largest face + expand box            a1,a2   (E1:1292-1304, E1:238-265)
  2  fused crop 224^2 + resize 224^2             a3,a4,a12 (E1:267-290, E1:1905)
  3  [classifier backbone: torchvision/cuDNN, not part of the path -- either the real
      MobileNetV3 `features` (mode "backbone") or stand-in tensors (mode "standin")]
  4  dense head + per-attribute softmax/argmax   a5      (E3:1387-1457)
  5  all-gather of {indicator, probs}            a6      (E3:1978-1986)
  6  balanced assignment + threshold + slice     a7-a10  (E1:1403-1447 / E3:1459-1569 / E4:1477-1615, E3:2022-2025)
  7  fairness CE, its gradient, head backward    a14     (E3:2114-2122)
  8  hook region/scale + dynamic weights         a11,a13 (E3:1751-1803)
  9  loss assembly                               a14     (E3:2146-2147)
  10 fused image gradient                        a11,a12,a3 backward

The reference runs stages 1-6 without gradient on one generation of the images and stages 2-4,7-10
with gradient on a second, numerically identical generation (E3:1956-2147); the owned arithmetic is
the same, so one pass with shared forward results is what is measured.
"""
import dataclasses
from typing import Optional

import torch

from . import api, dist as fdist, ops

KINDS = {
    # kind: (attribute widths, col_start, k_head, K classes of the assignment, e1_rule)
    "gender": ([2], [40], 80, 2, True),
    "gender_race": ([2, 4], [0, 2], 6, 8, False),
    "gender_race_age": ([2, 4, 2], [0, 2, 6], 8, 16, False),
}


@dataclasses.dataclass
class GuidanceConfig:
    kind: str = "gender_race"
    num_samples_per_device: int = 100            # E3:1460 keyword default
    uncertainty_threshold: float = 0.2           # every shipped YAML (e.g. exp-3 configs/debias-text-encoder.yaml:10)
    target_ratio: float = 0.5                    # E1:1404
    factors1: tuple = (0.2, 0.6, 0.6)            # dynamic weights: factor1_gender/race/age (E4:484-487 defaults)
    factors2: tuple = (0.2, 0.3, 0.3)            # hook: factor2_*
    weight_loss_img: float = 8.0
    weight_loss_face: float = 1.0
    size_face: int = 224
    img_size_small: int = 224
    fill_value: float = -1.0
    expand_coef: float = 0.5
    d_in: int = 960
    d_hid: int = 1280

    @property
    def n_attr(self):
        return len(KINDS[self.kind][0])


def make_head_weights(cfg: GuidanceConfig, dtype, device, seed=0, logit_scale=6.0):
    """Random-init head of the MobileNetV3 layout (no checkpoint is available offline); the last
    layer is scaled up so that a good share of faces pass the uncertainty threshold."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    k_head = KINDS[cfg.kind][2]
    w1 = torch.randn(cfg.d_hid, cfg.d_in, generator=g) / cfg.d_in ** 0.5
    b1 = torch.randn(cfg.d_hid, generator=g) * 0.1
    w2 = torch.randn(k_head, cfg.d_hid, generator=g) * (logit_scale / cfg.d_hid ** 0.5)
    b2 = torch.randn(k_head, generator=g) * 0.1
    return tuple(t.to(device=device, dtype=dtype) for t in (w1, b1, w2, b2))


def synth_batch(n, cfg: GuidanceConfig, dtype, device, seed=5991, H=512, W=512, max_faces=1, no_face_frac=0.05,
                host=False):
    """Synthetic images / candidate boxes / stand-ins of SURVEY.md section 8d, generated on the CPU
    generator (so the oracle sees the same values) and moved to ``device``."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    yy = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    xx = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    ph = torch.rand(n, 3, 1, 1, generator=g) * 6.28
    fr = 0.01 + 0.03 * torch.rand(n, 3, 1, 1, generator=g)
    images = torch.empty(n, 3, H, W)
    for s in range(0, n, 64):            # chunked: same values as one big expression (the noise is drawn in memory order), ~10x less peak memory
        e = min(s + 64, n)
        images[s:e] = 0.6 * torch.sin(fr[s:e] * xx + ph[s:e]) * torch.cos(fr[s:e] * 0.7 * yy - ph[s:e]) \
            + (torch.rand(e - s, 3, H, W, generator=g) - 0.5) * 0.2
    images = images.clamp_(-1, 1)
    ctr = 160 + 192 * torch.rand(n, max_faces, 2, generator=g)
    side = 96 + 192 * torch.rand(n, max_faces, 2, generator=g) * torch.tensor([1.0, 1.0])
    side[..., 1] = side[..., 0] * (0.8 + 0.4 * torch.rand(n, max_faces, generator=g))
    cand = torch.cat([ctr - side / 2, ctr + side / 2], dim=-1).to(torch.float32)
    if max_faces > 1:
        counts = torch.randint(1, max_faces + 1, (n,), generator=g, dtype=torch.int32)
        counts = torch.where(torch.rand(n, generator=g) < no_face_frac, torch.zeros_like(counts), counts)
    else:
        counts = (torch.rand(n, generator=g) >= no_face_frac).to(torch.int32)
    jitter = torch.randint(-12, 13, (n, 4), generator=g)
    k_head = KINDS[cfg.kind][2]
    batch = dict(
        images=images.to(dtype), cand_boxes=cand, counts=counts, bbox_jitter=jitter,
        pooled=torch.randn(n, cfg.d_in, generator=g).to(dtype),                                   # backbone stand-in (fwd)
        g_chips=(torch.randn(n, 3, cfg.size_face, cfg.size_face, generator=g) * 1e-3).to(dtype),  # backbone stand-in (bwd)
        g_small=(torch.randn(n, 3, cfg.img_size_small, cfg.img_size_small, generator=g) * 1e-3).to(dtype),  # CLIP/DINO grad stand-in
        loss_clip=torch.rand(n, generator=g).to(dtype), loss_dino=torch.rand(n, generator=g).to(dtype),
        loss_face=torch.rand(n, generator=g).to(dtype),
        preds_ori=[torch.randint(0, w, (n,), generator=g) for w in KINDS[cfg.kind][0]],
        k_head=k_head,
    )
    # probs_*_ori (E3:1751): only their dtype is read by the hook / weight functions; 0.9 on the original prediction
    batch["probs_ori"] = [(torch.nn.functional.one_hot(p, w) * (0.9 - 0.1 / max(w - 1, 1)) + 0.1 / max(w - 1, 1)).to(dtype)
                          for p, w in zip(batch["preds_ori"], KINDS[cfg.kind][0])]
    if not host:
        batch = {k: _to(v, device) for k, v in batch.items()}
    return batch


def synth_batch_device(n, cfg: GuidanceConfig, dtype, device, seed=5991, H=512, W=512, max_faces=1, no_face_frac=0.05):
    """Same distributions as synth_batch, generated on the device in chunks (BASELINE-size batches
    do not fit a CPU-side fp32 staging copy comfortably).  Used by bench.py only."""
    g = torch.Generator(device=device).manual_seed(seed)
    images = torch.empty((n, 3, H, W), dtype=dtype, device=device)
    yy = torch.arange(H, dtype=torch.float32, device=device).view(1, 1, H, 1)
    xx = torch.arange(W, dtype=torch.float32, device=device).view(1, 1, 1, W)
    for s in range(0, n, 32):
        m = min(32, n - s)
        ph = torch.rand(m, 3, 1, 1, generator=g, device=device) * 6.28
        fr = 0.01 + 0.03 * torch.rand(m, 3, 1, 1, generator=g, device=device)
        im = 0.6 * torch.sin(fr * xx + ph) * torch.cos(fr * 0.7 * yy - ph) + (torch.rand(m, 3, H, W, generator=g, device=device) - 0.5) * 0.2
        images[s:s + m] = im.clamp_(-1, 1).to(dtype)
    r = lambda *shape: torch.rand(*shape, generator=g, device=device)
    ctr = 160 + 192 * r(n, max_faces, 2)
    side = 96 + 192 * r(n, max_faces, 2)
    side[..., 1] = side[..., 0] * (0.8 + 0.4 * r(n, max_faces))
    cand = torch.cat([ctr - side / 2, ctr + side / 2], dim=-1).to(torch.float32)
    if max_faces > 1:
        counts = torch.randint(1, max_faces + 1, (n,), generator=g, device=device, dtype=torch.int32)
    else:
        counts = torch.ones(n, dtype=torch.int32, device=device)
    counts = torch.where(r(n) < no_face_frac, torch.zeros_like(counts), counts)
    rn = lambda *shape: torch.randn(*shape, generator=g, device=device)
    jitter = torch.randint(-12, 13, (n, 4), generator=g, device=device)
    # face_bboxs_ori (E3:1751): the expanded box found on the ORIGINAL model's image, an input of the path;
    # here the new box plus a few pixels of jitter
    ind0, boxes0 = ops.select_expand_boxes(cand, counts, images.shape[-2], cfg.expand_coef, 1.0, -1)
    bbox_ori = torch.where(ind0.unsqueeze(1), boxes0 + jitter, boxes0)
    batch = dict(
        images=images, cand_boxes=cand, counts=counts, bbox_jitter=jitter, bbox_ori=bbox_ori,
        pooled=rn(n, cfg.d_in).to(dtype),
        g_chips=(rn(n, 3, cfg.size_face, cfg.size_face) * 1e-3).to(dtype),
        g_small=(rn(n, 3, cfg.img_size_small, cfg.img_size_small) * 1e-3).to(dtype),
        loss_clip=r(n).to(dtype), loss_dino=r(n).to(dtype), loss_face=r(n).to(dtype),
        preds_ori=[torch.randint(0, w, (n,), generator=g, device=device) for w in KINDS[cfg.kind][0]],
        k_head=KINDS[cfg.kind][2],
    )
    batch["probs_ori"] = [(torch.nn.functional.one_hot(p, w) * (0.9 - 0.1 / max(w - 1, 1)) + 0.1 / max(w - 1, 1)).to(dtype)
                          for p, w in zip(batch["preds_ori"], KINDS[cfg.kind][0])]
    return batch


def _to(v, device):
    if torch.is_tensor(v):
        return v.to(device)
    if isinstance(v, list):
        return [_to(x, device) for x in v]
    return v


class _NullProbe:
    @staticmethod
    def begin(name):
        pass

    @staticmethod
    def end(name):
        pass


class GuidancePath:
    """Holds the frozen head and the configuration; ``step`` runs stages 1-10 on this rank's images."""

    def __init__(self, cfg: GuidanceConfig, head_weights, backbone=None, group=None, peer_exchange="auto"):
        """``peer_exchange``: "auto" = the one-shot NVLink transport (dist.PeerExchange) when several NCCL ranks can map each
        other's memory, else one NCCL call per exchange; False = NCCL / gloo always."""
        self.cfg, self.head, self.backbone, self.group = cfg, head_weights, backbone, group
        self.peer_mode, self.peer = peer_exchange, None

    def _peer_for(self, slot, n_all, K, device):
        """The PeerExchange for this batch shape: created collectively on the first multi-rank step; None = use NCCL (transport
        switched off or unavailable, rows not a multiple of 16 bytes, or a batch larger than the one it was sized for)."""
        if self.peer_mode in (False, "off") or self.peer is False:
            return None
        counts_bytes = ((n_all * K * 4 + 15) // 16) * 16
        if self.peer is None:
            if slot % 16 or (n_all // fdist._world(self.group)[0]) % 8 or not fdist.PeerExchange.available(device):
                return None
            try:
                self.peer = fdist.PeerExchange(slot, counts_bytes, device, self.group)
            except Exception as e:      # no peer mapping on this machine: say so once, stay on NCCL
                import sys
                print(f"fairguide: PeerExchange unavailable ({type(e).__name__}: {e}); the exchanges use NCCL", file=sys.stderr)
                self.peer = False
                return None
        if self.peer.slot_bytes != slot or self.peer.counts_bytes < counts_bytes:
            return None
        return self.peer

    # The step is written as three local phases around the two exchanges of the path, so that the
    # phases can be recorded into CUDA graphs while the collectives stay ordinary NCCL calls between them.
    @torch.no_grad()
    def phase_a(self, batch, probe=None):
        """Stages 1-4 on this rank's images, plus the packed {indicator, probs} row block the all-gather sends."""
        cfg = self.cfg
        P = probe if probe is not None else _NullProbe
        widths, col_start, k_head, K, e1_rule = KINDS[cfg.kind]
        images = batch["images"]
        n, _, H, W = images.shape
        ind, boxes = ops.select_expand_boxes(batch["cand_boxes"], batch["counts"], H, cfg.expand_coef, 1.0, -1)
        P.begin("sample_fwd")
        chips, small = ops.crop_resize_fwd(images, boxes, ind, (cfg.size_face,) * 2, (cfg.img_size_small,) * 2, cfg.fill_value)
        P.end("sample_fwd")
        pooled = batch["pooled"] if self.backbone is None else self.backbone(chips)
        logits, hidden = ops.head_fwd(pooled, *self.head)
        preds, probs, logits_attr = ops.head_attributes(logits, None, ind, n, col_start, widths, cfg.fill_value, images.dtype)
        st = dict(ind=ind, boxes=boxes, chips=chips, small=small, logits=logits, hidden=hidden, preds=preds, probs=probs,
                  logits_attr=logits_attr)
        world, _ = fdist._world(self.group)
        if world > 1:
            st["multi"] = True
            wtot = 1 + sum(widths)
            st["peer"] = self._peer_for(n * wtot * probs[0].element_size(), world * n, K, images.device) if images.is_cuda else None
            if st["peer"] is not None:
                # one-shot NVLink transport: the rows leave for every peer as soon as they exist, packed on the way
                st["peer"].begin_step()
                st["peer"].push_rows_from(ind, probs)
            else:
                st["packed"] = fdist.pack_probs(ind, probs)
                st["gathered"] = torch.empty((world * n, wtot), dtype=st["packed"].dtype, device=images.device)
        return st

    def exchange_1(self, st):
        """Stage 5: the all-gather (no-op with one rank)."""
        if st.get("peer") is not None:
            widths = KINDS[self.cfg.kind][0]
            world, _ = fdist._world(self.group)
            st["unpacked"] = st["peer"].wait_rows_into(world * st["ind"].shape[0], widths, st["probs"][0].dtype, st["ind"].device)
        elif "packed" in st:
            fdist.all_gather_packed(st["gathered"], st["packed"], self.group)

    @torch.no_grad()
    def phase_b(self, st, batch, rand_tensors=None, num_valid=None):
        """Stage 6 up to the second exchange: this rank's Monte-Carlo plan counts over the gathered rows."""
        cfg = self.cfg
        widths, col_start, k_head, K, e1_rule = KINDS[cfg.kind]
        images = batch["images"]
        if "unpacked" in st:
            ind_all, probs_all = st["unpacked"]
        elif "gathered" in st:
            ind_all, probs_all = fdist.unpack_probs(st["gathered"], widths)
        else:
            ind_all, probs_all = st["ind"], st["probs"]
        st["ind_all"], st["probs_all"] = ind_all, probs_all
        st["counts"] = None
        if cfg.kind != "gender":
            n_all = ind_all.shape[0]
            nv = num_valid if num_valid is not None else int(ind_all.sum().item())
            ws = ops.OtWorkspace(n_all, K, cfg.num_samples_per_device, images.device)
            if rand_tensors is None and nv > 0:
                # one generator launch for all attributes (api.generate_dynamic_targets_* keeps the reference's
                # one-torch.rand-per-attribute order for seeded parity)
                rand_tensors = tuple(torch.rand([len(widths), cfg.num_samples_per_device, nv], dtype=images.dtype,
                                                device=images.device).unbind(0))
            pa = probs_all[2] if len(widths) == 3 else None
            st["counts"] = ops.ot_plan_counts(probs_all[0], probs_all[1], pa, rand_tensors or (), nv, ws)
            st["nv"], st["ws"] = nv, ws
        return st

    def exchange_2(self, st):
        """The all-reduce of the plan counts (E3:1535; no-op with one rank or for the E1 rule)."""
        if st.get("counts") is not None and st["nv"] > 0:
            if st.get("peer") is not None and st["counts"].numel() * 4 <= st["peer"].counts_bytes:
                st["peer"].sum_counts(st["counts"])
            else:
                fdist.all_reduce_counts(st["counts"], self.group)

    @torch.no_grad()
    def phase_c(self, st, batch, probe=None, close_assign=False):
        """Rest of stage 6 (targets, threshold, local slice) and stages 7-10."""
        cfg = self.cfg
        P = probe if probe is not None else _NullProbe
        widths, col_start, k_head, K, e1_rule = KINDS[cfg.kind]
        images = batch["images"]
        n, _, H, W = images.shape
        _, rank = fdist._world(self.group)
        ind, boxes, probs_all = st["ind"], st["boxes"], st["probs_all"]
        if cfg.kind == "gender":
            t_all, _ = ops.assign_rank_binom(probs_all[0], cfg.target_ratio, cfg.uncertainty_threshold, True)
            targets_all = [t_all]
        else:
            targets_all, _ = ops.ot_targets(st["counts"], probs_all[0], probs_all[1], st["nv"], st["ws"], cfg.uncertainty_threshold, True)
        targets = [t[n * rank:n * (rank + 1)] for t in targets_all]
        if close_assign:
            P.end("assign")
        # 7-9: hook factors, then CE of every attribute + loss assembly + gradient wrt the head logits in one launch
        bbox_ori = batch["bbox_ori"] if "bbox_ori" in batch else torch.where(ind.unsqueeze(1), boxes + batch["bbox_jitter"], boxes)
        region, scale, dyn_w = ops.guidance_factors(ind, boxes, bbox_ori, targets, batch["preds_ori"],
                                                    cfg.factors2[:len(widths)], cfg.factors1[:len(widths)], e1_rule, H, W)
        loss_fair, loss, g_logits = ops.fair_loss_fused(
            st["logits_attr"], targets, col_start, k_head, ind, 1.0 / n, dyn_w, batch["loss_clip"], batch["loss_dino"],
            batch["loss_face"], cfg.weight_loss_img, cfg.weight_loss_face, -1.0)
        g_pooled = ops.head_bwd(g_logits, st["hidden"], self.head[0], self.head[2])
        # 10
        P.begin("image_grad")
        g_images = ops.image_grad(batch["g_chips"], batch["g_small"], boxes, ind, region, scale, tuple(images.shape),
                                  images.dtype, images.device)
        P.end("image_grad")
        return dict(indicators=ind, boxes=boxes, chips=st["chips"], small=st["small"], logits=st["logits"], preds=st["preds"],
                    probs=st["probs"], targets=targets, targets_all=targets_all, counts=st["counts"], loss_fair=loss_fair,
                    g_pooled=g_pooled, region=region, scale=scale, dyn_weights=dyn_w, loss=loss, loss_mean=loss.mean(),
                    g_images=g_images, bbox_ori=bbox_ori,
                    ot_status=st["ws"].status_tensor() if st.get("counts") is not None else None, num_valid=st.get("nv"),
                    indicators_all=st["ind_all"], probs_all=probs_all, peer=st.get("peer"))

    @torch.no_grad()
    def step(self, batch, rand_tensors=None, num_valid: Optional[int] = None, probe=None):
        """``probe``: optional object with begin(name)/end(name) (bench.py records CUDA events on the
        current stream around the named stages)."""
        P = probe if probe is not None else _NullProbe
        st = self.phase_a(batch, probe)
        P.begin("assign")
        Q = P if "multi" in st else _NullProbe           # the parts of the assignment stage are of interest with several ranks only
        Q.begin("exchange_rows")
        self.exchange_1(st)
        Q.end("exchange_rows")
        Q.begin("plan_counts")
        self.phase_b(st, batch, rand_tensors, num_valid)
        Q.end("plan_counts")
        Q.begin("exchange_counts")
        self.exchange_2(st)
        Q.end("exchange_counts")
        return self.phase_c(st, batch, probe, close_assign=True)


def validate_step(out):
    """Raise if the assignment kernels of the step that produced ``out`` flagged a problem -- above all a ``num_valid``
    that does not match the faces actually present (one 16-byte device-to-host read; synchronises)."""
    if out.get("ot_status") is not None:
        ops.check_ot_status(out["ot_status"].tolist(), out.get("num_valid"))
    if out.get("peer"):
        out["peer"].check()


class CapturedStep:
    """``GuidancePath.step`` recorded once into a CUDA graph and replayed: the ~15 launches (and, with several ranks, the
    two collectives) of a step are enqueued by one driver call, which removes the host-side gaps between the small
    kernels.  The batch tensors are static: refresh them in place (``copy_``) before ``replay``; ``out`` holds the
    step's result tensors, overwritten by every replay.  The Monte-Carlo draws advance with every replay (torch's
    graph-safe generator), exactly like consecutive eager steps.

    THE NUMBER OF FACES IS FROZEN AT CAPTURE: ``num_valid`` fixes the shape of the draws and of the plan counts inside the
    graph.  A refreshed batch with a different number of detected faces must be re-captured; ``replay(validate=True)``
    (or ``pipeline.validate_step(out)``) reads the status words the kernels leave behind and raises on a mismatch instead of
    returning wrong targets -- the kernels themselves stay inside their buffers either way."""

    def __init__(self, path, batch, num_valid, rand_tensors=None):
        """``rand_tensors``: optional static [S, num_valid] draw buffers (one per attribute) that the graph reads instead
        of drawing with torch.rand -- refresh them in place between replays (parity tests hand the same draws to the oracle)."""
        self.path, self.batch = path, batch
        world, _ = fdist._world(path.group)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capturing stream (allocator, lazy module loads)
            for _ in range(2):
                path.step(batch, rand_tensors=rand_tensors, num_valid=num_valid)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graphs = []
        if world == 1 or path.peer:
            # one rank, or the exchanges are kernels of ours (dist.PeerExchange): the whole step is ONE graph
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.out = path.step(batch, rand_tensors=rand_tensors, num_valid=num_valid)
            self.graphs = [g]
            self.state = None
        else:
            # one graph per local phase; the two collectives run between the replays as ordinary NCCL calls on the
            # static buffers the graphs read and write
            ga, gb, gc = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                st = path.phase_a(batch)
            path.exchange_1(st)
            with torch.cuda.graph(gb, pool=ga.pool()):
                path.phase_b(st, batch, rand_tensors, num_valid)
            path.exchange_2(st)
            with torch.cuda.graph(gc, pool=ga.pool()):
                self.out = path.phase_c(st, batch)
            self.graphs, self.state = [ga, gb, gc], st

    def replay(self, validate=False):
        if self.state is None:
            self.graphs[0].replay()
        else:
            self.graphs[0].replay()
            self.path.exchange_1(self.state)
            self.graphs[1].replay()
            self.path.exchange_2(self.state)
            self.graphs[2].replay()
        if validate:
            validate_step(self.out)
        return self.out
