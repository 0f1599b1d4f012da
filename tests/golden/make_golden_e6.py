#!/usr/bin/env python
"""Golden vectors for the rest of exp-6 (race only), produced by EXECUTING THE REFERENCE'S OWN FUNCTION BODIES
(exp-6-debias-race/1-main-debias.py), lifted unmodified with ``ast`` as in make_golden.py:

    get_face_race           E6:1365-1411   fp32 and fp16 logits, with / without selector, empty selection
    get_evaluate_metrics    E6:1624-1638   fp32 / bf16 / fp16 probabilities, rows of -1, values straddling 0.8
    apply_grad_hook_face    E6:1640-1673   with ``factor`` passed POSITIONALLY
    gen_dynamic_weights     E6:1675-1689   with ``factor`` passed positionally

    python tests/golden/make_golden_e6.py [--ref /root/reference]        (build container only) -> e6.npz
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

E6 = "exp-6-debias-race/1-main-debias.py"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    path = os.path.join(a.ref, E6)
    out = {}
    rng = np.random.Generator(np.random.PCG64(61))
    # ---- get_face_race
    n = 21
    selector = rng.uniform(size=n) > 0.3
    m = int(selector.sum())
    chips = torch.zeros(n, 1)
    for dn, dt in (("f32", torch.float32), ("f16", torch.float16)):
        logits = torch.tensor((rng.normal(size=(m, 6)) * 2.5).astype(np.float32)).to(dt)
        full = torch.tensor((rng.normal(size=(n, 6)) * 2.5).astype(np.float32)).to(dt)
        state = {"next": logits}
        ns = mg.make_namespace(mg._Recorder(), {"race_classifier": lambda x, s=state: s["next"].clone()})
        mg.compile_into(ns, mg.lift(path, ["get_face_race"]).values())
        sel_out = ns["get_face_race"](chips.to(dt), selector=torch.tensor(selector), fill_value=-1)
        state["next"] = full
        nosel_out = ns["get_face_race"](chips.to(dt), selector=None, fill_value=-1)
        empty_out = ns["get_face_race"](chips.to(dt), selector=torch.zeros(n, dtype=torch.bool), fill_value=-1)
        out[f"race_head_{dn}_logits"], out[f"race_head_{dn}_logits_full"] = mg._bits(logits), mg._bits(full)
        for k in range(3):
            out[f"race_head_{dn}_sel{k}"] = mg._bits(sel_out[k])
            out[f"race_head_{dn}_nosel{k}"] = mg._bits(nosel_out[k])
            out[f"race_head_{dn}_empty{k}"] = mg._bits(empty_out[k])
    out["race_head_selector"] = selector
    # ---- get_evaluate_metrics(probs_race_all)
    ns = {"torch": torch}
    mg.compile_into(ns, mg.lift(path, ["get_evaluate_metrics"]).values())
    for di, (dn, dt) in enumerate((("f32", torch.float32), ("bf16", torch.bfloat16), ("f16", torch.float16))):
        gen = torch.Generator().manual_seed(600 + di)
        pr = torch.softmax(torch.randn(301, 4, generator=gen) * 1.5, -1).to(dt)
        pr[5:9] = torch.tensor([0.796875, 0.1, 0.05, 0.053125]).to(dt)      # straddle the 0.8 threshold in every dtype
        pr[9] = torch.tensor([0.7998046875, 0.2001953125, 0.0, 0.0]).to(dt)
        pr[10] = 0.25                                                            # argmax tie
        pr[torch.rand(301, generator=gen) < 0.1] = -1
        out[f"race_metrics_{dn}_pr"] = mg._bits(pr)
        out[f"race_metrics_{dn}_out"] = np.array(ns["get_evaluate_metrics"](pr), dtype=np.float64)
    # ---- hook and weights with the factor passed positionally
    g = np.load(os.path.join(HERE, "hooks.npz"))
    ns = mg.make_namespace(mg._Recorder())
    mg.compile_into(ns, mg.lift(path, ["make_grad_hook", "apply_grad_hook_face", "gen_dynamic_weights"]).values())
    b = g["images"].shape[0]
    targets = rng.integers(-1, 4, size=b); preds = rng.integers(0, 4, size=b)
    preds[4] = -1; targets[0] = preds[0]
    probs = np.full((b, 4), 0.25, dtype=np.float32)
    x = torch.tensor(g["images"], requires_grad=True)
    y = ns["apply_grad_hook_face"](x, torch.tensor(g["box"]), torch.tensor(g["box_ori"]), torch.tensor(targets), torch.tensor(preds),
                                   torch.tensor(probs), 0.15)
    (y * torch.tensor(g["upstream"])).sum().backward()
    face_ind = ~(g["box"] == -1).all(axis=1)
    w = ns["gen_dynamic_weights"](torch.tensor(face_ind), torch.tensor(targets), torch.tensor(preds), torch.tensor(probs), 0.35)
    out["race_hook_targets"], out["race_hook_preds"] = targets.astype(np.int64), preds.astype(np.int64)
    out["race_hook_grad"], out["race_hook_weights"] = x.grad.numpy(), w.numpy()
    np.savez_compressed(os.path.join(HERE, "e6.npz"), **out)
    print("wrote e6.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
