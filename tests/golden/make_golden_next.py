#!/usr/bin/env python
"""Golden vectors for the "next" rows, produced by EXECUTING THE REFERENCE'S OWN STATEMENTS.

    python tests/golden/make_golden_next.py [--ref /root/reference]        (build container only)

* get_evaluate_metrics (E3:1716, E4:1780): the FunctionDef nodes are lifted unmodified with ``ast`` (as in
  make_golden.py) and run on seeded probability tensors (fp32, bf16, fp16; some rows -1).
* detector staging: the assignment ``images_np = ...`` inside get_face_app (E1:1317) and the subscript passed to
  ``face_app.get`` (E1:1326) are located in the syntax tree and their VALUE expressions evaluated with ``images`` /
  ``image_np`` bound to seeded tensors.  Nothing from the reference is written into this repo.
* generate_dynamic_targets_race (E6:1413-1482): FunctionDef lifted unmodified and run on seeded fp32 probabilities
  (N = 3..26 faces, some rows -1).  ``ot`` (POT 0.9.3) is not installed: ``ot.emd`` -> exact assignment by
  scipy.optimize.linear_sum_assignment (as in make_golden.py), ``ot.dist`` -> the arithmetic of POT's
  ``euclidean_distances`` on the numpy backend (einsum norms + ``-2 X.Y^T``, clamp, sqrt).
"""
import argparse
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
E1 = "exp-1-debias-gender/1-main-debias.py"
E3 = "exp-3-debias-gender-race/1-main-debias.py"
E4 = "exp-4-debias-gender-race-age/1-main-debias.py"
E6 = "exp-6-debias-race/1-main-debias.py"


def pot_dist(x1, x2, metric="euclidean", p=1):
    assert metric == "euclidean"
    a2 = np.einsum("ij,ij->i", x1, x1)
    b2 = np.einsum("ij,ij->i", x2, x2)
    c = -2 * np.dot(x1, x2.T)
    c += a2[:, None]
    c += b2[None, :]
    return np.sqrt(np.maximum(c, 0))


def lsa_emd(a, b, M):
    from scipy.optimize import linear_sum_assignment
    b = np.asarray(b, dtype=np.int64)
    cols = np.repeat(np.arange(M.shape[1]), b)
    r, c = linear_sum_assignment(np.asarray(M, dtype=np.float64)[:, cols])
    T = np.zeros(M.shape, dtype=np.float64)
    T[r, cols[c]] = 1.0
    return T


def find_function(tree, name):
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return node
    raise KeyError(name)


def compile_function(node, ns):
    mod = ast.Module(body=[node], type_ignores=[])
    exec(compile(mod, "<reference>", "exec"), ns)
    return ns[node.name]


def staging_expressions(tree):
    fn = find_function(tree, "get_face_app")
    assign = next(n for n in ast.walk(fn) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "images_np")
    call = next(n for n in ast.walk(fn) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute) and n.func.attr == "get"
                and isinstance(n.func.value, ast.Name) and n.func.value.id == "face_app")
    to_code = lambda e: compile(ast.Expression(body=e), "<reference>", "eval")
    return to_code(assign.value), to_code(call.args[0])


def probs(gen, n, w, dtype, sharp):
    p = torch.softmax(torch.randn(n, w, generator=gen) * sharp, -1).to(dtype)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    out = {}
    # ---- metrics
    for tag, rel, with_age in (("e3", E3, False), ("e4", E4, True)):
        tree = ast.parse(open(os.path.join(a.ref, rel)).read())
        fn = compile_function(find_function(tree, "get_evaluate_metrics"), {"torch": torch, "abs": abs})
        for di, dtype in enumerate((torch.float32, torch.bfloat16, torch.float16)):
            gen = torch.Generator().manual_seed(100 + di)
            n = 257
            pg, pr, pa = probs(gen, n, 2, dtype, 1.5), probs(gen, n, 4, dtype, 1.5), probs(gen, n, 2, dtype, 1.5)
            pg[5:9, 0] = 0.796875; pg[5:9, 1] = 1 - 0.796875          # straddle the 0.8 confidence threshold in every dtype
            pg[9, 0] = 0.7998046875; pg[9, 1] = 0.2001953125
            pg[10] = 0.5                                               # argmax tie
            miss = torch.rand(n, generator=gen) < 0.1
            for t in (pg, pr, pa):
                t[miss] = -1
            res = fn(pg, pr, pa) if with_age else fn(pg, pr)
            key = f"metrics_{tag}_{str(dtype).split('.')[-1]}"
            out[key + "_pg"], out[key + "_pr"], out[key + "_pa"] = pg.float().numpy(), pr.float().numpy(), pa.float().numpy()
            out[key + "_out"] = np.array(res, dtype=np.float64)
    # ---- staging
    tree = ast.parse(open(os.path.join(a.ref, E1)).read())
    expr_u8, expr_bgr = staging_expressions(tree)
    for di, dtype in enumerate((torch.float32, torch.bfloat16, torch.float16)):
        gen = torch.Generator().manual_seed(200 + di)
        images = (torch.rand(3, 3, 10, 12, generator=gen) * 2.1 - 1.05).to(dtype)      # a few values outside [-1,1]
        images[0, :, 0, :4] = torch.tensor([-1.0, 1.0, 0.0, 0.999])
        images_np = eval(expr_u8, {"np": np, "torch": torch, "images": images})
        bgr = np.stack([eval(expr_bgr, {"image_np": im}) for im in images_np])
        key = f"stage_{str(dtype).split('.')[-1]}"
        out[key + "_in"] = images.float().numpy()
        out[key + "_out"] = np.ascontiguousarray(bgr)
    # ---- E6 enumerated-composition assignment
    import itertools
    import math
    import types
    tree = ast.parse(open(os.path.join(a.ref, E6)).read())
    fn = compile_function(find_function(tree, "generate_dynamic_targets_race"),
                          {"torch": torch, "np": np, "math": math, "itertools": itertools,
                           "ot": types.SimpleNamespace(dist=pot_dist, emd=lsa_emd)})
    cases = [(3, 0, 2.0), (8, 2, 1.0), (17, 3, 2.0), (26, 4, 0.7), (24, 0, 3.0)]
    out["race_n_cases"] = np.array(len(cases))
    for c, (nv, nmiss, sharp) in enumerate(cases):
        gen = torch.Generator().manual_seed(300 + c)
        p = probs(gen, nv + nmiss, 4, torch.float32, sharp)
        if nmiss:
            p[torch.randperm(nv + nmiss, generator=gen)[:nmiss]] = -1
        t, u = fn(p, True)
        out[f"race_probs_{c}"], out[f"race_targets_{c}"], out[f"race_unc_{c}"] = p.numpy(), t.numpy(), u.numpy()
    # ---- image_pipeline (E1:292-312): the reference's glue around skimage / kornia (both absent -> oracle restatements)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import align as oalign

    class _Similarity:
        def estimate(self, src, dst):
            self.params = oalign.umeyama(src, dst, True)
            return True

    tree = ast.parse(open(os.path.join(a.ref, E1)).read())
    fn = compile_function(find_function(tree, "image_pipeline"),
                          {"torch": torch, "np": np, "transform": types.SimpleNamespace(SimilarityTransform=_Similarity),
                           "kornia": types.SimpleNamespace(geometry=types.SimpleNamespace(transform=types.SimpleNamespace(
                               warp_affine=lambda src, M, dsize, mode, padding_mode, align_corners:
                               oalign.warp_affine(src, M, dsize, align_corners=align_corners))))})
    rng = np.random.default_rng(7)
    out["align_n_cases"] = np.array(3)
    for c in range(3):
        gen = torch.Generator().manual_seed(400 + c)
        yy = torch.arange(160, dtype=torch.float32).view(1, 160, 1)
        xx = torch.arange(192, dtype=torch.float32).view(1, 1, 192)
        img = (0.6 * torch.sin(0.05 * xx + c) * torch.cos(0.04 * yy) + (torch.rand(3, 160, 192, generator=gen) - 0.5) * 0.2).clamp(-1, 1)
        s, th = [(1.2, 0.3), (0.9, -0.4), (0.5, 0.1)][c]
        R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        lm = ((oalign.TEMPLATE_112 - 56.0) @ (R.T * s) + [[90, 70], [30, 120], [150, 80]][c] + rng.normal(size=(5, 2))).astype(np.float32)
        res = fn(img, lm)
        out[f"align_img_{c}"], out[f"align_lm_{c}"], out[f"align_out_{c}"] = img.numpy(), lm, res.numpy()
    # ---- adjusted-DFT gradient coefficients (E1:1104-1109): the statements that build grad_coefs inside generate_image_w_gradient
    fn_node = find_function(tree, "generate_image_w_gradient")
    stmts = []
    for node in fn_node.body:
        seg = ast.get_source_segment(open(os.path.join(a.ref, E1)).read(), node) or ""
        if "grad_coefs" in seg and "register_hook" not in seg:
            stmts.append(node)
    assert len(stmts) == 4, [ast.dump(n)[:60] for n in stmts]
    code = compile(ast.Module(body=stmts, type_ignores=[]), "<reference>", "exec")
    for c, (T, steps) in enumerate([(1000, 20), (1000, 25), (500, 7)]):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, T, dtype=torch.float32) ** 2        # Stable Diffusion's scaled-linear schedule
        alphas = 1.0 - betas
        sched = types.SimpleNamespace(alphas=alphas, alphas_cumprod=torch.cumprod(alphas, dim=0),
                                      timesteps=torch.linspace(T - 1, 0, steps).round().long())
        ns = {"noise_scheduler": sched, "np": np, "math": math}
        exec(code, ns)
        out[f"dft_alphas_{c}"], out[f"dft_alphas_cumprod_{c}"] = sched.alphas.numpy(), sched.alphas_cumprod.numpy()
        out[f"dft_timesteps_{c}"], out[f"dft_coefs_{c}"] = sched.timesteps.numpy(), np.asarray(ns["grad_coefs"], dtype=np.float64)
    out["dft_n_cases"] = np.array(3)
    # ---- face-realism loss: FaceFeatsModel (E1:82-117) and the loss_face_ij block of the training loop (E1:1917-1929,
    #      E3:2124-2143, E4:2253-2272), lifted unmodified.  sentence-transformers is absent: util.semantic_search / dot_score
    #      are stubbed (dot-product matrix + topk); get_face_feats is replaced by the identity on pre-normalised feature rows
    #      (the network is external to the path).
    import pickle
    import tempfile

    def _dot_score(a, b):
        return a @ b.T

    def _semantic_search(query, corpus, score_function=None, top_k=1):
        scores = score_function(query, corpus)
        vals, idx = scores.topk(top_k, dim=1)
        return [[{"corpus_id": int(i), "score": float(v)} for v, i in zip(vr, ir)] for vr, ir in zip(vals, idx)]

    util_stub = types.SimpleNamespace(semantic_search=_semantic_search, dot_score=_dot_score)

    def loss_block(tree_):
        for node in ast.walk(tree_):
            body = getattr(node, "body", None)
            if not isinstance(body, list):
                continue
            for k, stmt in enumerate(body):
                if (isinstance(stmt, ast.Assign) and getattr(stmt.targets[0], "id", None) == "loss_face_ij"
                        and "torch.ones" in ast.unparse(stmt.value)):
                    stmts = []
                    for nxt in body[k:]:
                        if isinstance(nxt, ast.Assign) and getattr(nxt.targets[0], "id", None) == "dynamic_weights":
                            break
                        stmts.append(nxt)
                    return compile(ast.Module(body=stmts, type_ignores=[]), "<reference>", "exec")
        raise KeyError("loss_face_ij block")

    E1_tree = ast.parse(open(os.path.join(a.ref, E1)).read())
    cls_node = next(n for n in ast.walk(E1_tree) if isinstance(n, ast.ClassDef) and n.name == "FaceFeatsModel")
    ns_cls = {"torch": torch, "nn": torch.nn, "pkl": pickle, "util": util_stub}
    exec(compile(ast.Module(body=[cls_node], type_ignores=[]), "<reference>", "exec"), ns_cls)
    gen = torch.Generator().manual_seed(500)
    D, d, n = 300, 64, 23
    raw_db = torch.randn(D, d, generator=gen) * 2
    with tempfile.NamedTemporaryFile(suffix=".pkl") as tf:
        pickle.dump((raw_db, None, None), tf)
        tf.flush()
        model = ns_cls["FaceFeatsModel"](tf.name)
    out["face_db_raw"], out["face_db"] = raw_db.numpy(), model.face_feats.data.numpy()
    names = {1: (E1, ["targets_ij"], ["preds_gender_ori_ij"], ["probs_gender_ori_ij"]),
             2: (E3, ["targets_gender_ij", "targets_race_ij"], ["preds_gender_ori_ij", "preds_race_ori_ij"],
                 ["probs_gender_ori_ij", "probs_race_ori_ij"]),
             3: (E4, ["targets_gender_ij", "targets_race_ij", "targets_age_ij"],
                 ["preds_gender_ori_ij", "preds_race_ori_ij", "preds_age_ori_ij"],
                 ["probs_gender_ori_ij", "probs_race_ori_ij", "probs_age_ori_ij"])}
    widths = [2, 4, 2]
    for n_attr, (rel, tn, pn, qn) in names.items():
        code = loss_block(ast.parse(open(os.path.join(a.ref, rel)).read()))
        feats = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1)
        feats_ori = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=-1)
        face = torch.rand(n, generator=gen) > 0.2
        ns = {"torch": torch, "weight_dtype": torch.float32, "accelerator": types.SimpleNamespace(device="cpu"),
              "args": types.SimpleNamespace(face_gender_confidence_level=0.75, face_gender_race_confidence_level=0.75,
                                            face_gender_race_age_confidence_level=0.75),
              "idxs_ij": list(range(n)), "face_indicators_ij": face, "get_face_feats": lambda net, chips: chips,
              "face_feats_net": None, "aligned_face_chips_ij": feats, "face_feats_ori_ij": feats_ori, "face_feats_model": model}
        for k in range(n_attr):
            t = torch.randint(-1, widths[k], (n,), generator=gen)
            ns[tn[k]] = t
            ns[pn[k]] = torch.where(torch.rand(n, generator=gen) < 0.85, t.clamp(min=0), torch.randint(0, widths[k], (n,), generator=gen))
            ns[qn[k]] = torch.softmax(torch.randn(n, widths[k], generator=gen) * 4, -1)
            out[f"face_t{k}_{n_attr}"], out[f"face_p{k}_{n_attr}"], out[f"face_q{k}_{n_attr}"] = ns[tn[k]].numpy(), ns[pn[k]].numpy(), ns[qn[k]].numpy()
        exec(code, ns)
        out[f"face_feats_{n_attr}"], out[f"face_feats_ori_{n_attr}"], out[f"face_ind_{n_attr}"] = feats.numpy(), feats_ori.numpy(), face.numpy()
        out[f"face_loss_{n_attr}"] = ns["loss_face_ij"].numpy()
    # semantic_search on its own, with a selector and similarities
    q = torch.nn.functional.normalize(torch.randn(9, d, generator=gen), dim=-1)
    sel = torch.tensor([True, False, True, True, False, True, True, True, False])
    tgt, sim = model.semantic_search(q, sel, return_similarity=True)
    out["search_q"], out["search_sel"], out["search_target"], out["search_sim"] = q.numpy(), sel.numpy(), tgt.numpy(), sim.numpy()
    np.savez_compressed(os.path.join(HERE, "nextrows.npz"), **out)
    print("wrote nextrows.npz:", sorted(out))


if __name__ == "__main__":
    main()
