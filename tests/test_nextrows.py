"""The "next" rows of SURVEY.md 8f: detector staging (E1:1317/1326), get_evaluate_metrics (E3:1716, E4:1780) and the
E6 enumerated-composition assignment (E6:1413-1482).
Golden vectors come from the reference's own statements (tests/golden/make_golden_next.py)."""
import os

import numpy as np
import pytest
import torch

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nextrows.npz"))
DTYPES = {"float32": torch.float32, "bfloat16": torch.bfloat16, "float16": torch.float16}


def _metric_case(tag, dname):
    key = f"metrics_{tag}_{dname}"
    dt = DTYPES[dname]
    args = [torch.tensor(GOLD[key + s]).to(dt) for s in ("_pg", "_pr", "_pa")]
    return (args if tag == "e4" else args[:2]), GOLD[key + "_out"]


@pytest.mark.parametrize("dname", list(DTYPES))
@pytest.mark.parametrize("tag", ["e3", "e4"])
def test_oracle_metrics_golden(tag, dname):
    from oracle import nextrows
    args, want = _metric_case(tag, dname)
    got = np.array(nextrows.evaluate_metrics(*args))
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-7)


@pytest.mark.parametrize("dname", list(DTYPES))
def test_oracle_staging_golden(dname):
    from oracle import nextrows
    x = torch.tensor(GOLD[f"stage_{dname}_in"]).to(DTYPES[dname])
    assert np.array_equal(nextrows.stage_detector_input(x), GOLD[f"stage_{dname}_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("dname", list(DTYPES))
@pytest.mark.parametrize("tag", ["e3", "e4"])
def test_gpu_metrics_golden(tag, dname):
    import fairguide as fg
    args, want = _metric_case(tag, dname)
    got = np.array(fg.get_evaluate_metrics(*[a.cuda() for a in args]))
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-7)       # fp32 sums of <= 56 terms: order may differ in the last ulp


@pytest.mark.gpu
def test_gpu_metrics_large_vs_oracle():
    import fairguide as fg
    from oracle import nextrows
    g = torch.Generator().manual_seed(5)
    n = 4096
    ps = [torch.softmax(torch.randn(n, w, generator=g) * 2, -1).to(torch.bfloat16) for w in (2, 4, 2)]
    miss = torch.rand(n, generator=g) < 0.05
    for p in ps:
        p[miss] = -1
    want = np.array(nextrows.evaluate_metrics(*ps))
    got = np.array(fg.get_evaluate_metrics(*[p.cuda() for p in ps]))
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-7)
    # no valid row at all: means of empty selections are NaN in the reference as well
    empty = fg.get_evaluate_metrics(torch.full((3, 2), -1.0).cuda(), torch.full((3, 4), -1.0).cuda())
    assert all(np.isnan(v) for v in empty)


@pytest.mark.gpu
@pytest.mark.parametrize("dname", list(DTYPES))
def test_gpu_staging_golden(dname):
    import fairguide as fg
    x = torch.tensor(GOLD[f"stage_{dname}_in"]).to(DTYPES[dname]).cuda()
    got = fg.stage_detector_input(x)
    assert got.dtype == np.uint8 and np.array_equal(got, GOLD[f"stage_{dname}_out"])


@pytest.mark.gpu
@pytest.mark.parametrize("dname,shape", [("bfloat16", (4, 3, 512, 512)), ("float16", (2, 3, 64, 48)), ("float32", (3, 3, 33, 21))])
def test_gpu_staging_vs_oracle(dname, shape):
    """Full-size images (vector path) and odd shapes (scalar path), bit-exact, including values beyond [-1, 1]."""
    import fairguide as fg
    from oracle import nextrows
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(*shape, generator=g) * 2.2 - 1.1).to(DTYPES[dname])
    want = nextrows.stage_detector_input(x)
    got = fg.ops.stage_detector_input(x.cuda()).cpu().numpy()
    assert np.array_equal(got, want)


# ----------------------------------------------------------------------------- E6 enumerated-composition assignment
def test_oracle_race_assignment_golden():
    """The oracle's restatement of generate_dynamic_targets_race (E6:1413-1482) against the outputs of the reference's own
    function body (lifted with ast, `ot` stubbed: see make_golden_next.py)."""
    from oracle import assign as oassign
    for c in range(int(GOLD["race_n_cases"])):
        p = torch.tensor(GOLD[f"race_probs_{c}"])
        t, u = oassign.generate_dynamic_targets_race(p, True)
        assert np.array_equal(t.numpy(), GOLD[f"race_targets_{c}"]), c
        assert np.array_equal(u.numpy(), GOLD[f"race_unc_{c}"]), c
        assert torch.equal(oassign.generate_dynamic_targets_race(p), t)


def test_race_compositions_host_logic():
    """Product host code (api.race_compositions) == the oracle's statement-by-statement version, incl. the object-dtype
    regime (N >= 36: multinomial coefficients beyond int64) and the tie order at the 95 % cut."""
    import fairguide as fg
    from oracle import assign as oassign
    for N in (1, 2, 7, 24, 37):
        c1, w1 = oassign.race_compositions(N)
        c2, w2 = fg.api.race_compositions(N)
        assert np.array_equal(np.asarray(c1, dtype=np.int64), c2) and np.array_equal(np.asarray(w1, dtype=np.float64), w2)
        assert (c2.sum(1) == N).all() and w2.sum() > 0.95 and (np.diff(w2) <= 0).all()


@pytest.mark.gpu
def test_gpu_race_assignment_golden():
    import fairguide as fg
    for c in range(int(GOLD["race_n_cases"])):
        p = torch.tensor(GOLD[f"race_probs_{c}"]).cuda()
        t, u = fg.generate_dynamic_targets_race(p, True)
        assert np.array_equal(t.cpu().numpy(), GOLD[f"race_targets_{c}"]), c          # bit-exact targets
        assert np.array_equal(u.cpu().numpy(), GOLD[f"race_unc_{c}"]), c              # fp64 accumulation in the reference's order
        assert torch.equal(fg.generate_dynamic_targets_race(p), t)
        # fused thresholding == the caller's two lines
        t2, _ = fg.generate_dynamic_targets_race(p, True, uncertainty_threshold=0.2)
        ref = GOLD[f"race_targets_{c}"].copy(); ref[GOLD[f"race_unc_{c}"] > np.float32(0.2)] = -1
        assert np.array_equal(t2.cpu().numpy(), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("N,nmiss,sharp", [(1, 1, 2.0), (33, 5, 1.5), (48, 0, 2.5)])
def test_gpu_race_assignment_vs_oracle(N, nmiss, sharp):
    import fairguide as fg
    from oracle import assign as oassign
    g = torch.Generator().manual_seed(N)
    p = torch.softmax(torch.randn(N + nmiss, 4, generator=g) * sharp, -1)
    if nmiss:
        p[torch.randperm(N + nmiss, generator=g)[:nmiss]] = -1
    valid = (p != -1).all(-1)
    M = fg.ops.race_cost_matrix(p.cuda(), N).cpu().numpy()
    assert np.array_equal(M, oassign.pot_dist_euclidean(p[valid].numpy(), np.eye(4, dtype=np.int64)))   # fp64, bit-exact
    t, u = fg.generate_dynamic_targets_race(p.cuda(), True, num_valid=N)
    rt, ru = oassign.generate_dynamic_targets_race(p, True)
    assert torch.equal(t.cpu(), rt) and torch.equal(u.cpu(), ru)
    # every composition solved exactly: status word clean
    combs, w = fg.api.race_compositions(N)
    d16 = np.zeros((len(w), 16), np.int32); d16[:, :4] = combs
    _, _, ws = fg.ops.assign_race_enumerated(p.cuda(), N, torch.from_numpy(d16).cuda(), torch.from_numpy(w).cuda())
    assert ws[:4].view(torch.int32).item() == 0


@pytest.mark.gpu
def test_gpu_race_assignment_no_faces():
    import fairguide as fg
    p = torch.full((5, 4), -1.0).cuda()
    t, u = fg.generate_dynamic_targets_race(p, True)
    assert (t == -1).all() and (u == -1).all()


# ----------------------------------------------------------------------------- f4: adjusted-DFT coefficients, bucketed gradient sync
def test_adjusted_dft_coefs_golden():
    """Oracle and product host code against the reference's own statements (E1:1104-1109)."""
    import fairguide as fg
    from oracle import nextrows
    for c in range(int(GOLD["dft_n_cases"])):
        ac, al = torch.tensor(GOLD[f"dft_alphas_cumprod_{c}"]), torch.tensor(GOLD[f"dft_alphas_{c}"])
        ts = torch.tensor(GOLD[f"dft_timesteps_{c}"])
        want = GOLD[f"dft_coefs_{c}"]
        assert np.array_equal(nextrows.adjusted_dft_grad_coefs(ac, al, ts), want)
        assert np.array_equal(fg.adjusted_dft_grad_coefs(ac, al, ts), want)
        np.testing.assert_allclose(np.prod(want) ** (1 / len(want)), 1.0, rtol=1e-12)
    assert fg.make_grad_hook(0.5)(torch.tensor([2.0, 4.0])).tolist() == [1.0, 2.0]


def test_grad_bucket_layout():
    import fairguide as fg
    assert fg.GradBucket.layout([4, 1, 7]) == [0, 4, 5, 12]
    assert fg.GradBucket.layout([]) == [0]


def _lora_like_grads(seed, dtype, n_tensors=37):
    g = torch.Generator().manual_seed(seed)
    shapes = [(50, 320 + 64 * (k % 5)) if k % 2 == 0 else (320 + 64 * (k % 5), 50) for k in range(n_tensors)] + [(7,), (1,)]
    return [torch.randn(*s, generator=g).to(dtype) * 10 ** float(torch.randint(-3, 2, (1,), generator=g)) for s in shapes]


@pytest.mark.gpu
@pytest.mark.parametrize("dname", ["float32", "bfloat16"])
def test_gpu_grad_bucket_vs_oracle(dname):
    """Two simulated ranks on one GPU: pack each rank's gradients, add the buckets (what the all-reduce does), unpack with
    the two divisors; against the per-tensor loop of the reference."""
    import fairguide as fg
    from oracle import nextrows
    dt = DTYPES[dname]
    ga, gb = _lora_like_grads(1, dt), _lora_like_grads(2, dt)
    want, finite = nextrows.allreduce_average_gradients([[g.float() for g in ga], [g.float() for g in gb]], 3)
    pa = [torch.nn.Parameter(torch.zeros_like(g).cuda()) for g in ga]
    pb = [torch.nn.Parameter(torch.zeros_like(g).cuda()) for g in gb]
    for p, g in zip(pa + pb, ga + gb):
        p.grad = g.clone().cuda()
    A, B = fg.GradBucket(pa), fg.GradBucket(pb)
    A.pack(); B.pack()
    assert int(A.nonfinite.item()) == 0 and A.bucket[A.total].item() == 0
    flat = torch.cat([g.flatten().float() for g in ga])
    assert torch.equal(A.bucket[:-1].cpu(), flat)                                  # pack is a pure copy
    A.bucket.add_(B.bucket)
    A.unpack(2, 3)
    for p, w in zip(pa, want):
        if dt == torch.float32:
            # (x * (1/2)) * (1/3) in fp32 (ATen multiplies by the reciprocal of a scalar divisor) vs the oracle's two divisions
            np.testing.assert_allclose(p.grad.cpu().numpy(), w.numpy(), rtol=3e-7, atol=0)
        else:
            np.testing.assert_allclose(p.grad.float().cpu().numpy(), w.numpy(), rtol=2e-2, atol=0)
    # non-finite gradients are counted, and the count travels in the last slot of the bucket
    pb[3].grad[0, 5] = float("inf"); pb[10].grad[2, 2] = float("nan")
    B.pack()
    assert int(B.nonfinite.item()) == 2 and B.bucket[B.total].item() == 2.0
    # world size 1 through the drop-in entry point
    for p, g in zip(pa, ga):
        p.grad = g.clone().cuda()
    assert fg.allreduce_average_gradients(pa, n_backward=4) is True
    np.testing.assert_allclose(pa[0].grad.float().cpu().numpy(), (ga[0].float() / 1 / 4).numpy(), rtol=1e-2 if dt != torch.float32 else 3e-7)
    ok, total = B.sync(num_processes=1, n_backward=1)
    assert ok is False and total == 2
