"""The two "next" rows of SURVEY.md 8f: detector staging (E1:1317/1326) and get_evaluate_metrics (E3:1716, E4:1780).
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
