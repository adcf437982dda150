"""CPU tests of the oracle itself: every layer of oracle/attention_oracle.py against the committed golden vectors
(ground truth of the reference's own checks, see oracle/make_golden.py) and against each other."""
import glob
import math
import os

import numpy as np
import pytest
import torch

from oracle import attention_oracle as orc


def _load(path):
    z = np.load(path)
    t = {k: torch.from_numpy(z[k]) for k in z.files if z[k].ndim > 0}
    return t, bool(z["causal"]), float(z["scale"])


def _golden_files(golden_dir, prefix):
    files = sorted(glob.glob(os.path.join(golden_dir, prefix + "*.npz")))
    assert files, "golden fixtures missing: run python oracle/make_golden.py"
    return files


def test_golden_present(golden_dir):
    assert len(_golden_files(golden_dir, "sdpa_")) >= 5
    assert len(_golden_files(golden_dir, "tutorial_")) >= 1


@pytest.mark.parametrize("idx", range(5))
def test_fp64_restatement_matches_reference_ground_truth(golden_dir, idx):
    """attention_fp64 / attention_grads_fp64 vs the reference's SDPA ground truth, at the reference's own
    tolerances (test_correctness.py:40,60-62)."""
    t, causal, scale = _load(_golden_files(golden_dir, "sdpa_")[idx])
    g = orc.attention_grads_fp64(t["Q"], t["K"], t["V"], t["dO"], scale, causal)
    assert torch.allclose(g["O"].float(), t["O"], atol=1e-4, rtol=1e-5)
    assert torch.allclose(g["dQ"].float(), t["dQ"], atol=9e-4, rtol=1e-5)
    assert torch.allclose(g["dK"].float(), t["dK"], atol=7e-4, rtol=1e-5)
    assert torch.allclose(g["dV"].float(), t["dV"], atol=7e-5, rtol=1e-5)
    # L is log2(e) * logsumexp
    assert torch.allclose((g["L"].squeeze(-1) * math.log(2.0)).float(), t["lse"], atol=1e-5, rtol=1e-6)


@pytest.mark.parametrize("idx", range(5))
def test_sdpa_wrapper_reproduces_golden(golden_dir, idx):
    """reference_sdpa(_grads) is the same call the fixtures were generated with."""
    t, causal, scale = _load(_golden_files(golden_dir, "sdpa_")[idx])
    O, dQ, dK, dV = orc.reference_sdpa_grads(t["Q"], t["K"], t["V"], t["dO"], scale, causal)
    for got, key in ((O, "O"), (dQ, "dQ"), (dK, "dK"), (dV, "dV")):
        assert torch.allclose(got, t[key], atol=2e-5, rtol=1e-5), key


def test_tutorial_golden(golden_dir):
    """fp16 causal, sm_scale 0.5 (flash_attention_openai_tutorial.py:523-559): atol 1e-2, rtol 0."""
    t, causal, scale = _load(_golden_files(golden_dir, "tutorial_")[0])
    assert causal and scale == 0.5
    g = orc.attention_grads_fp64(t["Q"], t["K"], t["V"], t["dO"], scale, causal)
    for key in ("O", "dQ", "dK", "dV"):
        assert torch.allclose(g[key].float(), t[key], atol=1e-2, rtol=0), key


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("tiles", [(16, 16), (32, 16), (16, 64)])
def test_tiled_online_softmax_matches_closed_form(causal, tiles):
    torch.manual_seed(7)
    Q, K, V = (torch.randn(1, 2, 80, 24) for _ in range(3))
    O64, L64 = orc.attention_fp64(Q, K, V, 0.3, causal)
    O, L = orc.attention_tiled(Q, K, V, 0.3, causal, B_r=tiles[0], B_c=tiles[1])
    assert torch.allclose(O.double(), O64, atol=2e-6)
    assert torch.allclose(L.double(), L64, atol=2e-5)


def test_tiled_low_precision_p_stays_within_north_star_tolerance():
    torch.manual_seed(3)
    Q, K, V = (orc.round_trip(torch.randn(1, 1, 128, 64), torch.bfloat16) for _ in range(3))
    # non-causal: every row averages 128 keys.  (Causal rows with one or two keys have |O| ~ |V| and bf16
    # rounding of P alone costs 2^-9 * |V| > 2e-3 there — the tolerance trap of SURVEY.md §0-10.)
    O64, _ = orc.attention_fp64(Q, K, V, 0.125, False)
    O, _ = orc.attention_tiled(Q, K, V, 0.125, False, 64, 64, p_dtype="bf16")
    assert (O.double() - O64).abs().max() < 2e-3
    Oc64, _ = orc.attention_fp64(Q, K, V, 0.125, True)
    Oc, _ = orc.attention_tiled(Q, K, V, 0.125, True, 64, 64, p_dtype="bf16")
    assert ((Oc.double() - Oc64).abs() <= 2e-3 + 2.0 ** -8 * Oc64.abs()).all()


def test_baseline_config1_runs_on_cpu():
    """BASELINE.json configs[0]: reference torch attention fwd fp32 B=1 H=4 N=512 D=64 non-causal on CPU."""
    torch.manual_seed(0)
    Q, K, V = (torch.randn(1, 4, 512, 64) for _ in range(3))
    O = orc.reference_sdpa(Q, K, V, 1.0, False)
    O64, L64 = orc.attention_fp64(Q, K, V, 1.0, False)
    assert torch.allclose(O.double(), O64, atol=1e-4, rtol=1e-5)
    assert L64.shape == (1, 4, 512, 1)


def test_flop_accounting():
    # flash_attention_openai_tutorial.py:630-636 ; BASELINE.md §3
    assert orc.attention_flops(2, 32, 8192, 128, True, "fwd") == pytest.approx(1.0995e12, rel=1e-3)
    assert orc.attention_flops(2, 32, 8192, 128, True, "fwd_bwd") == pytest.approx(3.848e12, rel=1e-3)
    assert orc.attention_flops(4, 16, 4096, 64, False, "fwd") == pytest.approx(2.749e11, rel=1e-3)
