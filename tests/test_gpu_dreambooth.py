"""GPU parity: fused DreamBooth-PSO loss+grad kernel vs the restated trainer code
(train_pso_sdxl_turbo_dreambooth.py:1847-1865,1881-1935) and its fp64 closed form."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as olosses, make_golden
from tests import _util as U

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pso(built_lib):
    import pairwise_sample_optimization_b200 as pso
    return pso


def _c(a, dtype=None):
    return U.cuda(torch.from_numpy(np.asarray(a)), dtype)


@pytest.mark.parametrize("name", ["dreambooth_pso", "dreambooth_pso_db", "dreambooth_pso_noprior"])
def test_dreambooth_vs_fixture(pso, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    mp = _c(g["model_pred"]).requires_grad_(True)
    loss, lw, ll, logits = pso.pso_db_loss(mp, _c(g["ref_pred"]), _c(g["noisy"]), _c(g["x0"]), _c(g["sigmas"]),
                                           loss_type=str(g["loss_type"]), beta_pso=float(g["beta_pso"]),
                                           neg_defactor=float(g["neg_defactor"]),
                                           prior_loss_weight=float(g["prior_loss_weight"]))
    loss.backward()
    assert abs(loss.item() - float(g["loss_fp64"])) <= 1e-5 * abs(float(g["loss_fp64"]))
    assert abs(loss.item() - float(g["loss"])) <= 2e-5 * abs(float(g["loss"]))
    np.testing.assert_allclose(lw.cpu().numpy(), g["losses_w_fp64"], rtol=1e-5)
    np.testing.assert_allclose(ll.cpu().numpy(), g["losses_l_fp64"], rtol=1e-5)
    np.testing.assert_allclose(logits.cpu().numpy(), g["logits_fp64"], rtol=1e-5, atol=1e-9)
    assert U.rel_max(mp.grad, torch.from_numpy(g["grad_fp64"])) <= 1e-5
    assert U.rel_max(mp.grad, torch.from_numpy(g["grad"])) <= 1e-4
    pso.check_status()


@pytest.mark.parametrize("loss_type", ["pso", "pso_db"])
@pytest.mark.parametrize("pd,ld", [("bf16", "bf16"), ("fp16", "fp16"), ("fp32", "fp16")])
def test_dreambooth_config4_half_storage(pso, loss_type, pd, ld):
    """BASELINE config 4: batch 4 pairs per GPU, 64x64 latents."""
    d = make_golden.synth_dreambooth(4, (4, 64, 64), 51)
    mp, rp = d["model_pred"].to(U.DT[pd]), d["ref_pred"].to(U.DT[pd])
    nz, x0 = d["noisy"].to(U.DT[ld]), d["x0"].to(U.DT[ld])
    cf = olosses.dreambooth_closed_form(mp.double(), rp.double(), nz.double(), x0.double(), d["sigmas"], loss_type,
                                        5.0, 0.1, 0.3)
    mpc = U.cuda(mp).requires_grad_(True)
    loss, lw, ll, logits = pso.pso_db_loss(mpc, U.cuda(rp), U.cuda(nz), U.cuda(x0), U.cuda(d["sigmas"]),
                                           loss_type=loss_type, beta_pso=5.0, neg_defactor=0.1, prior_loss_weight=0.3)
    loss.backward()
    assert abs(loss.item() - cf["loss"].item()) <= 1e-5 * abs(cf["loss"].item())
    np.testing.assert_allclose(logits.cpu().numpy(), cf["logits"].numpy(), rtol=2e-5, atol=1e-8)
    if pd == "fp32":
        assert U.rel_max(mpc.grad, cf["grad"]) <= 1e-5
    else:
        U.assert_rounded_equal(mpc.grad, cf["grad"], U.DT[pd])


def test_dreambooth_hinge_inactive_and_odd_rows(pso):
    d = make_golden.synth_dreambooth(2, (4, 16, 16), 52)
    # pso_db with a tiny beta keeps 1 - beta*logits > 0; with a huge negative margin the hinge closes
    mp = U.cuda(d["model_pred"]).requires_grad_(True)
    loss, lw, ll, logits = pso.pso_db_loss(mp, None, U.cuda(d["noisy"]), U.cuda(d["x0"]), U.cuda(d["sigmas"]),
                                           loss_type="pso_db", beta_pso=-1e4, neg_defactor=0.0, prior_loss_weight=0.0)
    loss.backward()
    cf = olosses.dreambooth_closed_form(d["model_pred"], None, d["noisy"], d["x0"], d["sigmas"], "pso_db", -1e4, 0.0, 0.0)
    assert abs(loss.item() - cf["loss"].item()) <= 1e-5 * max(1.0, abs(cf["loss"].item()))
    assert U.rel_max(mp.grad, cf["grad"]) <= 1e-5 or float(cf["grad"].abs().max()) == 0.0
    with pytest.raises(Exception):
        pso.pso_db_loss(mp[:3], None, U.cuda(d["noisy"])[:3], U.cuda(d["x0"])[:3], U.cuda(d["sigmas"])[:3], loss_type="pso_db")
