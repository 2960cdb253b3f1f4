"""CPU: the oracle restatements against the golden vectors produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from util import ROOT, load_cfg, rel

GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name", ["tiny", "tiny3"])
def test_denoiser_restatement_matches_reference_golden(name):
    import ref_model
    g = np.load(os.path.join(GOLD, f"denoiser_{name}.npz"))
    cfg = load_cfg(name)
    sd = ref_model.random_state_dict(cfg, seed=int(g["seed"]))
    assert abs(sum(float(v.double().sum()) for v in sd.values()) - float(g["weight_checksum"])) < 1e-6 * abs(float(g["weight_checksum"])) + 1e-3
    sd = {k: v.requires_grad_(v.is_floating_point() and "inv_freq" not in k) for k, v in sd.items()}
    x0, noise, t = torch.from_numpy(g["x0"]), torch.from_numpy(g["noise"]), torch.from_numpy(g["t"])
    ids, mask = torch.from_numpy(g["ids"]), torch.from_numpy(g["mask"])
    xt = ref_model.add_noise(x0, noise, t)
    assert rel(xt, torch.from_numpy(g["xt"])) < 1e-6                       # DDPMScheduler.add_noise
    loss, pred = ref_model.train_step_loss(sd, cfg, x0, noise, t, ids, mask)
    assert rel(pred, torch.from_numpy(g["pred"])) < 1e-5                   # north-star fp32 tolerance
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * float(g["loss"])
    loss.backward()
    names = [str(n) for n in g["grad_names"]]
    for k, n_ref in zip(names, g["grad_norms"]):
        gr = sd[k].grad
        n = 0.0 if gr is None else float(gr.double().norm())
        assert abs(n - n_ref) <= 2e-5 * max(n_ref, 1e-8) + 1e-10, (k, n, n_ref)
    for key in g.files:
        if key.startswith("grad::"):
            assert rel(sd[key[6:]].grad, torch.from_numpy(g[key])) < 1e-5, key
    # DDPM sampling step (diffusers 0.15 DDPMScheduler.step with set_timesteps(100))
    prev = ref_model.ddpm_step(torch.from_numpy(g["pred"]), int(g["ddpm_t"]), torch.from_numpy(g["xt"]), torch.from_numpy(g["ddpm_noise"]))
    assert rel(prev, torch.from_numpy(g["ddpm_prev"])) < 1e-6


def test_param_table_matches_full_config():
    import ref_model
    shapes = ref_model.param_shapes(load_cfg("1d_config"))
    assert len(shapes) == 740                                                # 739 params + inv_freq
    n = sum(int(np.prod(s)) for k, s in shapes.items() if "inv_freq" not in k)
    assert n == 536_767_432


def _rvq_inputs(g):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "oracle", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    seed, B, T, grid = int(g["seed"]), int(g["B"]), int(g["T"]), bool(g["grid"])
    return mg.rvq_codebooks(seed, grid), mg.rvq_latents(seed, B, T, grid)


@pytest.mark.parametrize("name", ["grid", "gauss"])
def test_rvq_oracle_matches_encodec_golden(name):
    import rvq_oracle
    g = np.load(os.path.join(GOLD, f"rvq_{name}.npz"))
    cb, lat = _rvq_inputs(g)
    codes = rvq_oracle.encode(lat, cb)
    ref = g["codes"].astype(np.int64)
    if name == "grid":      # every partial sum exactly representable: order-independent, must be identical
        assert np.array_equal(codes, ref)
    else:                   # Gaussian: identical wherever the fp64 top-2 margin is not a near-tie
        c64, margin = rvq_oracle.encode_fp64(lat, cb)
        clear = np.minimum.accumulate(margin, axis=1) > 1e-4      # a flip cascades to later stages
        assert np.array_equal(codes[clear], ref[clear])
        assert (codes != ref).mean() < 1e-3
    dec = rvq_oracle.decode(ref, cb)
    assert np.array_equal(dec[:, :, : g["dec_slice"].shape[2]], g["dec_slice"])
    assert abs(dec.astype(np.float64).sum() - float(g["dec_sum"])) < 1e-9 * float(g["dec_abs_sum"]) + 1e-12


def test_codes_affine_oracle_matches_dataloader_formula():
    import rvq_oracle
    codes = np.arange(1024, dtype=np.int64)
    # tts/dataloader.py:64,77,168-170: torchvision Normalize(0.5, 0.5) on FloatTensor(codes / 1023)
    ref = ((torch.from_numpy(codes).float() / 1023) - 0.5) / 0.5
    assert np.array_equal(rvq_oracle.codes_affine(codes), ref.numpy())
