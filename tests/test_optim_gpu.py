"""GPU: the fused clip + AdamW step against torch's clip_grad_norm_ + AdamW -- the reference's optimiser
lines train.py:41-47,116-120 -- applied to the same gradients."""
import pytest
import torch

from util import load_cfg, rel, synth_inputs

pytestmark = pytest.mark.gpu


def test_fused_clip_adamw_matches_torch(cuda):
    from prompt_tts_b200.models import TTSSingleSpeaker
    from prompt_tts_b200.optim import FusedClipAdamW
    from prompt_tts_b200.train import DenoiserTrainStep
    cfg = load_cfg("tiny")
    torch.manual_seed(0)
    model = TTSSingleSpeaker(cfg).to(cuda)
    inp = synth_inputs(cfg, 2, 16, seed=2, device=cuda)
    stepper = DenoiserTrainStep(model)
    # lr large enough that two steps move the weights measurably; clip threshold below the actual norm so clipping is active
    kw = dict(lr=1e-3, betas=(0.95, 0.999), weight_decay=1e-2, eps=1e-8)
    opt = FusedClipAdamW(stepper, max_norm=0.05, **kw)
    ref_params = {k: p.detach().clone().cpu().requires_grad_(True) for k, p in model.named_parameters()}
    ref_opt = torch.optim.AdamW(list(ref_params.values()), **kw)
    names = [k for k, _ in model.named_parameters()]
    for it in range(3):
        for p in model.parameters():
            p.grad = None
        stepper(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"])
        for k, p in model.named_parameters():       # the same gradients go to the torch optimiser
            ref_params[k].grad = p.grad.detach().cpu().clone() if p.grad is not None else None
        live = [q for q in ref_params.values() if q.grad is not None]
        tn = torch.nn.utils.clip_grad_norm_(live, 0.05)
        ref_opt.step()
        gsq = opt.step()
        assert abs(float(gsq) ** 0.5 - float(tn)) / float(tn) < 1e-4
        assert float(tn) > 0.05, "the test must exercise the clipping branch"
    worst = max(rel(dict(model.named_parameters())[k].cpu(), ref_params[k]) for k in names)
    moved = max(rel(dict(model.named_parameters())[k].cpu() - ref_params[k].detach(), ref_params[k]) for k in names)
    assert worst < 1e-5, worst
    # state_dict still has the reference layout and reads the updated master buffer
    sd = model.state_dict()
    assert all(torch.equal(sd[k].cpu(), dict(model.named_parameters())[k].detach().cpu()) for k in names)
    # and the next forward uses the updated weights (packed copies refreshed): loss changes
    l0 = float(stepper.loss)
    for p in model.parameters():
        p.grad = None
    l1 = float(stepper(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"]))
    assert l1 != l0 and abs(l1 - l0) / l0 < 0.5
