"""Worker of tests/test_dp_gpu.py (one process per GPU, launched by torch.distributed.run): N-rank GradSync gradients against a
single-process step on the concatenated batch -- the reference's DDP contract (train.py:25-29,67-69,100-117)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import load_cfg, rel, synth_inputs  # noqa: E402


def main():
    from prompt_tts_b200.dp import GradSync
    from prompt_tts_b200.models import TTSSingleSpeaker
    from prompt_tts_b200.optim import FusedClipAdamW
    from prompt_tts_b200.train import DenoiserTrainStep
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg_name, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    cfg = load_cfg(cfg_name)
    out = {"world": world, "cfg": cfg_name, "B_per_rank": B, "T": T}
    for mode in ("fp32", "bf16"):
        torch.manual_seed(1 + rank)                     # ranks start from DIFFERENT weights: GradSync must broadcast rank 0's
        model = TTSSingleSpeaker(cfg).to(dev)
        gs = GradSync(model, world_size=world, bucket_mb=0.25, comm_dtype=torch.float32 if mode == "fp32" else torch.bfloat16)
        stepper = DenoiserTrainStep(model, grad_sync=gs)
        inp = synth_inputs(cfg, B, T, seed=100 + rank, device=dev)
        for _ in range(2):                               # step 1 learns the layout (one blocking reduce); step 2 is the bucketed, overlapped path
            stepper(inp["x0"], inp["noise"], inp["t"], inp["ids"], inp["mask"])
        torch.cuda.synchronize()
        buckets = gs.n_buckets_last
        flat = gs.flat.clone()
        # every rank holds the same averaged gradient
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        # single process, concatenated batch, same (rank 0) weights
        if rank == 0:
            torch.manual_seed(1)
            ref_model = TTSSingleSpeaker(cfg).to(dev)
            parts = [synth_inputs(cfg, B, T, seed=100 + r, device=dev) for r in range(world)]
            cat = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
            ref = DenoiserTrainStep(ref_model)
            for _ in range(2):
                ref(cat["x0"], cat["noise"], cat["t"], cat["ids"], cat["mask"])
            torch.cuda.synchronize()
            assert [g[1:] for g in ref.grad_sync.groups] == [g[1:] for g in gs.groups]
            out[mode] = {"rel_flat_vs_single_process": rel(flat, ref.grad_sync.flat), "identical_on_all_ranks": same, "buckets": buckets,
                         "grad_norm": float(flat.norm())}
        # one optimiser step on every rank: weights stay identical across ranks
        opt = FusedClipAdamW(stepper, lr=1e-3)
        opt.step()
        chk = opt.pflat.double().sum().reshape(1)
        allc = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        if rank == 0:
            out[mode]["weights_identical_after_step"] = all(torch.equal(allc[0], c) for c in allc)
        del model, stepper, gs, opt
    if rank == 0:
        print("DPRESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
