"""Worker of tests/test_gpu_ddp.py: launched as 2 ranks (torchrun, NCCL), one GPU each.  Multi-GPU parity as SURVEY.md
section 8e defines it: the all-reduced gradient must equal the mean of the per-rank single-GPU gradients, checked by
replaying every rank's scenes on ONE GPU (rank 0) -- fp32 path, rel 1e-5.  Covers the flat and the bucketed/overlapped
reducer (three steps each: the bucketed one freezes its bucket order after the first)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_path):
    import occuseg_b200.sparseconvnet as scn
    from occuseg_b200 import scenes
    from occuseg_b200.ddp import BucketedGradAllReduce, FlatGradAllReduce
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    scn.set_precision("fp32")

    def make_net():
        torch.manual_seed(7)
        return scn.Sequential().add(scn.InputLayer(3, 4096, mode=4)).add(scn.SubmanifoldConvolution(3, 3, 16, 3, False)) \
            .add(scn.UNet(3, 1, [16, 32, 48], True)).add(scn.BatchNormReLU(16)).add(scn.OutputLayer(3)).to(dev)

    def batch(r, step):
        c, f = scenes.make_batch("tiny", (10 * step + 2 * r, 10 * step + 2 * r + 1))
        return [torch.from_numpy(c).float(), torch.from_numpy(f).to(dev), None, 2]

    results = {}
    for kind in ("flat", "bucketed"):
        net = make_net()
        red = FlatGradAllReduce(net.parameters(), world) if kind == "flat" else \
            BucketedGradAllReduce(net.parameters(), world, bucket_mb=0.05)
        worst = 0.0
        for step in range(3):
            net(batch(rank, step)).square().mean().backward()
            red.finish()
            torch.cuda.synchronize()
            got = [p.grad.detach().clone() for p in net.parameters()]
            if rank == 0:
                # replay every rank's scenes on this one GPU with the same weights (a fresh copy of the module tree, so the
                # running statistics of `net` are untouched) and average
                ref = make_net()
                ref.load_state_dict(net.state_dict())
                acc = [torch.zeros_like(p) for p in ref.parameters()]
                for r in range(world):
                    ref.zero_grad(set_to_none=True)
                    ref2 = ref
                    ref2(batch(r, step)).square().mean().backward()
                    for a, p in zip(acc, ref2.parameters()):
                        a += p.grad
                    ref.load_state_dict(net.state_dict())      # undo the running-statistics update of this replay
                for a, g, (name, _) in zip(acc, got, net.named_parameters()):
                    want = (a / world).cpu().numpy()
                    err = float(np.abs(g.cpu().numpy() - want).max() / max(np.abs(want).max(), 1e-30))
                    worst = max(worst, err)
            net.zero_grad(set_to_none=False)
        if kind == "bucketed" and rank == 0:
            results["n_buckets"] = len(red.buckets)
        results[kind] = worst
    # every rank holds identical reduced gradients
    flat = torch.cat([g.flatten() for g in got])
    both = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    results["ranks_agree"] = bool(all(torch.equal(both[0], b) for b in both))
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(results, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
