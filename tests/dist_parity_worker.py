"""torchrun worker of tests/test_dist_gpu.py: NCCL data-parallel gradients through Trainer(graph=True) == gradients of
the concatenated batch on one GPU (SURVEY 7.5 "distributed" tier). Exit code 0 = parity."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from vit_torch_b200 import models, train
    from vit_torch_b200.dist import GradAllReducer
    mode = sys.argv[1] if len(sys.argv) > 1 else "deferred"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    bs = 4
    g = torch.Generator().manual_seed(123)
    X = torch.randn((world * bs, 3, 96, 96), generator=g)
    Y = torch.randint(0, 10, (world * bs,), generator=g)

    def make():
        torch.manual_seed(7)
        m = models.dino_vits16(pretrained=False).to(dev)
        train.reset_parameters_like_zoo(m)
        return m

    m = make()
    red = GradAllReducer(m, overlap=(mode == "overlap"), compress=(mode == "bf16"), split=(mode == "split"),
                         register_arena=(mode == "registered"))
    tr = train.Trainer(m, lr=0.0, momentum=0.0, reducer=red, graph=(mode != "overlap"), strict_graph=True)
    x, y = X[rank * bs:(rank + 1) * bs].to(dev), Y[rank * bs:(rank + 1) * bs].to(dev)
    for _ in range(4):                       # 2 eager steps, the capture step, one replay (lr 0: weights unchanged)
        tr.step(x, y)
    if mode == "registered" and rank == 0:
        print(f"arena registered with NCCL: {red.arena_registered}", flush=True)
    if mode == "split":
        assert tr._graph2 is not None and 0 < tr._split_off < tr.arena.off, "backward was not split into two graphs"
    torch.cuda.synchronize()
    got = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    red.close()
    ok = True
    if rank == 0:
        ref = make()
        tr1 = train.Trainer(ref, lr=0.0, momentum=0.0, graph=False)
        tr1.step(X.to(dev), Y.to(dev))
        want = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
        assert sorted(got) == sorted(want), (sorted(set(want) - set(got)), sorted(set(got) - set(want)))
        tol = 2e-2 if mode == "bf16" else 1e-3
        worst = 0.0
        for k in want:
            e = ((got[k] - want[k]).abs().max() / want[k].abs().max().clamp_min(1e-20)).item()
            worst = max(worst, e)
            if not e <= tol:
                print(f"MISMATCH {k}: {e:.3e}", flush=True)
                ok = False
        print(f"dist parity [{mode}] world {world}: worst normalised error {worst:.3e} over {len(want)} tensors; "
              f"collectives per step {red.collectives / 4:.2f}", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
