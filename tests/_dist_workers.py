"""Worker bodies of the multi-rank tests (one process per rank).

* ``cpu_worker``  -- gloo, world 2, no GPU: the data-parallel host logic of cara_b200.train (shard_batch, the flat
  gradient buffer, allreduce_grads, the 1/world grad scale) on a small differentiable stand-in for the model.
* ``__main__``    -- launched by torchrun with one rank per GPU (NCCL): the real CUDA step.  Two ranks on half
  batches (GraphedStep: all-reduce + AdamW captured in the graph, or eager with CARA_GRAPH_COLLECTIVE=0) must take the
  same optimizer step as one rank on the full batch.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _toy_params(seed=0):
    g = torch.Generator().manual_seed(seed)
    return [("CP_A2", torch.randn(6, 4, generator=g)), ("CP_R1", torch.randn(4, generator=g)),
            ("head.weight", torch.randn(3, 6, generator=g)), ("head.bias", torch.randn(3, generator=g))]


def _toy_loss(params, x, y):
    """A CaRA-shaped toy: logits = ((x A) * r) A^T W^T + b; mean cross-entropy over the rows given."""
    p = dict(params)
    h = ((x @ p["CP_A2"]) * p["CP_R1"]) @ p["CP_A2"].t()
    return torch.nn.functional.cross_entropy(h @ p["head.weight"].t() + p["head.bias"], y)


def cpu_worker(rank, world, port, out_dir):
    """gloo world-``world``: per-rank shard -> backward into the flat gradient -> allreduce_grads -> AdamW with
    grad_scale 1/world (oracle update).  Writes the resulting parameters; the parent compares them with a single
    full-batch step."""
    import torch.distributed as dist
    from cara_b200 import train as T
    from oracle import cara_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        params = [(n, torch.nn.Parameter(t.clone())) for n, t in _toy_params()]
        flat = T.FlatTrainable(params)
        g = torch.Generator().manual_seed(99)
        X, Y = torch.randn(8, 6, generator=g), torch.randint(0, 3, (8,), generator=g)
        lo, hi = T.shard_batch(X.shape[0], rank, world)
        flat.zero_grad()
        _toy_loss(params, X[lo:hi], Y[lo:hi]).backward()          # accumulates into the flat buffer's views
        local = flat.grad.clone()
        T.allreduce_grads(flat, world)
        new = {}
        for n, p in params:
            a, b = flat.slices[n]
            gr = flat.grad[a:b].view(p.shape) * (1.0 / world)      # what the fused kernel's grad_scale applies
            new[n] = O.adamw_update(p.detach(), gr, torch.zeros_like(gr), torch.zeros_like(gr), 1)[0]
        torch.save({"params": new, "local_grad": local, "summed_grad": flat.grad.clone(), "range": (lo, hi)},
                   os.path.join(out_dir, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


def _gpu_main():
    import torch.distributed as dist
    from cara_b200 import train as T
    from cara_b200.vit import create_model
    from oracle import cara_oracle as O
    from src.cara.cara import cara
    import warnings
    warnings.simplefilter("ignore")

    import signal
    signal.alarm(int(os.environ.get("CARA_DIST_ALARM", "400")))      # never outlive a hung collective
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    depth = int(os.environ.get("CARA_DIST_DEPTH", "2"))
    rank_r = int(os.environ.get("CARA_DIST_RANK", "8"))
    per = int(os.environ.get("CARA_DIST_PER_RANK", "4"))
    steps = int(os.environ.get("CARA_DIST_STEPS", "2"))
    g = O.Geometry(depth=depth, rank=rank_r, num_classes=10)
    st = O.synthetic_state(g)

    def build():
        vit = create_model("vit_base_patch16_224_in21k", depth=g.depth)
        vit = cara({"model": vit, "rank": g.rank, "scale": 1.0, "l_mu": 1.0, "l_std": 0.0})
        vit.reset_classifier(g.num_classes)
        vit.load_state_dict(st, strict=True)
        vit = vit.to(dev)
        vit.train()                                      # drop_path 0 in this model: deterministic up to atomics
        opt = T.FusedAdamW(T.FlatTrainable(T.freeze_backbone(vit)), lr=1e-3, weight_decay=1e-4)
        return vit, opt

    batches = [O.synthetic_batch(g, per * world, seed=500 + i) for i in range(steps)]
    # data-parallel: this rank's shard through the graphed step (collective + AdamW inside the graph by default)
    vit, opt = build()
    lo, hi = T.shard_batch(per * world, rank, world)
    x0, y0 = batches[0]
    step = T.GraphedStep(vit, opt, x0[lo:hi].to(dev), y0[lo:hi].to(dev), world)
    grads_dp, losses_dp = [], []
    for x, y in batches:
        losses_dp.append(float(step(x[lo:hi].to(dev), y[lo:hi].to(dev))))
        torch.cuda.synchronize()
        grads_dp.append(opt.flat.grad.detach().clone() / world)     # the graph leaves the all-reduced SUM in the buffer
    p_dp = opt.flat.flat.detach().clone()
    # single rank, full batch, eager
    vit1, opt1 = build()
    grads_1, losses_1 = [], []
    for x, y in batches:
        losses_1.append(float(T.train_step(vit1, opt1, x.to(dev), y.to(dev), 1)))
        grads_1.append(opt1.flat.grad.detach().clone())
    p_1 = opt1.flat.flat.detach().clone()
    torch.cuda.synchronize()

    def rel(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    # every rank holds the same parameters after the step
    gathered = [torch.empty_like(p_dp) for _ in range(world)]
    dist.all_gather(gathered, p_dp)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    # the global loss is the mean of the shard losses
    lt = torch.tensor(losses_dp, device=dev)
    dist.all_reduce(lt)
    lt = (lt / world).tolist()
    e_g = rel(grads_dp[0], grads_1[0])                   # same parameters, same samples: only fp32 summation order differs
    e_g_later = max([rel(a, b) for a, b in zip(grads_dp[1:], grads_1[1:])] + [0.0])
    e_p = rel(p_dp, p_1)
    e_l = max(abs(a - b) for a, b in zip(lt, losses_1))
    steps_dev = float(opt.state[1])
    # Adam's first steps move every element by ~lr * sign(g): elements whose gradient is rounding noise may flip, so
    # after the first update the two runs hold slightly different parameters (and later gradients)
    ok = same and e_g <= 1e-5 and e_g_later <= 2e-2 and e_p <= 1e-3 and e_l <= 1e-3 and steps_dev == float(steps)
    if rank == 0:
        print("DIST world=%d per_rank=%d depth=%d captured_update=%s: first-step grad rel %.3e (later steps %.3e), "
              "param rel %.3e, loss abs %.3e, ranks identical %s, device step count %.0f -> %s"
              % (world, per, depth, step.capture_update, e_g, e_g_later, e_p, e_l, same, steps_dev,
                 "DIST_OK" if ok else "DIST_FAIL"), flush=True)
    sys.stdout.flush()
    step.release()                                       # the graph holds the captured all-reduce
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    _gpu_main()
