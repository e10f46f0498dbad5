"""Training-step plumbing around the kernels: flat trainable buffer, fused AdamW, LR schedule, data-parallel
gradient exchange.

The reference's step is vit_cp.py:45-50 (forward, CE, zero_grad, backward, AdamW over ``CP*`` + ``head``
with lr 1e-3 / wd 1e-4, vit_cp.py:176-187).  Here the ~120 k trainable scalars live in ONE flat fp32 buffer
(parameters and ``.grad`` are views into it) so that the optimizer is one fused kernel launch and the
data-parallel exchange is one NCCL all-reduce of <= 1.1 MB per step; the frozen backbone is
replicated and never communicated.
"""
import math

import torch

from . import kernels as K


def freeze_backbone(model):
    """vit_cp.py:176-182: trainable iff the parameter name contains "CP" or "head"."""
    trainable = []
    for n, p in model.named_parameters():
        if "CP" in n or "head" in n:
            p.requires_grad = True
            trainable.append((n, p))
        else:
            p.requires_grad = False
    return trainable


class FlatTrainable:
    """Re-homes the trainable parameters into one contiguous fp32 buffer (and their grads into another)."""

    def __init__(self, named_params):
        self.named = list(named_params)
        total = sum(p.numel() for _, p in self.named)
        dev = self.named[0][1].device
        self.flat = torch.empty(total, device=dev, dtype=torch.float32)
        self.grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.slices = {}
        off = 0
        for n, p in self.named:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view(p.shape)
            p.grad = self.grad[off:off + k].view(p.shape)
            self.slices[n] = (off, off + k)
            off += k

    def zero_grad(self):
        self.grad.zero_()
        for _, p in self.named:           # re-attach if someone set grads to None
            if p.grad is None or p.grad.data_ptr() < self.grad.data_ptr() or \
               p.grad.data_ptr() >= self.grad.data_ptr() + self.grad.numel() * 4:
                a, b = self.slices[_]
                p.grad = self.grad[a:b].view(p.shape)

    def nbytes(self):
        return self.flat.numel() * 4


class FusedAdamW:
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction) as one kernel over the flat buffer."""

    def __init__(self, flat: FlatTrainable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.t = 0
        self.param_groups = [{"lr": lr}]   # scheduler-facing, like torch.optim

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def step(self, grad_scale=1.0):
        self.t += 1
        lr = self.param_groups[0]["lr"]
        K.adamw_step(self.flat.flat, self.flat.grad, self.m, self.v, lr, self.t, self.betas, self.eps,
                     self.weight_decay, gscale=grad_scale)


def cosine_lr(epoch, base_lr=1e-3, t_initial=100, warmup_t=10, lr_min=1e-5, warmup_lr_init=1e-6, decay_rate=0.1):
    """timm ``CosineLRScheduler(t_initial=100, warmup_t=10, lr_min=1e-5, warmup_lr_init=1e-6, decay_rate=0.1)``
    evaluated at an epoch index, as vit_cp.py:55-56,187 steps it (cycle_limit 1, no warmup prefix)."""
    if epoch < warmup_t:
        return warmup_lr_init + epoch * (base_lr - warmup_lr_init) / warmup_t
    i = epoch // t_initial
    if i >= 1:
        return lr_min
    lr_max = base_lr * (decay_rate ** i)
    return lr_min + 0.5 * (lr_max - lr_min) * (1.0 + math.cos(math.pi * (epoch - t_initial * i) / t_initial))


def allreduce_grads(flat: FlatTrainable, world_size):
    """One in-place sum all-reduce of the flat CP+head gradient (NCCL over NVLink on GPUs, gloo in CPU tests);
    the 1/world_size of the global-batch mean is applied inside the fused AdamW kernel (grad_scale)."""
    if world_size > 1:
        torch.distributed.all_reduce(flat.grad, op=torch.distributed.ReduceOp.SUM)


def shard_batch(global_batch, rank, world_size):
    """Contiguous even split of the global batch over the ranks (frozen backbone replicated)."""
    if global_batch % world_size != 0:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world_size))
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


def train_step(model, opt, x, y, world_size=1):
    """One vit_cp.py:45-50 iteration on this rank's shard; returns the (local) loss tensor."""
    out = model(x)
    loss = torch.nn.functional.cross_entropy(out, y)
    opt.zero_grad()
    loss.backward()
    allreduce_grads(opt.flat, world_size)
    opt.step(grad_scale=1.0 / world_size)
    return loss


class GraphedStep:
    """``train_step`` with zero_grad + forward + cross-entropy + backward replayed from ONE CUDA graph.

    A step is ~1,800 kernel launches whose host-side enqueue (autograd + ctypes) takes longer than half of the
    device time; the batch shape of vit_cp.py's loop is fixed (vtab.py:84-88, drop_last), so the whole launch
    sequence is captured once and replayed.  The gradient all-reduce and the fused AdamW kernel stay outside the
    graph (two launches; the optimizer's step count and learning rate are host scalars that change every step).
    Inputs are copied into the graph's static buffers on the launching stream before each replay.
    """

    def __init__(self, model, opt, x, y, world_size=1, warmup=3, accumulate=1):
        """``accumulate`` = k > 1: every call takes k micro-batches of the captured shape (x: [k*B, ...]) and sums
        their gradients before the single all-reduce + AdamW step (activations are kept for one micro-batch only:
        ViT-L at 1,024 images per GPU does not fit otherwise).  The loss returned is the mean over the k replays."""
        self.model, self.opt, self.world_size, self.accumulate = model, opt, world_size, int(accumulate)
        self.x = torch.empty_like(x)
        self.y = torch.empty_like(y)
        self.x.copy_(x)
        self.y.copy_(y)
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # eager runs: caches, lazily configured kernels, allocator warm-up
                self._fwd_bwd()
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        before = K.launch_count
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
        self.launches_per_replay = K.launch_count - before   # cara_* kernels inside the graph (bench: gpu_launches)

    def _fwd_bwd(self):
        if self.accumulate == 1:
            self.opt.zero_grad()              # inside the graph; with accumulation it happens once per call instead
        out = self.model(self.x)
        loss = torch.nn.functional.cross_entropy(out, self.y)
        loss.backward()
        return loss.detach()

    def __call__(self, x, y):
        k, B = self.accumulate, self.x.shape[0]
        if x.shape[0] != k * B or x.shape[1:] != self.x.shape[1:] or y.shape[0] != k * B:
            raise ValueError("GraphedStep was captured for %d micro-batch(es) of shape %s" % (k, tuple(self.x.shape)))
        if k == 1:
            if x.data_ptr() != self.x.data_ptr():
                self.x.copy_(x, non_blocking=True)
                self.y.copy_(y, non_blocking=True)
            self.graph.replay()
            K.launch_count += self.launches_per_replay
            loss = self.loss
        else:
            self.opt.zero_grad()
            loss = None
            for i in range(k):
                self.x.copy_(x[i * B:(i + 1) * B], non_blocking=True)
                self.y.copy_(y[i * B:(i + 1) * B], non_blocking=True)
                self.graph.replay()
                K.launch_count += self.launches_per_replay
                loss = self.loss.clone() if loss is None else loss + self.loss
            loss = loss / k
        allreduce_grads(self.opt.flat, self.world_size)
        self.opt.step(grad_scale=1.0 / (self.world_size * k))
        return loss
