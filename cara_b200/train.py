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
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction) as one kernel over the flat buffer.

    The learning rate and the step count live in a 4-float DEVICE tensor (``state``: lr, steps taken, exit ticket,
    unused) that the kernel reads and advances itself, so the launch can be captured into the step's CUDA graph;
    ``param_groups[0]["lr"]`` stays the scheduler-facing value and is uploaded (stream-ordered) whenever it changed."""

    def __init__(self, flat: FlatTrainable, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
        self.flat = flat
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.m = torch.zeros_like(flat.flat)
        self.v = torch.zeros_like(flat.flat)
        self.t = 0                         # host-side count of the steps issued (the device holds the one that is used)
        self.param_groups = [{"lr": lr}]   # scheduler-facing, like torch.optim
        self.state = torch.zeros(4, device=flat.flat.device, dtype=torch.float32)
        self._lr_uploaded = None

    def zero_grad(self, set_to_none=False):
        self.flat.zero_grad()

    def sync_lr(self):
        """Upload ``param_groups[0]["lr"]`` if it changed since the last upload (never inside a graph capture: the
        captured kernel reads whatever the device tensor holds at replay time)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_uploaded:
            self.state[0:1].copy_(torch.tensor([lr], dtype=torch.float32), non_blocking=False)
            self._lr_uploaded = lr

    def launch(self, grad_scale=1.0):
        """The kernel launch alone (capturable); callers run ``sync_lr()`` before the step / replay."""
        K.adamw_step_dev(self.flat.flat, self.flat.grad, self.m, self.v, self.state, self.betas, self.eps,
                         self.weight_decay, gscale=grad_scale)

    def step(self, grad_scale=1.0):
        self.sync_lr()
        self.t += 1
        self.launch(grad_scale)


def cosine_lr(epoch, base_lr=1e-3, t_initial=100, warmup_t=10, lr_min=1e-5, warmup_lr_init=1e-6, decay_rate=0.1,
              cycle_limit=0):
    """timm 0.4.12 ``CosineLRScheduler(t_initial=100, warmup_t=10, lr_min=1e-5, warmup_lr_init=1e-6, decay_rate=0.1)``
    (``_get_lr`` with t_mul 1, no warmup prefix, cycle_limit 0 = restart every ``t_initial``) evaluated at an epoch
    index, which is how vit_cp.py:55-56,187 steps it.  The reference's loop only ever asks for epochs 0..50."""
    if epoch < warmup_t:
        return warmup_lr_init + epoch * (base_lr - warmup_lr_init) / warmup_t
    i = epoch // t_initial
    if cycle_limit > 0 and i >= cycle_limit:
        return lr_min
    gamma = decay_rate ** i
    lo, hi = lr_min * gamma, base_lr * gamma
    return lo + 0.5 * (hi - lo) * (1.0 + math.cos(math.pi * (epoch - t_initial * i) / t_initial))


class EpochCosineSchedule:
    """The learning rate of every optimizer step of the reference's loop (vit_cp.py:26-59,187).

    timm's scheduler writes ``warmup_lr_init`` into the optimizer when it is constructed, and the loop calls
    ``sched.step(epoch)`` AFTER each ``opt.step()``: the first batch of epoch e therefore still runs at the value set
    during epoch e-1 (all of epoch 0 at 1e-6), every other batch at ``cosine_lr(e)``.  After the periodic test of an
    epoch >= 50 the scheduler is dropped (``sched = None``, vit_cp.py:58-59) and the rate stays where it was."""

    def __init__(self, base_lr=1e-3, **kw):
        self.base_lr, self.kw, self.active = base_lr, kw, True
        self.lr = cosine_lr(0, base_lr=base_lr, **kw)      # what CosineLRScheduler.__init__ leaves in the optimizer

    def after_step(self, epoch):
        """``sched.step(epoch)`` (vit_cp.py:55-56); returns the rate of the NEXT optimizer step."""
        if self.active:
            self.lr = cosine_lr(epoch, base_lr=self.base_lr, **self.kw)
        return self.lr

    def after_test(self, epoch):
        """vit_cp.py:57-59: the scheduler is dropped at the first periodic test with epoch >= 50."""
        if epoch >= 50:
            self.active = False


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
    """``train_step`` replayed from ONE CUDA graph: zero_grad + forward + cross-entropy + backward + the gradient
    all-reduce (NCCL, captured) + the fused AdamW kernel (learning rate and step count read from device memory).

    A step is ~1,500 kernel launches whose host-side enqueue (autograd + ctypes) takes longer than half of the
    device time; the batch shape of vit_cp.py's loop is fixed (vtab.py:84-88, drop_last), so the whole launch
    sequence is captured once and replayed -- the host's only per-step work is the copy of the inputs into the
    graph's static buffers and one ``cudaGraphLaunch``.  With micro-batch accumulation (``accumulate`` = k > 1) the
    forward/backward graph is replayed k times and a second small graph holds the all-reduce + AdamW.
    On several GPUs the all-reduce and the optimizer stay outside the graph (two eager launches) unless
    ``CARA_GRAPH_COLLECTIVE=1``.
    """

    def __init__(self, model, opt, x, y, world_size=1, warmup=3, accumulate=1, capture_update=None):
        """``accumulate`` = k > 1: every call takes k micro-batches of the captured shape (x: [k*B, ...]) and sums
        their gradients before the single all-reduce + AdamW step (activations are kept for one micro-batch only:
        ViT-L at 1,024 images per GPU does not fit otherwise).  The loss returned is the mean over the k replays."""
        import os
        self.model, self.opt, self.world_size, self.accumulate = model, opt, world_size, int(accumulate)
        if capture_update is None:
            # one GPU: AdamW rides in the graph.  Several GPUs: the NCCL all-reduce (and AdamW after it) stay eager
            # launches unless CARA_GRAPH_COLLECTIVE=1 -- a process group cannot be destroyed while a captured graph
            # still holds its collectives (the caller must drop the GraphedStep first; see release()).
            env = os.environ.get("CARA_GRAPH_COLLECTIVE")
            capture_update = (world_size == 1) if env is None else env != "0"
        self.capture_update = bool(capture_update)
        self.gscale = 1.0 / (self.world_size * self.accumulate)
        self.x = torch.empty_like(x)
        self.y = torch.empty_like(y)
        self.x.copy_(x)
        self.y.copy_(y)
        opt.sync_lr()
        side = torch.cuda.Stream(device=x.device)
        side.wait_stream(torch.cuda.current_stream(x.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # eager runs: caches, lazily configured kernels, allocator warm-up
                self._fwd_bwd()
            if self.capture_update and world_size > 1:   # the communicator must exist before a capture can use it
                allreduce_grads(opt.flat, world_size)
        torch.cuda.current_stream(x.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        self.update_graph = None
        before = K.launch_count
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
            if self.capture_update and self.accumulate == 1:
                self._update()
        self.launches_per_replay = K.launch_count - before   # cara_* kernels inside the graph (bench: gpu_launches)
        self.launches_per_update = 0
        if self.capture_update and self.accumulate > 1:
            self.update_graph = torch.cuda.CUDAGraph()
            before = K.launch_count
            with torch.cuda.graph(self.update_graph):
                self._update()
            self.launches_per_update = K.launch_count - before

    def release(self):
        """Drop the captured graphs (needed before ``torch.distributed.destroy_process_group()`` when the all-reduce
        was captured: NCCL waits for every graph that references the communicator)."""
        self.graph = None
        self.update_graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def _fwd_bwd(self):
        if self.accumulate == 1:
            self.opt.zero_grad()              # inside the graph; with accumulation it happens once per call instead
        out = self.model(self.x)
        loss = torch.nn.functional.cross_entropy(out, self.y)
        loss.backward()
        return loss.detach()

    def _update(self):
        allreduce_grads(self.opt.flat, self.world_size)
        self.opt.launch(self.gscale)

    def __call__(self, x, y):
        k, B = self.accumulate, self.x.shape[0]
        if x.shape[0] != k * B or x.shape[1:] != self.x.shape[1:] or y.shape[0] != k * B:
            raise ValueError("GraphedStep was captured for %d micro-batch(es) of shape %s" % (k, tuple(self.x.shape)))
        self.opt.sync_lr()
        self.opt.t += 1
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
        if not self.capture_update:
            self._update()
        elif self.update_graph is not None:
            self.update_graph.replay()
            K.launch_count += self.launches_per_update
        if self.capture_update:
            K.param_generation += 1           # the captured AdamW changed the parameters behind autograd's back
        return loss
