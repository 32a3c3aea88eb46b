"""Data-parallel train step for MoE-GAN: one iteration of the reference loop body
(moegan/t2i_moe_gan.py:1262-1421, minus the CLIP terms :1384-1387 whose weights are not available offline)
with the B200 MoE blocks inside the generator.

What is new relative to the reference:
  * one process per GPU; gradients are all-reduced with NCCL in ~25 MB buckets launched from grad-ready hooks on a
    communication stream, so the transfer over NVLink overlaps the rest of backward (`GradReducer`);
  * the balance loss is made exact under data parallelism by all-reducing the [E] load vector (`dp_allreduce_load`);
  * no host synchronisation inside the step (the reference's `.item()` NaN guards and the `kl > 50` branch are
    expressed on the device), which lets the whole step (D-step + G-step + both AdamW updates) be captured in ONE
    CUDA graph and replayed (`use_cuda_graph=True`).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn

from . import host_model as hm
from .sparse_moe import moe_balance_loss
from .functional import BalanceLossFunction, bump_weight_generation


# --------------------------------------------------------------------------------------------------
# bucketed, overlapped gradient all-reduce
# --------------------------------------------------------------------------------------------------
def _flat_storage_view(t: torch.Tensor) -> Optional[torch.Tensor]:
    """1-D view of a dense tensor in STORAGE order (no copy): contiguous tensors and channels-last 4-D tensors (the conv
    weights / gradients of a channels_last model).  An all-reduce is element-wise, so any order works as long as pack
    and unpack use the same one and every rank has the same layout.  None if the tensor is not dense in either format."""
    if t.is_contiguous():
        return t.view(-1)
    if t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last):
        return t.permute(0, 2, 3, 1).view(-1)
    return None


class GradReducer:
    """Averages .grad of `params` across the process group, bucket by bucket, as soon as a bucket's gradients exist.

    Buckets are filled in reverse parameter order (the order backward produces gradients).  Each bucket is one flat
    fp32 buffer; when its last gradient has been accumulated the bucket is packed and all-reduced on `comm_stream`
    while the compute stream keeps running backward.  `finish()` joins the streams and scatters the averages back.
    """

    def __init__(self, params: Sequence[nn.Parameter], group=None, bucket_bytes: int = 25 << 20, overlap: bool = True,
                 tail_bytes: int = 8 << 20, comm_priority: Optional[int] = None):
        import os
        # MOEGAN_DP_MODE (bench / profiling): "overlap" (default) = buckets all-reduced on a communication stream while backward runs;
        # "single" = ONE all-reduce of all gradients after backward, on the compute stream (no SM contention with the persistent
        # 1-CTA-per-SM GEMMs of backward, which leave an overlapped NCCL kernel waiting for SMs: tools/dp_timeline.py)
        mode = os.environ.get("MOEGAN_DP_MODE", "overlap")
        if mode == "single":
            overlap, bucket_bytes, tail_bytes = False, 1 << 62, 0
        if comm_priority is None and os.environ.get("MOEGAN_DP_COMM_PRIORITY"):
            comm_priority = int(os.environ["MOEGAN_DP_COMM_PRIORITY"])
        self.comm_priority = comm_priority
        self.defer = mode == "single"          # launch nothing from the hooks: finish() packs and reduces everything at once
        # profiling only (tools/dp_diag3.sh): pack the buckets but launch no collective - the result is NOT a data-parallel step
        self.skip_collective = os.environ.get("MOEGAN_DP_SKIP_COLLECTIVE", "0") == "1"
        self._comm_used = False
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.overlap = overlap
        self.buckets: List[List[nn.Parameter]] = []
        cur, cur_bytes = [], 0
        for p in reversed(self.params):
            nbytes = p.numel() * 4
            if cur and cur_bytes + nbytes > bucket_bytes:
                self.buckets.append(cur); cur, cur_bytes = [], 0
            cur.append(p); cur_bytes += nbytes
        if cur:
            self.buckets.append(cur)
        # The LAST bucket holds the first parameters of the model, whose gradients appear at the very end of backward: nothing is left
        # to overlap its all-reduce with, so it is kept short - its trailing <= tail_bytes are split off into a bucket of their own and
        # the rest of it can go out (and overlap) earlier.
        if self.buckets and tail_bytes > 0:
            last = self.buckets[-1]
            tail, tb = [], 0
            while len(last) > 1 and tb + last[-1].numel() * 4 <= tail_bytes:
                q = last.pop(); tail.insert(0, q); tb += q.numel() * 4
            if tail and last:
                self.buckets.append(tail)
            elif tail:
                self.buckets[-1] = tail
        self.bucket_of = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self.flat: List[Optional[torch.Tensor]] = [None] * len(self.buckets)
        self.pending = [0] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.comm_stream = None
        self.enabled = False
        # NCCL averages inside the collective (ReduceOp.AVG); other backends (gloo in the CPU tests) sum, then scale
        self.avg_in_collective = bool(self.world > 1 and dist.get_backend(group) == "nccl")
        self._handles = []
        # profiling (tools/dp_timeline.py): a list to which every bucket appends CUDA events - gradients ready (compute stream),
        # all-reduce begin / end (comm stream) - so that the exposed part of the collective can be read off one step
        self.trace: Optional[list] = None
        if self.world > 1:
            for p in self.params:
                self._handles.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _flat(self, bi: int) -> torch.Tensor:
        if self.flat[bi] is None:
            n = sum(p.numel() for p in self.buckets[bi])
            self.flat[bi] = torch.empty(n, dtype=torch.float32, device=self.buckets[bi][0].device)
        return self.flat[bi]

    def begin(self) -> None:
        """Call before backward."""
        if self.world == 1:
            return
        self.enabled = True
        for bi, b in enumerate(self.buckets):
            self.pending[bi] = len(b)
            self.launched[bi] = False
        if self.overlap and self.buckets and self.buckets[0][0].is_cuda and self.comm_stream is None:
            # a high-priority stream lets the few CTAs of an NCCL kernel be placed ahead of the next compute kernel's CTAs
            self.comm_stream = torch.cuda.Stream(priority=self.comm_priority) if self.comm_priority is not None else torch.cuda.Stream()

    def _on_grad(self, p: nn.Parameter) -> None:
        if not self.enabled:
            return
        bi = self.bucket_of[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0 and not self.defer:
            self._launch(bi)

    def _grad_view(self, p: nn.Parameter, flat_slice: torch.Tensor) -> Optional[torch.Tensor]:
        """The slice of a bucket seen as p's gradient (same shape and strides as p), or None if p is dense in neither format."""
        if p.is_contiguous():
            return flat_slice.view(p.shape)
        if p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last):
            n, c, h, w = p.shape
            return flat_slice.view(n, h, w, c).permute(0, 3, 1, 2)
        return None

    def _launch(self, bi: int) -> None:
        self.launched[bi] = True
        flat = self._flat(bi)
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.buckets[bi]]
        views = flat.split([p.numel() for p in self.buckets[bi]])
        srcs = [_flat_storage_view(g) for g in grads]
        srcs = [s if s is not None else g.reshape(-1) for s, g in zip(srcs, grads)]
        # a gradient that already lives in its bucket slot (left there by the previous finish() and accumulated into in place) needs no copy
        todo = [(v, s) for v, s in zip(views, srcs) if v.data_ptr() != s.data_ptr()]
        if todo:
            torch._foreach_copy_([v for v, _ in todo], [s for _, s in todo])
        op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
        rec = None
        if self.trace is not None and flat.is_cuda:
            rec = {"bucket": bi, "bytes": flat.numel() * 4, "ready": torch.cuda.Event(enable_timing=True),
                   "begin": torch.cuda.Event(enable_timing=True), "end": torch.cuda.Event(enable_timing=True)}
            rec["ready"].record()
            self.trace.append(rec)
        if self.skip_collective:
            return
        if self.comm_stream is not None:
            self._comm_used = True
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm_stream):
                if rec is not None:
                    rec["begin"].record()
                dist.all_reduce(flat, op=op, group=self.group)
                if not self.avg_in_collective:
                    flat.mul_(1.0 / self.world)
                if rec is not None:
                    rec["end"].record()
        else:
            dist.all_reduce(flat, op=op, group=self.group)
            if not self.avg_in_collective:
                flat.mul_(1.0 / self.world)

    def finish(self) -> None:
        """Call after backward: reduces stragglers, joins the comm stream and makes every .grad a VIEW of its (averaged) bucket slot -
        nothing is copied back: clip_grad_norm_ and the optimizer read the reduced gradients where the collective left them."""
        if self.world == 1 or not self.enabled:
            return
        for bi in range(len(self.buckets)):
            if not self.launched[bi]:
                self._launch(bi)          # parameters that received no gradient this step
        if self.comm_stream is not None and self._comm_used:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._comm_used = False
        for bi, b in enumerate(self.buckets):
            views = self._flat(bi).split([p.numel() for p in b])
            for p, v in zip(b, views):
                gv = self._grad_view(p, v)
                if gv is not None:
                    p.grad = gv
                else:                                   # dense in neither memory format (none in this model): copy back
                    if p.grad is None:
                        p.grad = torch.empty_like(p)
                    p.grad.copy_(v.view_as(p))
        self.enabled = False


class _AllReduceLoad(torch.autograd.Function):
    """Global per-expert load for the balance loss.  forward: sum over ranks.  backward: x world, because the
    data-parallel gradient average divides by world a loss term that is already global."""

    @staticmethod
    def forward(ctx, load, group, world):
        ctx.world = world
        out = load.clone()
        dist.all_reduce(out, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        return g * ctx.world, None, None


def dp_allreduce_load(load: torch.Tensor, group=None) -> torch.Tensor:
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return load
    return _AllReduceLoad.apply(load, group, world)


# --------------------------------------------------------------------------------------------------
# schedule helpers (t2i_moe_gan.py:1168-1178)
# --------------------------------------------------------------------------------------------------
def effective_kl_weight(kl_weight: float, epoch: int, kl_annealing_epochs: int = 5) -> float:
    warm = min(1.0, (epoch / kl_annealing_epochs) ** 2)
    return kl_weight * (1e-5 + (1.0 - 1e-5) * warm)


def temperature_factor(epoch: int) -> float:
    return max(1.0, 3.0 - epoch * 0.1)


# --------------------------------------------------------------------------------------------------
# the train step
# --------------------------------------------------------------------------------------------------
class MoEGANTrainer:
    def __init__(self, generator: nn.Module, discriminator: nn.Module, lr: float = 2e-4, beta1: float = 0.5, beta2: float = 0.999,
                 weight_decay: float = 0.01, r1_gamma: float = 10.0, kl_weight: float = 1e-3, balance_weight: float = 0.01,
                 amp_dtype: Optional[torch.dtype] = torch.bfloat16, group=None, bucket_bytes: int = 25 << 20,
                 use_cuda_graph: bool = False, exact_dp_balance: bool = True, fused_real_pass: bool = True):
        self.G, self.D = generator, discriminator
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.r1_gamma, self.kl_weight, self.balance_weight = r1_gamma, kl_weight, balance_weight
        self.amp_dtype = amp_dtype
        self.exact_dp_balance = exact_dp_balance
        self.use_cuda_graph = use_cuda_graph
        # D(real), D(real, mismatched text) and R1 from one shared feature pass with an explicit adjoint (host_model.
        # discriminator_real_pass); False = the reference's three separate evaluations + autograd.grad(create_graph=True)
        self.fused_real_pass = fused_real_pass and isinstance(discriminator, hm.AuroraDiscriminator)
        on_cuda = next(self.G.parameters()).is_cuda
        cap = dict(capturable=True) if use_cuda_graph else {}
        if on_cuda:
            cap["fused"] = True        # one multi-tensor kernel per optimizer instead of ~10 foreach passes
        # AdamW(lr 2e-4, betas (0.5, 0.999), wd 0.01) for both networks (:1100-1102)
        self.opt_g = torch.optim.AdamW(self.G.parameters(), lr=lr, betas=(beta1, beta2), weight_decay=weight_decay, **cap)
        self.opt_d = torch.optim.AdamW(self.D.parameters(), lr=lr, betas=(beta1, beta2), weight_decay=weight_decay, **cap)
        self.red_g = GradReducer(list(self.G.parameters()), group, bucket_bytes)
        self.red_d = GradReducer(list(self.D.parameters()), group, bucket_bytes)
        self.set_epoch(0)
        self._graph = None
        self._static: Dict[str, torch.Tensor] = {}
        self._out: Dict[str, torch.Tensor] = {}

    def set_epoch(self, epoch: int, kl_annealing_epochs: int = 5) -> None:
        self.annealing_factor = temperature_factor(epoch)
        self.eff_kl_weight = effective_kl_weight(self.kl_weight, epoch, kl_annealing_epochs)
        self._graph = None    # schedule constants are baked into a captured graph

    def _autocast(self):
        dev = next(self.G.parameters()).device.type
        return torch.autocast(device_type=dev, dtype=self.amp_dtype, enabled=self.amp_dtype is not None)

    # ---- D-step (:1268-1348) ----
    def d_step(self, real: torch.Tensor, text: torch.Tensor, z: torch.Tensor, perm: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        self.opt_d.zero_grad(set_to_none=True)
        real = real.detach()
        with self._autocast():
            if perm is None:
                perm = torch.rand(text.shape[0], device=text.device).argsort()      # device-side randperm (:1303)
            if self.fused_real_pass:
                real_pred, mismatched_pred, r1 = hm.discriminator_real_pass(self.D, real, text, text[perm], self.r1_gamma)
            else:
                real = real.requires_grad_(True)
                real_pred = self.D(real, text)
                r1 = hm.r1_penalty(real_pred, real, self.r1_gamma)
                mismatched_pred = self.D(real.detach(), text[perm])
            with torch.no_grad():      # module stays in train mode: dense-soft routing, eps redrawn (:1289-1298)
                fake16, _, _, _ = self.G(z, text, return_intermediate=True, return_routing=True, annealing_factor=self.annealing_factor)
            fake_pred = self.D(fake16.detach(), text)
            d_gan = hm.discriminator_adv_loss(real_pred, fake_pred, mismatched_pred)
            d_loss = d_gan + r1
        self.red_d.begin()
        d_loss.backward()
        self.red_d.finish()
        torch.nn.utils.clip_grad_norm_(self.D.parameters(), max_norm=0.7)           # (:1336), after the all-reduce
        self.opt_d.step()
        return {"d_loss": d_gan.detach(), "r1": r1.detach()}

    # ---- G-step (:1350-1421) ----
    def g_step(self, text: torch.Tensor, z: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.opt_g.zero_grad(set_to_none=True)
        with self._autocast():
            fake16, fake8, kl, routing = self.G(z, text, return_intermediate=True, return_routing=True,
                                                annealing_factor=self.annealing_factor)
            kl = torch.clamp(kl, max=50.0)                                          # `if kl > 50: clamp(max=50)` without the sync (:1369)
            kl = torch.nan_to_num(kl, nan=0.0, posinf=50.0, neginf=0.0)
            fake_pred = self.D(fake16, text)
            g_gan = hm.generator_adv_loss(fake_pred)
            last = routing[-1]
            load = getattr(last, "_moegan_load", None)
            if load is None:
                load = last.float().sum(dim=0)
            T_glob = last.shape[0] * (self.world if self.exact_dp_balance else 1)
            if self.exact_dp_balance:
                load = dp_allreduce_load(load, self.group)
            balance = BalanceLossFunction.apply(load, T_glob, self.balance_weight)
            g_loss = g_gan + balance + self.eff_kl_weight * kl                       # (:1393, :1402-1404)
        self.red_g.begin()
        g_loss.backward()
        self.red_g.finish()
        torch.nn.utils.clip_grad_norm_(self.G.parameters(), max_norm=0.8)           # (:1420)
        self.opt_g.step()
        return {"g_loss": g_gan.detach(), "kl": kl.detach(), "balance": balance.detach()}

    def _step_impl(self, real: torch.Tensor, text: torch.Tensor, z: Optional[torch.Tensor] = None,
                   perm: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """z / perm: caller-supplied latent batch and mismatched-text permutation (parity runs against the reference step, which
        draws them at :1266 / :1303); None = drawn here on the device."""
        if z is None:
            z = torch.randn(real.shape[0], hm.LATENT_DIM, device=real.device)       # (:1266)
        out = self.d_step(real, text, z, perm)
        out.update(self.g_step(text, z))
        return out

    def step(self, real: torch.Tensor, text: torch.Tensor) -> Dict[str, torch.Tensor]:
        """One train iteration on device-resident `real` [B,3,64,64] and `text` [B,512]."""
        if not self.use_cuda_graph:
            return self._step_impl(real, text)
        if self._graph is None or self._static["real"].shape != real.shape:
            self._capture(real, text)
        self._static["real"].copy_(real, non_blocking=True)
        self._static["text"].copy_(text, non_blocking=True)
        self._graph.replay()
        bump_weight_generation()     # the captured AdamW wrote the expert weights without bumping Parameter._version
        return self._out

    def step_from_host(self, real_host: torch.Tensor, text_host: torch.Tensor) -> Dict[str, float]:
        """The end-to-end call: pinned host batch in, host floats out (H2D copy + step + D2H loss read)."""
        dev = next(self.G.parameters()).device
        if self.use_cuda_graph and self._graph is not None and self._static["real"].shape == real_host.shape:
            self._static["real"].copy_(real_host, non_blocking=True)
            self._static["text"].copy_(text_host, non_blocking=True)
            self._graph.replay()
            bump_weight_generation()
            out = self._out
        else:
            out = self.step(real_host.to(dev, non_blocking=True), text_host.to(dev, non_blocking=True))
        packed = torch.stack([out["d_loss"].float(), out["g_loss"].float()]).cpu()
        return {"d_loss": float(packed[0]), "g_loss": float(packed[1])}

    def _training_state(self) -> List[torch.Tensor]:
        """Every tensor a train step mutates: parameters and buffers of both networks, and the optimizer states that exist."""
        ts = [t for m in (self.G, self.D) for t in list(m.parameters()) + list(m.buffers())]
        for opt in (self.opt_g, self.opt_d):
            for st in opt.state.values():
                ts += [v for v in st.values() if torch.is_tensor(v)]
        return ts

    def _capture(self, real: torch.Tensor, text: torch.Tensor) -> None:
        """Capture one step into a CUDA graph WITHOUT advancing training: the three warm-up steps capture needs (lazy
        initialisation, cuDNN autotuning, optimizer state allocation, NCCL) run on the live tensors - a graph bakes their
        addresses in - and are then undone in place: parameters / buffers / existing optimizer state are restored from a
        snapshot, optimizer state the warm-up created is zeroed (= AdamW's initial state, step 0).  A re-capture after
        set_epoch() therefore costs no optimizer steps (the reference takes none there either)."""
        self._static = {"real": real.clone(), "text": text.clone()}
        before = self._training_state()
        snapshot = [t.detach().clone() for t in before]
        known = {id(t) for t in before}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                       # warm-up outside capture: lazy inits, optimizer state, NCCL
                self._step_impl(self._static["real"], self._static["text"])
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        # Capture on the SAME stream the warm-up ran on.  Under data parallelism the gradient hooks keep every AccumulateGrad node
        # alive, and autograd pins such a node to the stream it was created on (the warm-up stream): capturing on another stream makes
        # every one of the ~300 gradient accumulations hop streams through an event pair inside the graph (torch warns "AccumulateGrad
        # node's stream does not match") - measured, that was most of the fixed ~1.1-1.7 ms per step that data parallelism cost.
        with torch.cuda.graph(g, stream=side):
            self._out = self._step_impl(self._static["real"], self._static["text"])
        with torch.no_grad():
            for t, s0 in zip(before, snapshot):
                t.copy_(s0)
            for t in self._training_state():
                if id(t) not in known:
                    t.zero_()
        bump_weight_generation()
        torch.cuda.synchronize()
        self._graph = g
