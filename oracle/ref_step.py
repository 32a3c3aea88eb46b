"""One train iteration of the UNMODIFIED reference (models, losses, optimiser settings from oracle/_ref/t2i_moe_gan.py).

TEST / BENCH INFRASTRUCTURE ONLY.  ``train_aurora_gan`` (t2i_moe_gan.py:1029-1669) cannot be called offline: it
constructs ``AuroraGANLoss`` (-> ``clip.load``, :1124) and evaluates the CLIP loss (:1385-1386).  This file therefore
restates ONE iteration of its batch loop (:1262-1421) statement by statement around the reference's own
``AuroraGenerator`` / ``AuroraDiscriminator`` / ``AuroraGANLoss`` objects, leaving out only the CLIP terms
(:1384-1387) - the same omission as the B200 trainer, so both arms compute the same losses.  Host synchronisations of
the reference loop (`.item()` NaN guards :1315, :1372, :1396; ``kl_loss > 50.0`` :1369) are kept: they are part of what
the reference costs.  gradient_accumulation_steps = 1; schedule constants are those of epoch 0 unless set_epoch() is
called (:1168-1178).
"""
from __future__ import annotations

from contextlib import nullcontext
from typing import Dict, Optional

import torch

from . import ref_loader


class ReferenceTrainStep:
    def __init__(self, device, amp_dtype: Optional[torch.dtype] = None, checkpoint_activation: bool = False, seed: int = 0,
                 lr: float = 2e-4, beta1: float = 0.5, beta2: float = 0.999, r1_gamma: float = 10.0, kl_weight: float = 1e-3,
                 balance_weight: float = 0.01, generator=None, discriminator=None):
        self.ref = ref = ref_loader.load_reference()
        self.device = torch.device(device)
        torch.manual_seed(seed)
        self.G = generator if generator is not None else ref.AuroraGenerator(max_resolution=16).to(self.device)      # (:1091)
        self.D = discriminator if discriminator is not None else ref.AuroraDiscriminator(max_resolution=16).to(self.device)
        if checkpoint_activation:
            self.G.enable_checkpointing()                                                      # (:1094-1097)
        wd = 0.01                                                                              # (:1100-1102)
        self.opt_g = torch.optim.AdamW(self.G.parameters(), lr=lr, betas=(beta1, beta2), weight_decay=wd)
        self.opt_d = torch.optim.AdamW(self.D.parameters(), lr=lr, betas=(beta1, beta2), weight_decay=wd)
        self.amp_dtype = amp_dtype
        # the reference's own AMP is fp16 autocast + GradScaler (:1121, :1278); bf16 needs no scaler
        self.scaler = torch.amp.GradScaler("cuda") if (amp_dtype == torch.float16 and self.device.type == "cuda") else None
        self.gan_loss = ref_loader.make_loss(ref, self.device)
        self.r1_gamma, self.kl_weight, self.balance_weight = r1_gamma, kl_weight, balance_weight
        self.set_epoch(0)

    def set_epoch(self, epoch: int, kl_annealing_epochs: int = 5) -> None:
        kl_factor = min(1.0, (epoch / kl_annealing_epochs) ** 2)                                # (:1168-1174)
        self.effective_kl_weight = self.kl_weight * (1e-5 + (1.0 - 1e-5) * kl_factor)
        self.temperature_factor = max(1.0, 3.0 - epoch * 0.1)                                   # (:1178)

    def _autocast(self):
        if self.amp_dtype is None:
            return nullcontext()
        return torch.autocast(device_type=self.device.type, dtype=self.amp_dtype)

    def step(self, real_images: torch.Tensor, text_embeddings: torch.Tensor, z: Optional[torch.Tensor] = None,
             shuffled_indices: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        ref, G, D, gan_loss, device = self.ref, self.G, self.D, self.gan_loss, self.device
        scaler = self.scaler
        real_images = real_images.to(device).detach().clone()                                   # (:1262-1263)
        text_embeddings = text_embeddings.to(device)
        batch_size = real_images.size(0)
        if z is None:
            z = torch.randn(batch_size, ref.LATENT_DIM, device=device)                          # (:1266)

        # ---- discriminator (:1268-1348) ----
        self.opt_d.zero_grad()
        real_images.requires_grad = True
        with self._autocast():
            real_pred = D(real_images, text_embeddings)
            grad_real, = torch.autograd.grad(outputs=real_pred.sum(), inputs=real_images, create_graph=True)
            grad_penalty = grad_real.view(grad_real.size(0), -1).norm(2, dim=1) ** 2
            r1_loss = (self.r1_gamma / 2) * grad_penalty.mean()
            with torch.no_grad():
                fake_images_16, fake_images_8, _, _ = G(z, text_embeddings, return_intermediate=True, return_routing=True,
                                                        annealing_factor=self.temperature_factor)
            fake_pred = D(fake_images_16.detach(), text_embeddings)
            if shuffled_indices is None:
                shuffled_indices = torch.randperm(batch_size)                                  # (:1303)
            mismatched_pred = D(real_images.detach(), text_embeddings[shuffled_indices])
            d_loss_gan = gan_loss.discriminator_loss(real_pred, fake_pred, mismatched_pred)
            d_loss = d_loss_gan + r1_loss
        if torch.isnan(d_loss).item() or torch.isinf(d_loss).item():                            # (:1315)
            return {"skipped": torch.ones(())}
        if scaler:
            scaler.scale(d_loss).backward(retain_graph=True)
            scaler.unscale_(self.opt_d)
            torch.nn.utils.clip_grad_norm_(D.parameters(), max_norm=0.7)
            scaler.step(self.opt_d)
            scaler.update()
        else:
            d_loss.backward(retain_graph=True)
            torch.nn.utils.clip_grad_norm_(D.parameters(), max_norm=0.7)                        # (:1336)
            self.opt_d.step()
        out = {"d_loss": d_loss_gan.detach(), "r1": r1_loss.detach()}
        del grad_real, grad_penalty, real_pred, fake_pred, mismatched_pred

        # ---- generator (:1350-1421) ----
        self.opt_g.zero_grad()
        with self._autocast():
            fake_images_16, fake_images_8, kl_loss, routing_probs = G(z, text_embeddings, return_intermediate=True, return_routing=True,
                                                                       annealing_factor=self.temperature_factor)
            if kl_loss is not None:
                if kl_loss > 50.0:                                                              # host sync (:1369)
                    kl_loss = torch.clamp(kl_loss, max=50.0)
                if torch.isnan(kl_loss).item() or torch.isinf(kl_loss).item():
                    kl_loss = torch.tensor(0.0, device=device, requires_grad=kl_loss.requires_grad)
            fake_pred_g = D(fake_images_16, text_embeddings)
            g_loss_gan = gan_loss.generator_loss(fake_pred_g)
            balance_loss = gan_loss.moe_balance_loss(routing_probs, balance_weight=self.balance_weight)   # (:1390)
            g_loss = g_loss_gan + balance_loss                                                  # CLIP terms (:1384-1387) omitted
            if torch.isnan(g_loss).item() or torch.isinf(g_loss).item():                        # (:1396)
                g_loss = torch.tensor(0.0, device=device, requires_grad=g_loss.requires_grad)
            if kl_loss is not None:
                g_loss = g_loss + self.effective_kl_weight * kl_loss                            # (:1402-1404)
        if scaler:
            scaler.scale(g_loss).backward()
            scaler.unscale_(self.opt_g)
            torch.nn.utils.clip_grad_norm_(G.parameters(), max_norm=0.8)
            scaler.step(self.opt_g)
            scaler.update()
        else:
            g_loss.backward()
            torch.nn.utils.clip_grad_norm_(G.parameters(), max_norm=0.8)                        # (:1420)
            self.opt_g.step()
        out.update({"g_loss": g_loss_gan.detach(), "kl": kl_loss.detach(), "balance": balance_loss.detach()})
        return out
