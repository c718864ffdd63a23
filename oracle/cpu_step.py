"""ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's training step (runners/holycow.py:224-257) restated on top of
oracle/reference_model.py with plain torch CPU ops: E -> G -> D x3 -> criteria, loss_G.backward(retain_graph),
optimizer_G.step, loss_D.backward, optimizer_D.step, EMA.  Used (a) by tests as the step-level checker and (b) by
bench.py as the CPU baseline / `--impl reference` arm (the reference itself is Python code that cannot travel to the
GPU box; this port runs the same torch operators the reference's modules would call).
"""
import math

import torch

from . import reference_model as R
from . import synth


class RAdamPort:
    """utils/radam.py:29-95 (the optimizer configs/finetuning-base.yaml selects), per-tensor form."""

    def __init__(self, params, lr, betas, eps):
        self.params, self.lr, self.betas, self.eps = list(params), lr, betas, eps
        self.state = [dict(step=0, m=torch.zeros_like(p), v=torch.zeros_like(p)) for p in self.params]

    @torch.no_grad()
    def step(self):
        b1, b2 = self.betas
        for p, st in zip(self.params, self.state):
            if p.grad is None:
                continue
            g = p.grad
            st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
            st["m"].mul_(b1).add_(g, alpha=1 - b1)
            st["step"] += 1
            t = st["step"]
            b2t = b2 ** t
            n_max = 2 / (1 - b2) - 1
            n_sma = n_max - 2 * t * b2t / (1 - b2t)
            if n_sma >= 5:
                step_size = math.sqrt((1 - b2t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) \
                    / (1 - b1 ** t)
                p.addcdiv_(st["m"], st["v"].sqrt().add_(self.eps), value=-step_size * self.lr)
            else:
                p.add_(st["m"], alpha=-self.lr / (1 - b1 ** t))

    def zero_grad(self):
        for p in self.params:
            p.grad = None


class OracleTrainer:
    def __init__(self, cfg, finetune=True, criteria=("adversarial", "featmat", "idt_embed", "perceptual", "dice"),
                 optimizer="RAdam", lr_gen=5e-4, lr_dis=8e-4, beta1=0.0, seed=1, with_embedder=True):
        import torchvision
        self.cfg, self.finetune, self.criteria = dict(cfg), finetune, tuple(criteria)
        if finetune:
            self.cfg["num_labels"] = 1
            self.cfg["embed_eps"] = 1e-12
        g_sd = synth.generator_state_dict(cfg, seed=seed, finetuned=finetune)
        d_sd = synth.discriminator_state_dict(self.cfg, seed=seed + 1, finetuned=finetune)

        def leafify(sd):
            return {k: (v.clone().requires_grad_(True) if not (k.endswith("weight_u") or k.endswith("weight_v")) else v.clone())
                    for k, v in sd.items()}
        self.g_sd, self.d_sd = leafify(g_sd), leafify(d_sd)
        self.vgg = synth.vgg_state_dict("vgg19", seed=3)
        self.vggface = synth.vgg_state_dict("vgg16", seed=5)
        self.pose_encoder = self.identity_encoder = None
        if with_embedder:
            torch.manual_seed(seed)
            self.pose_encoder = torchvision.models.mobilenet_v2(num_classes=cfg["pose_embedding_size"])
            if not finetune:
                self.identity_encoder = torchvision.models.resnext50_32x4d(num_classes=cfg["embed_channels"])
        g_params = [v for v in self.g_sd.values() if v.requires_grad]
        if not finetune and with_embedder:
            g_params += list(self.identity_encoder.parameters()) + list(self.pose_encoder.parameters())
        d_params = [v for v in self.d_sd.values() if v.requires_grad]
        if optimizer == "RAdam":
            self.opt_G = RAdamPort(g_params, lr_gen, (beta1, 0.999), 1e-5)
            self.opt_D = RAdamPort(d_params, lr_dis, (beta1, 0.999), 1e-5)
        else:
            self.opt_G = torch.optim.Adam(g_params, lr=lr_gen, betas=(beta1, 0.999), eps=1e-5)
            self.opt_D = torch.optim.Adam(d_params, lr=lr_dis, betas=(beta1, 0.999), eps=1e-5)
        self.ema = {k: v.detach().clone() for k, v in self.g_sd.items()}
        self.alpha = 0.972 if finetune else 0.999

    def embed(self, data, emb=None):
        if emb is not None:
            return emb["embeds"], emb["pose_embedding"], emb.get("embeds_elemwise")
        pose = self.pose_encoder(data["pose_input_rgbs"][:, 0])
        if self.finetune:
            ident = self.g_sd["identity_embedding"].expand(len(pose), -1)
            return ident, pose, None
        b, k, c, h, w = data["enc_rgbs"].shape
        per = self.identity_encoder(data["enc_rgbs"].view(-1, c, h, w)).view(b, k, -1)
        return per.mean(1), pose, per

    def step(self, data, target, emb=None):
        ident, pose, elem = self.embed(data, emb)
        label = torch.zeros_like(target["label"]) if self.finetune else target["label"]
        out, lg, ld = R.forward_losses(self.g_sd, self.d_sd, self.vgg, self.vggface, self.cfg, ident, pose,
                                       data["target_rgbs"][:, 0], target["real_segm"][:, 0], label, training=True,
                                       embeds_elemwise=elem, criteria=self.criteria)
        loss_G, loss_D = sum(lg.values()), sum(ld.values())
        self.opt_G.zero_grad()
        loss_G.backward(retain_graph=True)
        self.opt_G.step()
        self.opt_D.zero_grad()
        loss_D.backward()
        self.opt_D.step()
        with torch.no_grad():
            for k, v in self.g_sd.items():
                if v.requires_grad:
                    self.ema[k].mul_(self.alpha).add_(v.detach(), alpha=1 - self.alpha)
                else:
                    self.ema[k].copy_(v)
        return out, lg, ld
