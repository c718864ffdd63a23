"""Shared helpers for the parity tests (test infrastructure only)."""
import os
from argparse import Namespace

import torch

from oracle import synth


def make_args(cfg, device="cpu", **over):
    a = Namespace(gen_padding="zero", in_channels=3, out_channels=3, num_channels=cfg["num_channels"],
                  max_num_channels=cfg["max_num_channels"], embed_channels=cfg["embed_channels"],
                  pose_embedding_size=cfg["pose_embedding_size"], norm_layer="in", gen_constant_input_size=4,
                  gen_num_residual_blocks=2, image_size=cfg["image_size"], device=device, average_function="sum",
                  dis_padding="zero", dis_num_blocks=cfg["dis_num_blocks"], num_labels=cfg["num_labels"],
                  gan_type=cfg.get("gan_type", "gan"), fm_weight=cfg["fm_weight"], dice_weight=cfg["dice_weight"],
                  perc_weight=cfg["perc_weight"], idt_embed_weight=cfg["idt_embed_weight"],
                  dis_embed_weight=cfg["dis_embed_weight"], optimizer="Adam", lr_gen=5e-5, lr_dis=2e-4, beta1=0.0,
                  finetune=False, num_gpus=1)
    for k, v in over.items():
        setattr(a, k, v)
    return a


def write_vgg_files(dirname, vgg19_seed=3, vggface_seed=5):
    """Weight files in the layout criterions/common/perceptual_loss.py reads (features.* keys are all it uses)."""
    v19 = synth.vgg_state_dict("vgg19", seed=vgg19_seed)
    torch.save({"features." + k: v for k, v in v19.items()}, os.path.join(dirname, "vgg19-d01eb7cb.pth"))
    v16 = synth.vgg_state_dict("vgg16", seed=vggface_seed)
    torch.save(dict(v16), os.path.join(dirname, "vgg_face_weights.pth"))
    return v19, v16


class StubEmbedder(torch.nn.Module):
    """Same stand-in as oracle/make_golden.py: precomputed embeddings times a trainable scale."""

    def __init__(self, emb):
        super().__init__()
        self.emb = emb
        self.scale = torch.nn.Parameter(torch.ones(()))
        self.finetuning = False

    def forward(self, data_dict):
        data_dict["embeds"] = self.emb["embeds"] * self.scale
        data_dict["embeds_elemwise"] = self.emb["embeds_elemwise"] * self.scale
        data_dict["pose_embedding"] = self.emb["pose_embedding"] * self.scale


def to_dev(d, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


def max_abs(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def rel_err(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
