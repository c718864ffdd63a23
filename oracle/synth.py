"""ORACLE — TEST INFRASTRUCTURE ONLY.  Deterministic synthetic weights and inputs shared by the golden-vector
generator (which loads them into the *reference* modules), the oracle restatement and the CUDA parity tests.

Nothing is stored for weights: every state_dict is re-created from a seed with CPU torch.Generator streams (bit
stable for a given torch build; the GPU box runs the same image), in the reference's key layout
(weight_orig / weight_u / weight_v, torchvision `features.N` indices for VGG).
Spectral-norm vectors u, v are brought to convergence with explicit power iterations, because a freshly
initialised reference net (random u, v) is numerically wild and disagrees with itself.
"""
import math

import torch

from .reference_model import (VGG16_CONVS, VGG19_CONVS, discriminator_layout, generator_layout)

SMALL_CFG = dict(  # every conv stays a multiple of 32 channels so the tensor-core path is exercised
    num_channels=32, max_num_channels=64, embed_channels=64, pose_embedding_size=32, image_size=32,
    dis_num_blocks=7, num_labels=5,
    perc_weight=3e-2, idt_embed_weight=6e-3, fm_weight=10.0, dice_weight=1.0, dis_embed_weight=1e-2, gan_type="gan")
FULL_CFG = dict(  # configs/default.yaml + train.py defaults
    num_channels=64, max_num_channels=512, embed_channels=512, pose_embedding_size=256, image_size=256,
    dis_num_blocks=7, num_labels=16,
    perc_weight=3e-2, idt_embed_weight=6e-3, fm_weight=10.0, dice_weight=1.0, dis_embed_weight=1e-2, gan_type="gan")


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def _sn_entries(sd, prefix, weight, g, bias=None, iters=30, eps=1e-4):
    """weight_orig (+bias) + converged weight_u / weight_v (power iterations, torch SpectralNorm semantics)."""
    wm = weight.reshape(weight.shape[0], -1).double()
    u = torch.randn(wm.shape[0], generator=g, dtype=torch.float64)
    u = u / u.norm().clamp_min(eps)
    v = torch.mv(wm.t(), u)
    v = v / v.norm().clamp_min(eps)
    for _ in range(iters):
        v = torch.mv(wm.t(), u)
        v = v / v.norm().clamp_min(eps)
        u = torch.mv(wm, v)
        u = u / u.norm().clamp_min(eps)
    if bias is not None:
        sd[prefix + ".bias"] = bias
    sd[prefix + ".weight_orig"] = weight
    sd[prefix + ".weight_u"] = u.float()
    sd[prefix + ".weight_v"] = v.float()


def generator_state_dict(cfg, seed=1, gamma_bias=1.0, finetuned=False):
    g = _gen(seed)
    blocks, c_last = generator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["image_size"])
    sd = {}
    sd["constant.constant"] = 1.0 + _randn(g, 1, blocks[0][0], 4, 4, std=0.5)
    affine_sizes = []
    for i, (cin, cout, up) in enumerate(blocks):
        p = f"decoder_blocks.{i}"
        i0, i1 = (4, 8) if up else (3, 7)
        _sn_entries(sd, f"{p}.block.{i0}", _randn(g, cout, cin, 3, 3, std=1 / math.sqrt(9 * cin)), g)
        _sn_entries(sd, f"{p}.block.{i1}", _randn(g, cout, cout, 3, 3, std=1 / math.sqrt(9 * cout)), g)
        if cin != cout or up:
            _sn_entries(sd, f"{p}.skip.1", _randn(g, cout, cin, 1, 1, std=1 / math.sqrt(cin)), g,
                        bias=_randn(g, cout, std=0.1))
        affine_sizes += [cin, cout]
    affine_sizes.append(c_last)
    nb = len(blocks)
    _sn_entries(sd, f"decoder_blocks.{nb + 2}", _randn(g, 4, c_last, 3, 3, std=1 / math.sqrt(9 * c_last)), g,
                bias=_randn(g, 4, std=0.1))
    joint = cfg["embed_channels"] + cfg["pose_embedding_size"]
    hidden = max(joint, 512)
    n_aff = 2 * sum(affine_sizes)
    _sn_entries(sd, "affine_params_projector.0", _randn(g, hidden, joint, std=1 / math.sqrt(joint)), g,
                bias=_randn(g, hidden, std=0.1))
    b2 = _randn(g, n_aff, std=0.2)
    off = 0
    for c in affine_sizes:          # layout [beta(C) | gamma(C)] per AdaIN: make gains O(1) like a trained net
        b2[off + c:off + 2 * c] += gamma_bias
        off += 2 * c
    _sn_entries(sd, "affine_params_projector.2", _randn(g, n_aff, hidden, std=1 / math.sqrt(hidden)), g, bias=b2)
    if finetuned:
        sd["identity_embedding"] = _randn(g, 1, cfg["embed_channels"], std=1.0)
    return sd


def discriminator_state_dict(cfg, seed=2, finetuned=False):
    g = _gen(seed)
    nc = cfg["num_channels"]
    sd = {}
    _sn_entries(sd, "down_block.0", _randn(g, nc, 3, 3, 3, std=1 / math.sqrt(27)), g, bias=_randn(g, nc, std=0.1))
    _sn_entries(sd, "down_block.2", _randn(g, nc, nc, 3, 3, std=1 / math.sqrt(9 * nc)), g, bias=_randn(g, nc, std=0.1))
    _sn_entries(sd, "skip.0", _randn(g, nc, 3, 1, 1, std=1 / math.sqrt(3)), g, bias=_randn(g, nc, std=0.1))
    layout = discriminator_layout(nc, cfg["max_num_channels"], cfg["embed_channels"], cfg["dis_num_blocks"],
                                  cfg["image_size"])
    for i, (cin, cout, down) in enumerate(layout):
        p = f"blocks.{i}"
        _sn_entries(sd, f"{p}.block.2", _randn(g, cout, cin, 3, 3, std=1 / math.sqrt(9 * cin)), g,
                    bias=_randn(g, cout, std=0.1))
        _sn_entries(sd, f"{p}.block.5", _randn(g, cout, cout, 3, 3, std=1 / math.sqrt(9 * cout)), g,
                    bias=_randn(g, cout, std=0.1))
        if cin != cout or down:
            _sn_entries(sd, f"{p}.skip.0", _randn(g, cout, cin, 1, 1, std=1 / math.sqrt(cin)), g,
                        bias=_randn(g, cout, std=0.1))
    e = cfg["embed_channels"]
    _sn_entries(sd, "linear", _randn(g, 1, e, std=1 / math.sqrt(e)), g, bias=_randn(g, 1, std=0.1))
    n_labels = 1 if finetuned else cfg["num_labels"]
    emb = (torch.rand(n_labels, e, generator=g) * 0.2 - 0.1)
    _sn_entries(sd, "embed", emb, g, eps=1e-12 if finetuned else 1e-4)
    return sd


def vgg_state_dict(kind="vgg19", seed=3):
    """Keys as in `torchvision.models.vgg{19,16}().features.state_dict()` restricted to the first 30 layers.
    He-style scaling keeps activations O(1)-O(100) through 13 conv+ReLU layers on caffe-normalised input."""
    g = _gen(seed)
    convs = VGG19_CONVS if kind == "vgg19" else VGG16_CONVS
    plan19 = [64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512]
    plan16 = [64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512]
    plan = plan19 if kind == "vgg19" else plan16
    sd = {}
    cin = 3
    for idx, cout in zip(convs, plan):
        sd[f"{idx}.weight"] = _randn(g, cout, cin, 3, 3, std=math.sqrt(2.0 / (9 * cin)))
        sd[f"{idx}.bias"] = _randn(g, cout, std=0.05)
        cin = cout
    if kind == "vgg19":
        sd["0.weight"] = sd["0.weight"] / 64.0   # caffe input range is +-128
    else:
        sd["0.weight"] = sd["0.weight"] / 64.0
    return sd


def make_inputs(cfg, batch, seed=4, n_identity_frames=1):
    """Synthetic batch with the dataloader's output contract."""
    g = _gen(seed)
    s = cfg["image_size"]
    yy, xx = torch.meshgrid(torch.arange(s, dtype=torch.float32), torch.arange(s, dtype=torch.float32), indexing="ij")
    mask = (((yy - (s - 1) / 2) ** 2 + (xx - (s - 1) / 2) ** 2) <= (0.4 * s) ** 2).float()
    img = torch.rand(batch, 1, 3, s, s, generator=g)
    data = dict(
        enc_rgbs=torch.rand(batch, n_identity_frames, 3, s, s, generator=g),
        pose_input_rgbs=torch.rand(batch, 1, 3, s, s, generator=g),
        target_rgbs=img * mask[None, None, None],
    )
    target = dict(
        real_segm=mask[None, None, None].expand(batch, 1, 3, s, s).contiguous(),
        label=torch.randint(0, cfg["num_labels"], (batch,), generator=g),
    )
    emb = dict(  # stand-ins for the embedder outputs (the embedders themselves are stock torchvision nets)
        embeds=torch.randn(batch, cfg["embed_channels"], generator=g),
        pose_embedding=torch.randn(batch, cfg["pose_embedding_size"], generator=g),
        embeds_elemwise=torch.randn(batch, n_identity_frames, cfg["embed_channels"], generator=g),
    )
    return data, target, emb


def pose_encoder_state_dict(num_classes=32, seed=7):
    """Deterministic weights for the pose encoder (torchvision MobileNetV2, the reference's
    embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:28): torchvision's own key / shape layout, values
    from a seeded generator (He-scaled convs, non-trivial BatchNorm affine parameters AND running statistics, so that
    eval mode and the running-statistics update are both exercised)."""
    import torchvision
    g = _gen(seed)
    layout = torchvision.models.mobilenet_v2(num_classes=num_classes).state_dict()
    sd = {}
    for k, v in layout.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            sd[k] = _randn(g, *v.shape, std=0.3)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(v.shape, generator=g) * 1.5 + 0.5
        elif v.dim() == 4:                                  # conv: He scaling on the fan-in
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            sd[k] = _randn(g, *v.shape, std=(2.0 / fan_in) ** 0.5)
        elif v.dim() == 2:                                  # classifier
            sd[k] = _randn(g, *v.shape, std=(1.0 / v.shape[1]) ** 0.5)
        elif k.endswith("bias") and "classifier" in k:
            sd[k] = _randn(g, *v.shape, std=0.1)
        elif k.endswith("weight"):                          # BatchNorm gamma
            sd[k] = torch.rand(v.shape, generator=g) + 0.5
        else:                                               # BatchNorm beta
            sd[k] = _randn(g, *v.shape, std=0.2)
    return sd


def pose_inputs(batch=3, image_size=64, seed=8):
    g = _gen(seed)
    return torch.rand((batch, 1, 3, image_size, image_size), generator=g)


def identity_encoder_state_dict(num_classes=512, seed=9):
    """Deterministic weights for the identity encoder (torchvision ResNeXt50-32x4d, the reference's
    embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:27): torchvision's key / shape layout, values from
    a seeded generator (He-scaled convs, non-trivial BatchNorm affine parameters and running statistics)."""
    import torchvision
    g = _gen(seed)
    layout = torchvision.models.resnext50_32x4d(num_classes=num_classes).state_dict()
    sd = {}
    for k, v in layout.items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            sd[k] = _randn(g, *v.shape, std=0.3)
        elif k.endswith("running_var"):
            sd[k] = torch.rand(v.shape, generator=g) * 1.5 + 0.5
        elif v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            sd[k] = _randn(g, *v.shape, std=(2.0 / fan_in) ** 0.5)
        elif v.dim() == 2:
            sd[k] = _randn(g, *v.shape, std=(1.0 / v.shape[1]) ** 0.5)
        elif k.startswith("fc.") and k.endswith("bias"):
            sd[k] = _randn(g, *v.shape, std=0.1)
        elif k.endswith("weight"):                          # BatchNorm gamma (block-final ones included: O(1))
            sd[k] = torch.rand(v.shape, generator=g) * 0.5 + 0.5
        else:
            sd[k] = _randn(g, *v.shape, std=0.2)
    return sd


def identity_inputs(batch=2, frames=4, image_size=128, seed=10):
    g = _gen(seed)
    return torch.rand((batch, frames, 3, image_size, image_size), generator=g)
