"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A CPU restatement (plain torch ops, any dtype: run it in float64 for ground truth) of the reference's hot path,
written from the reference's algorithm, each function citing the reference file:line it follows.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may import this package.

Parity pinning: the reference ships no tests or golden vectors.  This restatement is pinned
against the *reference modules themselves*, imported from /root/reference in the build container:
  * tests/test_oracle_vs_reference.py runs both on identical state_dicts and inputs (skipped where the reference
    tree is absent, i.e. on the GPU box);
  * oracle/make_golden.py dumps the reference's outputs into tests/golden/*.pt, and tests/test_oracle_golden.py
    checks this file against them everywhere.

State dicts use the reference's own key names (weight_orig / weight_u / weight_v for spectral-normalised layers).
"""
import math

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------------------------
# spectral norm — torch.nn.utils.spectral_norm (legacy hook), SpectralNorm.compute_weight; call sites
# generators/common/blocks.py:78-80,86-88,98-100; generators/vector_pose_..._noBottleneck.py:84-86,98,100;
# discriminators/no_landmarks.py:54-66,81,86.  n_power_iterations = 1.
# --------------------------------------------------------------------------------------------------------------
def _l2normalize(v, eps):
    return v / torch.clamp(v.norm(), min=eps)


def spectral_norm_weight(sd, prefix, training, eps=1e-4, update=True):
    """Returns (W / sigma, sigma).  In training mode performs one power iteration and (if update) writes the new
    u, v back into `sd` in place, exactly like the reference's forward pre-hook does with its buffers."""
    w = sd[prefix + ".weight_orig"]
    u = sd[prefix + ".weight_u"]
    v = sd[prefix + ".weight_v"]
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v_new = _l2normalize(torch.mv(wm.t(), u), eps)
            u_new = _l2normalize(torch.mv(wm, v_new), eps)
        if update:
            sd[prefix + ".weight_u"] = u_new.clone()
            sd[prefix + ".weight_v"] = v_new.clone()
        u, v = u_new, v_new
    sigma = torch.dot(u.detach(), torch.mv(wm, v.detach()))
    return w / sigma, sigma


# --------------------------------------------------------------------------------------------------------------
# generator — generators/vector_pose_unsupervised_segmentation_noBottleneck.py, generators/common/blocks.py
# --------------------------------------------------------------------------------------------------------------
def generator_layout(num_channels=64, max_num_channels=512, image_size=256, const_size=4, num_res_blocks=2):
    """Channel plan of the decoder (generator :60-78): list of (cin, cout, upsample) per ResBlock + final width."""
    n_up = int(math.log2(image_size / const_size))
    nonclamped = num_channels * (2 ** n_up)
    c = min(nonclamped, max_num_channels)
    blocks = [(c, c, False) for _ in range(num_res_blocks)]
    for _ in range(n_up):
        cin = c
        nonclamped //= 2
        c = min(nonclamped, max_num_channels)
        blocks.append((cin, c, True))
    return blocks, c


def adain(x, gamma, beta, eps=1e-4):
    """AdaptiveNorm2d.forward (blocks.py:18-26): InstanceNorm2d(eps, affine=False) then per-(n,c) scale and shift."""
    out = F.instance_norm(x, eps=eps)
    return out * gamma[:, :, None, None] + beta[:, :, None, None]


def generator_forward(sd, identity_embedding, pose_embedding, layout, training, update_sn=True):
    """Generator.forward (:165-181).  identity_embedding: (B, E) (data_dict['embeds'], or the expanded
    `identity_embedding` parameter in fine-tuning mode, :128-129).  Returns fake_rgbs, fake_segm, affine_params."""
    blocks, c_last = layout
    joint = torch.cat((identity_embedding, pose_embedding), dim=1)                       # :133
    w0, _ = spectral_norm_weight(sd, "affine_params_projector.0", training, eps=1e-12, update=update_sn)   # torch default eps
    h = F.relu(F.linear(joint, w0, sd["affine_params_projector.0.bias"]))                # :98-99
    w2, _ = spectral_norm_weight(sd, "affine_params_projector.2", training, eps=1e-12, update=update_sn)
    affine = F.linear(h, w2, sd["affine_params_projector.2.bias"])                       # :100

    off = 0

    def take(c):                                                                         # :108-125: [beta | gamma]
        nonlocal off
        beta = affine[:, off:off + c]
        gamma = affine[:, off + c:off + 2 * c]
        off += 2 * c
        return gamma, beta

    b = pose_embedding.shape[0]
    x = sd["constant.constant"].expand(b, -1, -1, -1)                                    # :31-37
    for i, (cin, cout, up) in enumerate(blocks):                                         # blocks.py:70-111
        p = f"decoder_blocks.{i}"
        i0, i1 = (4, 8) if up else (3, 7)
        g0, b0 = take(cin)
        g1, b1 = take(cout)
        h = F.relu(adain(x, g0, b0))
        if up:
            h = F.interpolate(h, scale_factor=2, mode="nearest")
        w, _ = spectral_norm_weight(sd, f"{p}.block.{i0}", training, update=update_sn)
        h = F.conv2d(h, w, None, padding=1)
        h = F.relu(adain(h, g1, b1))
        w, _ = spectral_norm_weight(sd, f"{p}.block.{i1}", training, update=update_sn)
        h = F.conv2d(h, w, None, padding=1)
        if cin != cout or up:
            s = x
            if up:
                s = F.interpolate(s, scale_factor=2, mode="nearest")
            w, _ = spectral_norm_weight(sd, f"{p}.skip.1", training, update=update_sn)
            s = F.conv2d(s, w, sd[f"{p}.skip.1.bias"])
        else:
            s = x
        x = h + s
    nb = len(blocks)
    g, bt = take(c_last)
    h = F.relu(adain(x, g, bt))                                                          # :81-82
    w, _ = spectral_norm_weight(sd, f"decoder_blocks.{nb + 2}", training, update=update_sn)
    t = torch.tanh(F.conv2d(h, w, sd[f"decoder_blocks.{nb + 2}.bias"], padding=1))       # :84-87
    rgb = t[:, :-1] * 0.75 + 0.5                                                         # :170-174
    segm = t[:, -1:] * 0.5 + 0.5                                                         # :177-178
    return rgb * segm, segm, affine                                                      # :180-181


# --------------------------------------------------------------------------------------------------------------
# discriminator — discriminators/no_landmarks.py
# --------------------------------------------------------------------------------------------------------------
def discriminator_layout(num_channels=64, max_num_channels=512, embed_channels=512, dis_num_blocks=7, image_size=256):
    """(cin, cout, downsample) for self.blocks (:68-79)."""
    num_down = min(int(math.log(image_size, 2)) - 2, dis_num_blocks)
    cin = num_channels
    cout = cin
    blocks = []
    for i in range(1, num_down):
        cout = min(cin * 2, max_num_channels)
        if i == dis_num_blocks - 1:
            cout = embed_channels
        blocks.append((cin, cout, True))
        cin = cout
    for i in range(num_down, dis_num_blocks):
        if i == dis_num_blocks - 1:
            cout = embed_channels
        blocks.append((cin, cout, False))
    return blocks


def discriminator_pass(sd, x, embed, layout, training, update_sn=True):
    """Discriminator.pass_inputs (:90-108), including the in-place-ReLU aliasing of ResBlock(norm='none')
    (blocks.py:73: the block's first ReLU(inplace) overwrites the block input, so the skip branch and the feature
    previously appended to `feats` both become post-ReLU)."""
    w, _ = spectral_norm_weight(sd, "down_block.0", training, update=update_sn)
    h = F.relu(F.conv2d(x, w, sd["down_block.0.bias"], padding=1))
    w, _ = spectral_norm_weight(sd, "down_block.2", training, update=update_sn)
    h = F.avg_pool2d(F.conv2d(h, w, sd["down_block.2.bias"], padding=1), 2)
    w, _ = spectral_norm_weight(sd, "skip.0", training, update=update_sn)
    out = h + F.avg_pool2d(F.conv2d(x, w, sd["skip.0.bias"]), 2)
    feats = []
    for i, (cin, cout, down) in enumerate(layout):
        p = f"blocks.{i}"
        r = F.relu(out)            # in place in the reference: `out` (already in feats) becomes r
        feats.append(r)
        w, _ = spectral_norm_weight(sd, f"{p}.block.2", training, update=update_sn)
        h = F.relu(F.conv2d(r, w, sd[f"{p}.block.2.bias"], padding=1))
        w, _ = spectral_norm_weight(sd, f"{p}.block.5", training, update=update_sn)
        h = F.conv2d(h, w, sd[f"{p}.block.5.bias"], padding=1)
        if down:
            h = F.avg_pool2d(h, 2)
        if cin != cout or down:
            w, _ = spectral_norm_weight(sd, f"{p}.skip.0", training, update=update_sn)
            s = F.conv2d(r, w, sd[f"{p}.skip.0.bias"])
            if down:
                s = F.avg_pool2d(s, 2)
        else:
            s = r
        out = h + s
    feats.append(out)              # last feature stays pre-ReLU (torch.relu below is out of place, :100)
    o = F.relu(out)
    o = o.view(o.shape[0], o.shape[1], -1).sum(2)
    w, _ = spectral_norm_weight(sd, "linear", training, update=update_sn)
    out_linear = F.linear(o, w, sd["linear.bias"])[:, 0]
    score = (o * embed).sum(1) + out_linear if embed is not None else out_linear
    return score, feats


def discriminator_forward(sd, fake_rgbs, target_rgbs, label, layout, training, embed_eps=1e-4, update_sn=True):
    """Discriminator.forward (:138-166): three passes sharing weights; sigma differs per pass in training mode
    because every call runs one more power iteration."""
    w, _ = spectral_norm_weight(sd, "embed", training, eps=embed_eps, update=update_sn)
    embed = F.embedding(label, w)
    fake_score_G, fake_features = discriminator_pass(sd, fake_rgbs, embed, layout, training, update_sn)
    fake_score_D, _ = discriminator_pass(sd, fake_rgbs.detach(), embed.detach(), layout, training, update_sn)
    real_score, real_features = discriminator_pass(sd, target_rgbs, embed, layout, training, update_sn)
    return dict(fake_features=fake_features, real_features=real_features, real_embedding=embed,
                fake_score_G=fake_score_G, fake_score_D=fake_score_D, real_score=real_score)


# --------------------------------------------------------------------------------------------------------------
# criterions
# --------------------------------------------------------------------------------------------------------------
VGG19_CONVS = (0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28)      # torchvision vgg19().features[:30]
VGG19_POOLS = (4, 9, 18, 27)
VGG16_CONVS = (0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28)      # torchvision vgg16().features[:30]
VGG16_POOLS = (4, 9, 16, 23)
CAFFE_MEAN = (103.939 / 255., 116.779 / 255., 123.680 / 255.)
CAFFE_STD = (1. / 255., 1. / 255., 1. / 255.)


def perceptual_loss(vgg_sd, fake, real, weight, convs=VGG19_CONVS, pools=VGG19_POOLS, num_layers=30):
    """PerceptualLoss.forward (criterions/common/perceptual_loss.py:91-110): `(x+1)/2` quirk, caffe normalisation,
    MaxPool->AvgPool (:76-77), L1 after each ReLU, target detached."""
    mean = torch.tensor(CAFFE_MEAN, dtype=fake.dtype, device=fake.device)[None, :, None, None]
    std = torch.tensor(CAFFE_STD, dtype=fake.dtype, device=fake.device)[None, :, None, None]
    a = ((fake + 1) / 2 - mean) / std
    b = ((real.detach() + 1) / 2 - mean) / std
    loss = 0
    for i in range(num_layers):
        if i in convs:
            w, bias = vgg_sd[f"{i}.weight"], vgg_sd[f"{i}.bias"]
            a = F.conv2d(a, w, bias, padding=1)
            b = F.conv2d(b, w, bias, padding=1)
        elif i in pools:
            a = F.avg_pool2d(a, 2)
            b = F.avg_pool2d(b, 2)
        else:
            a = F.relu(a)
            b = F.relu(b)
            loss = loss + F.l1_loss(a, b)
    return loss * weight


def crop_and_resize_center(images, crop_factor=1 / 1.8):
    """criterions/idt_embed.py:38-50,58-83: centre crop by affine_grid + grid_sample(bilinear, reflection)."""
    n, c, h, w = images.shape
    t = h * (1 - crop_factor) / 2
    l = w * (1 - crop_factor) / 2
    b = h - t
    r = w - l
    theta = torch.zeros(n, 2, 3, dtype=torch.float32, device=images.device)
    theta[:, 0, 0] = (r - l) / w
    theta[:, 1, 1] = (b - t) / h
    theta[:, 0, 2] = (l + r) / w - 1
    theta[:, 1, 2] = (t + b) / h - 1
    grid = F.affine_grid(theta.to(images.dtype), (n, c, h, w), align_corners=False)
    return F.grid_sample(images, grid, mode="bilinear", padding_mode="reflection", align_corners=False)


def idt_embed_loss(vggface_sd, fake, real, weight):
    """criterions/idt_embed.py:21-56 without `dec_keypoints`."""
    return perceptual_loss(vggface_sd, crop_and_resize_center(fake), crop_and_resize_center(real), weight,
                           convs=VGG16_CONVS, pools=VGG16_POOLS)


def featmat_loss(fake_feats, real_feats, fm_weight):
    """criterions/featmat.py:18-20."""
    return sum(F.l1_loss(a, b.detach()) for a, b in zip(fake_feats, real_feats)) / len(fake_feats) * fm_weight


def adversarial_losses(fake_score_G, fake_score_D, real_score, gan_type="gan"):
    """criterions/adversarial.py:20-57 -> (loss_G, loss_D)."""
    def preds(real, fake):
        if gan_type == "gan":
            return real, fake
        if gan_type == "rgan":
            return real - fake, fake - real
        if gan_type == "ragan":
            return real - fake.mean(), fake - real.mean()
        raise Exception("Incorrect `gan_type` argument")
    real_pred, fake_pred_D = preds(real_score, fake_score_D)
    _, fake_pred_G = preds(real_score, fake_score_G)
    loss_D = torch.relu(1. - real_pred).mean() + torch.relu(1. + fake_pred_D).mean()
    if gan_type == "gan":
        loss_G = -fake_pred_G.mean()
    else:
        loss_G = torch.relu(1. + real_pred).mean() + torch.relu(1. - fake_pred_G).mean()
    return loss_G, loss_D


def dice_loss(fake_segm, real_segm, dice_weight):
    """criterions/dice.py:21-37 (real_segm (B,3,S,S) broadcasts against fake_segm (B,1,S,S))."""
    numer = (2 * fake_segm * real_segm).sum()
    denom = (fake_segm ** 2).sum() + (real_segm ** 2).sum()
    return -torch.log(numer / denom) * dice_weight


def dis_embed_loss(embeds_elemwise, real_embedding, weight):
    """criterions/dis_embed.py:19-33."""
    a = embeds_elemwise[:, 0] if embeds_elemwise.dim() > 2 else embeds_elemwise
    b = real_embedding[:, 0] if real_embedding.dim() > 2 else real_embedding
    return F.l1_loss(a, b.detach()) * weight


# --------------------------------------------------------------------------------------------------------------
# one generator+discriminator+criteria forward, as TrainingModule.forward (runners/holycow.py:153-201) wires it,
# with the embedder outputs given (the embedders are stock torchvision networks on both sides).
# --------------------------------------------------------------------------------------------------------------
def forward_losses(g_sd, d_sd, vgg_sd, vggface_sd, cfg, identity_embedding, pose_embedding, target_rgbs, real_segm,
                   label, training=True, embeds_elemwise=None, criteria=("perceptual", "adversarial", "featmat", "dice")):
    """Returns (outputs dict, losses_G dict, losses_D dict).  g_sd / d_sd may contain tensors requiring grad."""
    g_layout = generator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["image_size"])
    d_layout = discriminator_layout(cfg["num_channels"], cfg["max_num_channels"], cfg["embed_channels"],
                                    cfg.get("dis_num_blocks", 7), cfg["image_size"])
    fake_rgbs, fake_segm, affine = generator_forward(g_sd, identity_embedding, pose_embedding, g_layout, training)
    d = discriminator_forward(d_sd, fake_rgbs, target_rgbs, label, d_layout, training,
                              embed_eps=cfg.get("embed_eps", 1e-4))
    out = dict(fake_rgbs=fake_rgbs, fake_segm=fake_segm, affine_params=affine, **d)
    lg, ld = {}, {}
    for name in criteria:
        if name == "perceptual":
            lg["VGG"] = perceptual_loss(vgg_sd, fake_rgbs, target_rgbs, cfg["perc_weight"])
        elif name == "idt_embed":
            lg["VGGFace"] = idt_embed_loss(vggface_sd, fake_rgbs, target_rgbs, cfg["idt_embed_weight"])
        elif name == "adversarial":
            g, dd = adversarial_losses(d["fake_score_G"], d["fake_score_D"], d["real_score"], cfg.get("gan_type", "gan"))
            lg["adversarial_G"] = g
            ld["adversarial_D"] = dd
        elif name == "featmat":
            lg["feature_matching"] = featmat_loss(d["fake_features"], d["real_features"], cfg["fm_weight"])
        elif name == "dice":
            lg["segmentation_dice"] = dice_loss(fake_segm, real_segm, cfg["dice_weight"])
        elif name == "dis_embed":
            lg["embedding_matching"] = dis_embed_loss(embeds_elemwise, d["real_embedding"], cfg["dis_embed_weight"])
        else:
            raise ValueError(name)
    return out, lg, ld
