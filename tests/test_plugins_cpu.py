"""Host-side checks of the plugin mirror (no GPU): state_dict keys / shapes / parameter order identical to the
reference (recorded in tests/golden/small.pt by oracle/make_golden.py), deepcopy / load_state_dict round trips,
fine-tuning structure changes, and: the product path must fail loudly without a B200 (no CPU fallback)."""
import copy
import importlib
from argparse import Namespace

import pytest
import torch

from oracle import synth


def make_args(cfg, device="cpu", **over):
    a = Namespace(gen_padding="zero", in_channels=3, out_channels=3, num_channels=cfg["num_channels"],
                  max_num_channels=cfg["max_num_channels"], embed_channels=cfg["embed_channels"],
                  pose_embedding_size=cfg["pose_embedding_size"], norm_layer="in", gen_constant_input_size=4,
                  gen_num_residual_blocks=2, image_size=cfg["image_size"], device=device, average_function="sum",
                  dis_padding="zero", dis_num_blocks=cfg["dis_num_blocks"], num_labels=cfg["num_labels"],
                  gan_type="gan", fm_weight=10.0, dice_weight=1.0, perc_weight=3e-2, idt_embed_weight=6e-3,
                  dis_embed_weight=1e-2, optimizer="Adam", lr_gen=5e-5, lr_dis=2e-4, beta1=0.0, finetune=False,
                  num_gpus=1)
    for k, v in over.items():
        setattr(a, k, v)
    return a


@pytest.fixture(scope="module")
def nets():
    cfg = synth.SMALL_CFG
    args = make_args(cfg)
    G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
    D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
    return cfg, args, G, D


def test_state_dict_keys_shapes_and_param_order(nets, golden_small):
    cfg, args, G, D = nets
    assert list(G.state_dict().keys()) == golden_small["g_state_keys"]
    assert list(D.state_dict().keys()) == golden_small["d_state_keys"]
    assert {k: tuple(v.shape) for k, v in G.state_dict().items()} == golden_small["g_state_shapes"]
    assert {k: tuple(v.shape) for k, v in D.state_dict().items()} == golden_small["d_state_shapes"]
    # optimizer state in checkpoints is indexed by parameter order
    assert [k for k, _ in G.named_parameters()] == golden_small["g_param_order"]
    assert [k for k, _ in D.named_parameters()] == golden_small["d_param_order"]


def test_load_reference_layout_state_dict_and_deepcopy(nets):
    cfg, args, G, D = nets
    g_sd = synth.generator_state_dict(cfg, seed=1)
    d_sd = synth.discriminator_state_dict(cfg, seed=2)
    G.load_state_dict(g_sd, strict=True)
    D.load_state_dict(d_sd, strict=True)
    G2 = copy.deepcopy(G)
    for (k, a), (_, b) in zip(G.state_dict().items(), G2.state_dict().items()):
        assert torch.equal(a, b), k
    G2.eval().requires_grad_(False)
    assert all(not p.requires_grad for p in G2.parameters())
    assert G.get_num_affine_params() == g_sd["affine_params_projector.2.bias"].numel()


def test_finetuning_structure(nets, golden_small):
    cfg, args, _, _ = nets
    G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
    D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
    e = torch.randn(1, cfg["embed_channels"])
    G.enable_finetuning({"embeds": e.clone()})
    D.enable_finetuning({"embeds": e.clone()})
    assert list(G.state_dict().keys()) == golden_small["ft.g_state_keys"]
    assert {k: tuple(v.shape) for k, v in D.state_dict().items() if k.startswith("embed")} == golden_small["ft.d_state_shapes"]
    assert [k for k, _ in G.named_parameters()][0] == "identity_embedding"   # own parameters precede children
    assert D.embed.eps == 1e-12 and torch.equal(D.embed.weight_orig.data, e)
    G.enable_finetuning({"embeds": e * 2})          # second call copies in place
    assert torch.equal(G.identity_embedding.data, e * 2)
    # fine-tune optimizer covers generator parameters only (runners/holycow.py:34-41)
    runner = importlib.import_module("runners.holycow")
    E = torch.nn.Linear(2, 2)
    opt = runner.get_optimizer(E, G, make_args(cfg, finetune=True))
    n = sum(len(g["params"]) for g in opt.param_groups)
    assert n == len(list(G.parameters()))
    opt = runner.get_optimizer(E, G, make_args(cfg, finetune=False))
    assert sum(len(g["params"]) for g in opt.param_groups) == len(list(G.parameters())) + 2


def test_spectral_norm_power_iteration_matches_oracle(nets):
    from oracle import reference_model as R
    cfg, args, G, D = nets
    g_sd = synth.generator_state_dict(cfg, seed=1)
    G.load_state_dict(g_sd, strict=True)
    conv = G.decoder_blocks.slot(0).block.slot(3)
    G.train()
    inv = conv.inv_sigma()
    sd = {k: v.clone() for k, v in g_sd.items()}
    _, sigma = R.spectral_norm_weight(sd, "decoder_blocks.0.block.3", training=True)
    torch.testing.assert_close(1.0 / inv[0], sigma, rtol=1e-6, atol=0)
    torch.testing.assert_close(conv.weight_u, sd["decoder_blocks.0.block.3.weight_u"], rtol=1e-6, atol=1e-7)
    G.eval()
    u_before = conv.weight_u.clone()
    conv.inv_sigma()
    assert torch.equal(conv.weight_u, u_before)     # eval: no power iteration
    # d(1/sigma)/dW is the rank-1 term of SURVEY Appendix D
    inv = conv.inv_sigma()
    (g,) = torch.autograd.grad(inv.sum(), conv.weight_orig)
    sigma = 1.0 / inv.detach()
    expect = -(torch.outer(conv.weight_u, conv.weight_v).view_as(g)) / sigma ** 2
    torch.testing.assert_close(g, expect, rtol=1e-5, atol=1e-8)


def test_no_cpu_fallback(nets):
    """Without a CUDA device the hot path raises; it never silently computes on the host."""
    cfg, args, G, D = nets
    from b200lp.lib import B200lpError
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    _, _, emb = synth.make_inputs(cfg, batch=2, seed=4)
    with pytest.raises(B200lpError):
        G({"embeds": emb["embeds"], "pose_embedding": emb["pose_embedding"]})
    with pytest.raises(B200lpError):
        D({"fake_rgbs": torch.rand(2, 3, 32, 32), "target_rgbs": torch.rand(2, 3, 32, 32),
           "label": torch.zeros(2, dtype=torch.long)})
