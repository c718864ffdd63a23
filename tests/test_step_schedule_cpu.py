"""The criteria and one full runner step on the CPU over the kernel emulations (tests/kernel_emulators.py), in float64,
against the golden vectors of the unmodified reference (tests/golden/small.pt): VGG19 / VGG-Face node, feature
matching, the 3-pass discriminator, both backward passes, the discarded-weight-gradient shortcut, Adam + EMA — and the
same step with gradients accumulated in place through ops.direct_grads (forced on: CPU buckets never offer sinks).
GPU counterpart with the real kernels: tests/test_parity_gpu.py."""
import importlib
import tempfile

import pytest
import torch

import kernel_emulators as E
from helpers import StubEmbedder, make_args, max_abs, write_vgg_files
from oracle import synth


@pytest.fixture()
def emu(monkeypatch):
    E.install(monkeypatch)
    return E


@pytest.fixture(scope="module")
def small():
    cfg = synth.SMALL_CFG
    data, target, emb = synth.make_inputs(cfg, batch=2, seed=4)
    return dict(cfg=cfg, g_sd=synth.generator_state_dict(cfg, seed=1), d_sd=synth.discriminator_state_dict(cfg, seed=2),
                data=data, target=target, emb=emb)


def _dbl(d):
    return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}


def _nets(small, names, **over):
    cfg = small["cfg"]
    with tempfile.TemporaryDirectory() as vgg_dir:
        write_vgg_files(vgg_dir)
        args = make_args(cfg, device="cpu", vgg_weights_dir=vgg_dir, **over)
        crit = [importlib.import_module(f"criterions.{n}").Wrapper.get_net(args).double() for n in names]
    G = importlib.import_module("generators.vector_pose_unsupervised_segmentation_noBottleneck").Wrapper.get_net(args)
    D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_net(args)
    G.load_state_dict(small["g_sd"]); D.load_state_dict(small["d_sd"])
    return args, crit, G.double(), D.double()


def test_criteria_schedule_values(emu, small, golden_small):
    names = ("perceptual", "idt_embed", "adversarial", "featmat", "dice", "dis_embed")
    _, crit, _, D = _nets(small, names)
    crit = dict(zip(names, crit))
    D.eval()
    with torch.no_grad():
        dd = _dbl(dict(fake_rgbs=golden_small["g_eval.fake_rgbs"], target_rgbs=small["data"]["target_rgbs"],
                       label=small["target"]["label"], fake_segm=golden_small["g_eval.fake_segm"],
                       real_segm=small["target"]["real_segm"], embeds_elemwise=small["emb"]["embeds_elemwise"]))
        D(dd)
        got = {"VGG": crit["perceptual"](dd)["VGG"], "VGGFace": crit["idt_embed"](dd)["VGGFace"]}
        lg, ld = crit["adversarial"](dd)
        got["adversarial_G"], got["adversarial_D"] = lg["adversarial_G"], ld["adversarial_D"]
        got["feature_matching"] = crit["featmat"](dd)["feature_matching"]
        got["segmentation_dice"] = crit["dice"](dd)["segmentation_dice"]
        got["embedding_matching"] = crit["dis_embed"](dd)["embedding_matching"]
    for k, v in got.items():
        ref = float(golden_small["crit." + k])
        assert abs(float(v) - ref) <= 5e-5 * abs(ref) + 1e-7, (k, float(v), ref)


def _check_step(r, golden_small):
    for k, v in {**r["lG"], **r["lD"]}.items():
        ref = float(golden_small["step.loss." + k])
        assert abs(float(v) - ref) <= 5e-5 * abs(ref) + 1e-7, (k, float(v), ref)
    g_floor = 1e-5 * max(golden_small["step.gradG.norms"].values())
    d_floor = 1e-5 * max(golden_small["step.gradD.norms"].values())
    for k, ref_norm in golden_small["step.gradG.norms"].items():
        assert abs(float(r["gradG"][k].norm()) - ref_norm) <= 2e-4 * ref_norm + g_floor, (k, float(r["gradG"][k].norm()), ref_norm)
    for k, ref_norm in golden_small["step.gradD.norms"].items():
        assert abs(float(r["gradD"][k].norm()) - ref_norm) <= 2e-4 * ref_norm + d_floor, (k, float(r["gradD"][k].norm()), ref_norm)
    for k, v in golden_small.items():
        if k.startswith("step.gradG.") and k != "step.gradG.norms":
            name = k[len("step.gradG."):]
            assert max_abs(r["gradG"][name], v) <= 2e-4 * float(v.abs().max()) + g_floor, name
        if k.startswith("step.gradD.") and k != "step.gradD.norms":
            name = k[len("step.gradD."):]
            assert max_abs(r["gradD"][name], v) <= 2e-4 * float(v.abs().max()) + d_floor, name
    g_after = dict(r["G"].named_parameters())["decoder_blocks.0.block.3.weight_orig"]
    assert max_abs(g_after, golden_small["step.after.G.decoder_blocks.0.block.3.weight_orig"]) < 2e-6
    ema = r["tm"].running_averages["generator"].state_dict()["decoder_blocks.0.block.3.weight_orig"]
    assert max_abs(ema, golden_small["step.after.ema.G.decoder_blocks.0.block.3.weight_orig"]) < 1e-6


def _module(small):
    runner = importlib.import_module("runners.holycow")
    args, crit, G, D = _nets(small, ("idt_embed", "perceptual", "adversarial", "featmat", "dis_embed", "dice"))
    Emb = StubEmbedder(_dbl(small["emb"])).double()
    tm = runner.TrainingModule(Emb, G, D, crit, [], {})
    tm.train()
    opt_G = runner.get_optimizer(Emb, G, args)
    opt_D = importlib.import_module("discriminators.no_landmarks").Wrapper.get_optimizer(D, args)
    return runner, tm, Emb, G, D, opt_G, opt_D


@pytest.mark.parametrize("skip_discarded", [True, False])
def test_step_schedule_by_hand(emu, small, golden_small, skip_discarded):
    """The step driven call by call (plain autograd accumulation into the buckets), with and without the shortcut that
    skips the discriminator weight gradients the reference computes and then zeroes (runners/holycow.py:247)."""
    runner, tm, Emb, G, D, opt_G, opt_D = _module(small)
    bucket_G, bucket_D = tm.grad_buckets(opt_G, opt_D)
    D.skip_discarded_wgrad = skip_discarded
    _, lG, lD = tm(_dbl(small["data"]), _dbl(small["target"]))
    loss_G, loss_D = sum(lG.values()), sum(lD.values())
    bucket_G.zero()
    if not skip_discarded:
        bucket_D.zero()
    loss_G.backward(retain_graph=True)
    gradG = {k: p.grad.detach().clone() for k, p in G.named_parameters()}
    ref = float(golden_small["step.gradE.scale"])
    assert abs(float(Emb.scale.grad) - ref) <= 2e-4 * abs(ref) + 1e-9
    opt_G.step()
    bucket_D.zero()
    loss_D.backward()
    gradD = {k: p.grad.detach().clone() for k, p in D.named_parameters()}
    opt_D.step()
    tm.update_running_average(0.999)
    _check_step(dict(lG=lG, lD=lD, gradG=gradG, gradD=gradD, G=G, D=D, tm=tm), golden_small)


def test_step_schedule_train_step_with_gradient_sinks(emu, small, golden_small, monkeypatch):
    """runners.holycow.train_step end to end with the in-place gradient path forced on."""
    runner, tm, Emb, G, D, opt_G, opt_D = _module(small)

    def sinks(self):
        self.attach()
        return {p.data_ptr(): p.grad for p in self.params}

    monkeypatch.setattr(runner.GradBucket, "sinks", sinks)
    calls = {"n": 0}
    real = E.conv_wgrad_sn_acc

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    from b200lp import kernels as K
    monkeypatch.setattr(K, "conv_wgrad_sn_acc", counting)
    _, lG, lD = runner.train_step(tm, _dbl(small["data"]), _dbl(small["target"]), opt_G, opt_D, finetune=False)
    assert calls["n"] > 20                       # generator convs once, discriminator convs in two passes
    gradG = {k: p.grad.detach().clone() for k, p in G.named_parameters()}
    gradD = {k: p.grad.detach().clone() for k, p in D.named_parameters()}
    _check_step(dict(lG=lG, lD=lD, gradG=gradG, gradD=gradD, G=G, D=D, tm=tm), golden_small)
