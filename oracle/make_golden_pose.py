"""ORACLE — TEST INFRASTRUCTURE ONLY.  tests/golden/pose.pt: outputs of the UNMODIFIED reference embedder's pose path
(/root/reference/embedders/unsupervised_pose_separate_embResNeXt_segmentation.py:56-58, `Embedder.get_pose_embedding`)
on the deterministic weights / inputs of oracle/synth.py (`pose_encoder_state_dict`, `pose_inputs`).

    python oracle/make_golden_pose.py          # needs /root/reference (build container only)

Stored: the pose embedding in eval mode and in train mode (batch statistics; the classifier's Dropout probability is set
to 0 on the instantiated module — RNG streams differ between devices), and every BatchNorm's running statistics after
the train-mode call.  Like oracle/make_golden.py this process must not import the product packages.
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import REPO, setup_reference_imports  # noqa: E402


def main():
    setup_reference_imports()
    import importlib
    from oracle import synth
    mod = importlib.import_module("embedders.unsupervised_pose_separate_embResNeXt_segmentation")
    assert "/reference/" in mod.__file__.replace("\\", "/"), mod.__file__
    num_classes = 32
    torch.manual_seed(0)
    emb = mod.Embedder(16, num_classes, "sum")
    emb.pose_encoder.load_state_dict(synth.pose_encoder_state_dict(num_classes, seed=7), strict=True)
    x = synth.pose_inputs(batch=3, image_size=128, seed=8)
    out = {"num_classes": num_classes}
    with torch.no_grad():
        emb.eval()
        d = {"pose_input_rgbs": x}
        emb.get_pose_embedding(d)
        out["eval.pose_embedding"] = d["pose_embedding"].clone()
        emb.train()
        emb.pose_encoder.classifier[0].p = 0.0
        d = {"pose_input_rgbs": x}
        emb.get_pose_embedding(d)
        out["train.pose_embedding"] = d["pose_embedding"].clone()
    bns = [m for m in emb.pose_encoder.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    out["train.running_mean"] = torch.cat([m.running_mean for m in bns]).clone()
    out["train.running_var"] = torch.cat([m.running_var for m in bns]).clone()
    out["train.num_batches_tracked"] = int(bns[0].num_batches_tracked)
    path = REPO / "tests" / "golden" / "pose.pt"
    torch.save(out, path)
    print(f"wrote {path}: eval |y| max {float(out['eval.pose_embedding'].abs().max()):.3f}, "
          f"train |y| max {float(out['train.pose_embedding'].abs().max()):.3f}, {len(bns)} BatchNorm layers")


if __name__ == "__main__":
    main()
