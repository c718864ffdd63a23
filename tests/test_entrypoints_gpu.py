"""The reference's three entry-point flows on the B200 path, end to end with the synthetic dataset plugin:
meta-training (train.py --config default) -> checkpoint -> fine-tuning (train.py --config finetuning-base
--checkpoint_path ...) -> checkpoint -> rendering (drive.py).  Small image size / channel widths, a handful of steps."""
import subprocess
import sys
import tempfile
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))
from helpers import write_vgg_files  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "latent-pose-reenactment_b200"

SMALL = ["--image_size", "32", "--num_channels", "32", "--max_num_channels", "64", "--embed_channels", "64",
         "--pose_embedding_size", "32", "--num_workers", "0", "--synthetic_num_samples", "8",
         "--synthetic_num_identities", "4", "--n_frames_for_encoder", "2", "--no-logging"]


def run(cmd, cwd):
    p = subprocess.run([sys.executable] + cmd, cwd=cwd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, f"{' '.join(cmd)}\n--- stdout\n{p.stdout[-3000:]}\n--- stderr\n{p.stderr[-3000:]}"
    return p.stdout


def test_train_finetune_drive_roundtrip():
    import torch
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        vgg = tmp / "vgg"
        vgg.mkdir()
        write_vgg_files(str(vgg))
        exp = tmp / "exp"
        common = SMALL + ["--vgg_weights_dir", str(vgg), "--experiments_dir", str(exp)]
        # meta-training: 2 epochs x 2 iterations (8 samples, batch 4), eager and graph-replayed steps both exercised
        out = run([str(PKG / "train.py"), "--config", "default", "--num_gpus", "1", "--batch_size", "4", "--num_epochs", "2",
                   "--experiment_name", "meta"] + common, cwd=str(PKG))
        assert "Entering training loop" in out and "Saving checkpoint" in out
        ckpts = sorted((exp / "meta" / "checkpoints").glob("model_*.pth"))
        assert ckpts, out[-2000:]
        ck = torch.load(ckpts[-1], map_location="cpu", weights_only=False)
        assert set(ck) == {"embedder", "generator", "discriminator", "optimizer_G", "optimizer_D", "running_averages", "args"}
        assert "decoder_blocks.0.block.3.weight_orig" in ck["generator"] and "embed.weight_u" in ck["discriminator"]
        assert all(torch.isfinite(v).all() for v in ck["generator"].values())
        # fine-tuning from that checkpoint (identity embedding initialised from the EMA embedder, RAdam, 1-row embedding)
        out = run([str(PKG / "train.py"), "--config", "finetuning-base", "--checkpoint_path", str(ckpts[-1]),
                   "--batch_size", "4", "--num_epochs", "2", "--experiment_name", "ft"] + common, cwd=str(PKG))
        assert "computing an averaged identity embedding" in out
        ft = sorted((exp / "ft" / "checkpoints").glob("model_*.pth"))
        assert ft, out[-2000:]
        ck = torch.load(ft[-1], map_location="cpu", weights_only=False)
        assert ck["generator"]["identity_embedding"].shape == (1, 64) and ck["discriminator"]["embed.weight_orig"].shape == (1, 64)
        assert ck["args"].finetune is True
        # rendering with the fine-tuned checkpoint: 6 synthetic driver frames, batches of 4 -> PNG frames
        dest = tmp / "render"
        out = run([str(PKG / "drive.py"), str(ft[-1]), "--destination", str(dest), "--synthetic_frames", "6",
                   "--batch_size", "4", "--format", "png"], cwd=str(PKG))
        frames = sorted((dest / "synthetic").glob("*.png"))
        assert len(frames) == 6, out[-2000:]
        import cv2
        img = cv2.imread(str(frames[0]))
        assert img.shape == (32, 64, 3)          # driver | rendered frame side by side, as the reference's drive.py
