"""Entry point: render "puppeteering" frames with a fine-tuned model — the reference's `drive.py` (:43-98):
load checkpoint (EMA weights) -> per driver batch: pose embedder -> generator -> clamp -> uint8 -> writer.

    python drive.py <checkpoint.pth> --destination out/ [--batch_size 64] [--synthetic_frames 10000]

Differences from the reference: `--batch_size` (the reference hard-codes 1 at :57; BASELINE config 4 drives batches
of 64), driver frames come from a folder of images (`--images_paths`) or from the synthetic generator, frames are
written as an MJPG .avi or a folder of PNGs (no face_alignment / ffmpeg dependency), and the device->host copy of
finished frames is pinned + asynchronous (double-buffered) instead of a per-frame `.cpu()` sync.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import argparse
import copy
import logging
from pathlib import Path

import torch

torch.set_grad_enabled(False)

from utils import utils


class FrameWriter:
    """Folder-of-PNGs or MJPG .avi sink for (H, W, 3) uint8 RGB frames."""

    def __init__(self, path):
        self.path = Path(path)
        self.video = None
        self.count = 0
        if self.path.suffix in ('.avi', '.mp4'):
            self.path = self.path.with_suffix('.avi')
        else:
            self.path.mkdir(parents=True, exist_ok=True)

    def add(self, frame_rgb):
        import cv2
        bgr = cv2.cvtColor(frame_rgb, cv2.COLOR_RGB2BGR)
        if self.path.suffix == '.avi':
            if self.video is None:
                h, w = bgr.shape[:2]
                self.video = cv2.VideoWriter(str(self.path), cv2.VideoWriter_fourcc(*'MJPG'), 25.0, (w, h))
            self.video.write(bgr)
        else:
            cv2.imwrite(str(self.path / f'{self.count:06}.png'), bgr)
        self.count += 1

    def close(self):
        if self.video is not None:
            self.video.release()


def load_driver_frames(folder, image_size):
    import cv2
    files = sorted(p for p in Path(folder).iterdir() if p.suffix.lower() in ('.jpg', '.jpeg', '.png'))
    for f in files:
        img = cv2.cvtColor(cv2.imread(str(f)), cv2.COLOR_BGR2RGB)
        if img.shape[0] != image_size or img.shape[1] != image_size:
            img = cv2.resize(img, (image_size, image_size), interpolation=cv2.INTER_AREA)
        yield torch.from_numpy(img).permute(2, 0, 1).float().div_(255.)


def batches(frame_iter, batch_size):
    buf = []
    for f in frame_iter:
        buf.append(f)
        if len(buf) == batch_size:
            yield torch.stack(buf)
            buf = []
    if buf:
        yield torch.stack(buf)


def render(embedder, generator, frames_batches, device, sink=None, with_driver=True):
    """Generator over rendered batches; D2H of batch i overlaps the kernels of batch i+1."""
    pending = None
    n_frames = 0
    for batch in frames_batches:
        pinned = batch.pin_memory() if device.startswith('cuda') else batch
        data_dict = {'pose_input_rgbs': pinned.to(device, non_blocking=True)[:, None]}
        embedder.get_pose_embedding(data_dict)
        generator(data_dict)
        out = data_dict['fake_rgbs'].clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1)
        if with_driver:
            drv = data_dict['pose_input_rgbs'][:, 0].clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1)
            out = torch.cat((drv, out), dim=2)
        host = torch.empty(out.shape, dtype=torch.uint8, pin_memory=device.startswith('cuda'))
        host.copy_(out, non_blocking=True)
        ev = torch.cuda.Event() if device.startswith('cuda') else None
        if ev is not None:
            ev.record()
        if pending is not None:
            n_frames += _flush(pending, sink)
        pending = (host, ev)
    if pending is not None:
        n_frames += _flush(pending, sink)
    return n_frames


def _flush(pending, sink):
    host, ev = pending
    if ev is not None:
        ev.synchronize()
    if sink is not None:
        for frame in host.numpy():
            sink.add(frame)
    return host.shape[0]


def load_models(checkpoint_path, device):
    checkpoint_object = utils.load_checkpoint_file(checkpoint_path)
    saved_args = copy.copy(checkpoint_object['args'])
    saved_args.finetune = True
    saved_args.inference = True
    saved_args.world_size = 1
    saved_args.num_workers = 1
    saved_args.device = device
    embedder, generator, _, running_averages, _, _, _ = utils.load_model_from_checkpoint(checkpoint_object, saved_args)
    if 'embedder' in running_averages:
        embedder.load_state_dict(running_averages['embedder'])
    if 'generator' in running_averages:
        generator.load_state_dict(running_averages['generator'])
    eval_mode = getattr(saved_args, 'set_eval_mode_in_test', True)
    embedder.train(not eval_mode)
    generator.train(not eval_mode)
    return embedder, generator, saved_args


def main():
    logging.basicConfig(level=logging.INFO, stream=sys.stdout, format="%(asctime)s - %(levelname)s - %(message)s")
    logger = logging.getLogger('drive')
    ap = argparse.ArgumentParser(description="Render driving frames with a fine-tuned model.")
    ap.add_argument('checkpoint_path', type=Path)
    ap.add_argument('data_root', type=Path, nargs='?', default=None,
                    help="root folder that contains the driver image folders (omit for synthetic drivers)")
    ap.add_argument('--images_paths', type=Path, nargs='+', default=[])
    ap.add_argument('--destination', type=Path, required=True)
    ap.add_argument('--batch_size', type=int, default=1)
    ap.add_argument('--synthetic_frames', type=int, default=0, help="drive with N synthetic U[0,1) frames")
    ap.add_argument('--format', choices=['avi', 'png'], default='avi')
    args = ap.parse_args()

    device = 'cuda:0' if torch.cuda.is_available() else 'cpu'
    logger.info(f"Will run on device '{device}'")
    embedder, generator, saved_args = load_models(args.checkpoint_path, device)
    args.destination.mkdir(parents=True, exist_ok=True)

    jobs = []
    for p in args.images_paths:
        folder = (args.data_root / getattr(saved_args, 'img_dir', 'images-cropped') / p) if args.data_root else p
        if not folder.is_dir():
            folder = args.data_root / p if args.data_root else p
        jobs.append((str(p).replace('/', '_'), load_driver_frames(folder, saved_args.image_size)))
    if args.synthetic_frames:
        g = torch.Generator().manual_seed(123)
        s = saved_args.image_size
        jobs.append(('synthetic', (torch.rand(3, s, s, generator=g) for _ in range(args.synthetic_frames))))
    for name, frames in jobs:
        out = args.destination / (name + ('.avi' if args.format == 'avi' else ''))
        sink = FrameWriter(out)
        n = render(embedder, generator, batches(frames, args.batch_size), device, sink)
        sink.close()
        logger.info(f"Wrote {n} frames to {out}")


if __name__ == '__main__':
    main()
