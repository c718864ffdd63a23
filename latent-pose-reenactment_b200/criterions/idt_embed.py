"""VGG-Face identity criterion — drop-in for the reference's `criterions/idt_embed.py:14-104`:
centre crop (factor 1/1.8) resampled to full size by affine_grid + grid_sample(bilinear, reflection), then the
VGG16-Face perceptual loss.  The VGG part runs on libb200lp kernels (same path as criterions/perceptual.py);
the crop is the one gather-interpolate step still issued through torch."""
import torch

from criterions.common.perceptual_loss import PerceptualLoss


class Wrapper:
    @staticmethod
    def get_args(parser):
        parser.add('--idt_embed_weight', type=float, default=2e-3)

    @staticmethod
    def get_net(args):
        criterion = Criterion(args.idt_embed_weight, args.vgg_weights_dir)
        return criterion.to(args.device)


class Criterion(torch.nn.Module):
    def __init__(self, idt_embed_weight, vgg_weights_dir):
        super().__init__()
        self.idt_embed_crit = PerceptualLoss(idt_embed_weight, vgg_weights_dir, net='face').eval()
        self._center_grid = {}      # (batch, H, W, device) -> sampling grid of the constant centre crop

    def forward(self, data_dict):
        fake_rgb = data_dict['fake_rgbs']
        real_rgb = data_dict['target_rgbs']
        if len(fake_rgb.shape) > 4:
            fake_rgb = fake_rgb[:, 0]
        if len(real_rgb.shape) > 4:
            real_rgb = real_rgb[:, 0]

        if 'dec_keypoints' in data_dict:
            bboxes = compute_bboxes_from_keypoints(data_dict['dec_keypoints'])
            h, w = real_rgb.shape[2:]
            bboxes[:, 0:2] *= h
            bboxes[:, 2:4] *= w
            fake_cropped = crop_and_resize(fake_rgb, bboxes)
            real_cropped = crop_and_resize(real_rgb, bboxes)
        else:
            # the centre crop (factor 1/1.8) is the same for every batch: build its sampling grid once
            # (host->device tensor construction is not something to repeat per step, nor capturable in a CUDA graph)
            key = (len(real_rgb),) + tuple(real_rgb.shape[2:]) + (str(real_rgb.device),)
            grid = self._center_grid.get(key)
            if grid is None:
                crop_factor = 1 / 1.8
                h, w = real_rgb.shape[2:]
                t = h * (1 - crop_factor) / 2
                l = w * (1 - crop_factor) / 2
                bboxes = torch.tensor([[t, h - t, l, w - l]], dtype=torch.float32, device=real_rgb.device)
                # crop kernel (bilinear gather forward, gather-form backward): the box is strictly inside the image
                from b200lp import ops
                assert ops.crop_boxes_inside([[t, h - t, l, w - l]], h, w, h, w)
                grid = ('boxes', bboxes.expand(len(real_rgb), 4).contiguous())
                self._center_grid[key] = grid
            if isinstance(grid, tuple):
                from b200lp import ops
                fake_cropped = ops.crop_bilinear(fake_rgb, grid[1])
                real_cropped = ops.crop_bilinear(real_rgb.detach(), grid[1])
            else:
                fake_cropped = sample(fake_rgb, grid)
                real_cropped = sample(real_rgb, grid)
        return {'VGGFace': self.idt_embed_crit(fake_cropped, real_cropped)}


def sampling_grid(bboxes, image_shape, target_size=None):
    """bboxes B x 4 = [t, b, l, r] in pixels -> affine sampling grid (B, h, w, 2) for images of `image_shape`."""
    t, b, l, r = bboxes.t().float()
    batch_size, num_channels, h, w = image_shape
    theta = torch.zeros(batch_size, 2, 3, dtype=torch.float32, device=bboxes.device)
    theta[:, 0, 0] = (r - l) / w
    theta[:, 1, 1] = (b - t) / h
    theta[:, 0, 2] = (l + r) / w - 1
    theta[:, 1, 2] = (t + b) / h - 1
    return torch.nn.functional.affine_grid(theta, (batch_size, num_channels) + tuple(target_size or (h, w)),
                                           align_corners=False)


def sample(images, grid):
    return torch.nn.functional.grid_sample(images, grid.to(images.dtype), mode='bilinear', padding_mode='reflection',
                                           align_corners=False)


def crop_and_resize(images, bboxes, target_size=None):
    """images B x C x H x W; bboxes B x 4 = [t, b, l, r] in pixels -> crops resized to `target_size` (default H x W)."""
    return sample(images, sampling_grid(bboxes, images.shape, target_size))


def compute_bboxes_from_keypoints(keypoints):
    """keypoints B x 68*2 -> rough face boxes B x 4 (t, b, l, r), as the reference (:85-104)."""
    x, y = keypoints.float().view(-1, 68, 2).transpose(0, 2)
    face_height = y[8] - y[27]
    b = y[8] + face_height * 0.2
    t = y[27] - face_height * 0.47
    midpoint_x = (x.min() + x.max()) / 2
    half_height = (b - t) * 0.5
    return torch.stack([t, b, midpoint_x - half_height, midpoint_x + half_height], dim=1)
