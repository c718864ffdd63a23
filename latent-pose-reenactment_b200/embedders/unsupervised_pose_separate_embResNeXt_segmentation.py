"""Identity / pose embedder plugin — drop-in for the reference's
`embedders/unsupervised_pose_separate_embResNeXt_segmentation.py` (Wrapper :7-16, Embedder :19-63).

identity = torchvision ResNeXt50-32x4d -> `embed_channels`, averaged over the K identity frames;
pose     = torchvision MobileNetV2     -> `pose_embedding_size`.
The two torchvision modules are kept as the OWNERS of parameters and buffers, so the state_dict (634 torchvision keys
under `identity_encoder.` / `pose_encoder.`) is the reference's; on a B200 their arithmetic — forward and backward —
runs as libb200lp kernel schedules over those parameters (embedders/resnext_native.py: tcgen05 1x1 / stem GEMMs, FP32
grouped convolutions, fused BatchNorm kernels; embedders/mobilenet_native.py).
"""
import os

import torch
from torch import nn


class Wrapper:
    @staticmethod
    def get_args(parser):
        parser.add('--average_function', type=str, default='sum', help='sum|max')

    @staticmethod
    def get_net(args):
        net = Embedder(args.embed_channels, args.pose_embedding_size, args.average_function)
        return net.to(args.device)


class Embedder(nn.Module):
    def __init__(self, identity_embedding_size, pose_embedding_size, average_function):
        super().__init__()
        import torchvision
        self.identity_embedding_size = identity_embedding_size
        self.pose_embedding_size = pose_embedding_size
        self.identity_encoder = torchvision.models.resnext50_32x4d(num_classes=identity_embedding_size)
        self.pose_encoder = torchvision.models.mobilenet_v2(num_classes=pose_embedding_size)
        if average_function not in ('sum', 'max'):
            raise ValueError("Incorrect `average_function` argument, expected `sum` or `max`")
        self.average_function = average_function
        self.finetuning = False

    def enable_finetuning(self, data_dict=None):
        self.finetuning = True

    def get_identity_embedding(self, data_dict):
        inputs = data_dict['enc_rgbs']
        batch_size, num_faces, c, h, w = inputs.shape
        frames = inputs.reshape(-1, c, h, w)
        if self._native_identity_path(frames):
            from embedders import resnext_native
            per_frame = resnext_native.apply(self.identity_encoder, frames).view(batch_size, num_faces, -1)
        else:
            per_frame = self.identity_encoder(frames).view(batch_size, num_faces, -1)
        assert per_frame.shape[2] == self.identity_embedding_size
        data_dict['embeds'] = per_frame.mean(1) if self.average_function == 'sum' else per_frame.max(1)[0]
        data_dict['embeds_elemwise'] = per_frame

    def get_pose_embedding(self, data_dict):
        x = data_dict['pose_input_rgbs'][:, 0]
        if self._native_pose_path(x):
            # libb200lp schedule of the same MobileNetV2 (csrc/mobilenet.cu, mobilenet_bwd.cu): forward, and — when the
            # encoder's parameters are being trained (meta-training) — backward as one autograd node
            from embedders import mobilenet_native
            data_dict['pose_embedding'] = mobilenet_native.apply(self.pose_encoder, x)
        else:
            data_dict['pose_embedding'] = self.pose_encoder(x)

    def _native_identity_path(self, x):
        """Kernel schedule of the identity encoder (forward + backward) for float32 CUDA frames whose planes the
        tensor-core kernels tile (powers of two, >= 64 x 64, a multiple of 8 frames)."""
        if not x.is_cuda or x.dtype != torch.float32 or os.environ.get('B200LP_TORCH_IDENTITY_ENCODER'):
            return False
        n, _, h, w = x.shape
        if h < 64 or w < 64 or (h & (h - 1)) or (w & (w - 1)) or n % 8:
            return False
        from embedders import resnext_native
        if self.__dict__.get('_native_id_ok') is None:
            self.__dict__['_native_id_ok'] = resnext_native.supported(self.identity_encoder)
        return self.__dict__['_native_id_ok']

    def _native_pose_path(self, x):
        """Kernel schedule of the pose encoder for float32 CUDA frames (any size the torchvision module takes)."""
        if not x.is_cuda or x.dtype != torch.float32 or os.environ.get('B200LP_TORCH_POSE_ENCODER'):
            return False
        if torch.is_grad_enabled() and x.requires_grad:
            return False          # no gradient w.r.t. the image is implemented (the reference never needs one)
        from embedders import mobilenet_native
        if self.__dict__.get('_native_ok') is None:
            self.__dict__['_native_ok'] = mobilenet_native.supported(self.pose_encoder)
        return self.__dict__['_native_ok']

    def forward(self, data_dict):
        if not self.finetuning:
            self.get_identity_embedding(data_dict)
        self.get_pose_embedding(data_dict)
