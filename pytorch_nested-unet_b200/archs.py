"""Drop-in for the reference's `archs` module (finished/archs1.py:11-143): VGGBlock, UNet, NestedUNet.

Same constructor signatures, same sub-module / parameter / buffer names and shapes (state_dict-compatible, fp32 OIHW
masters, default nn.Conv2d / nn.BatchNorm2d initialisation consumed in the same RNG order), same forward contract
(NCHW fp32 CUDA input -> logits tensor, or a list of four with deep supervision).  What runs underneath is the
sm_100a engine in `engine.py`: no nn.Conv2d / nn.BatchNorm2d forward is ever called, nothing falls back to ATen.
"""
import torch
from torch import nn

import engine

__all__ = ['UNet', 'NestedUNet']


class VGGBlock(nn.Module):
    """conv3x3 -> BN -> ReLU, twice (finished/archs1.py:14-32). The sub-modules only hold parameters and buffers."""

    def __init__(self, in_channels, middle_channels, out_channels):
        super().__init__()
        self.relu = nn.ReLU(inplace=True)
        self.conv1 = nn.Conv2d(in_channels, middle_channels, 3, padding=1)
        self.bn1 = nn.BatchNorm2d(middle_channels)
        self.conv2 = nn.Conv2d(middle_channels, out_channels, 3, padding=1)
        self.bn2 = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        """finished/archs1.py:23-32 on its own (inside UNet / NestedUNet the engine schedules the block's launches itself): NCHW fp32
        in, NCHW fp32 out, trainable (engine.run_block)."""
        return engine.run_block(self, x)


class _EngineModel(nn.Module):
    """Shared forward of UNet / NestedUNet: hands the module's parameters to the fused engine."""

    _graph = None
    _heads = None

    def forward(self, input):
        """input: fp32 NCHW [N,C,H,W] as in the reference (trains.py:117) — or, with the device-side input pipeline, the raw uint8
        NHWC image batch [N,H,W,C] (normalised on the GPU with the loader's arithmetic, see set_input_normalization)."""
        return engine.run_model(self, input)

    def set_input_normalization(self, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), max_pixel=255.0, post_div=255.0):
        """How a uint8 image batch becomes the network input: ((u8/max_pixel - mean)/std)/post_div.  The defaults are the reference
        loader's: albumentations Normalize() (trains.py:266) followed by dataset.py:71's `/ 255`."""
        self.__dict__["_nunet_input_norm"] = dict(mean=tuple(mean), std=tuple(std), max_pixel=float(max_pixel), post_div=float(post_div))


class UNet(_EngineModel):
    """finished/archs1.py:35-71 (accepts and ignores deep_supervision, as the reference does)."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        nb_filter = [32, 64, 128, 256, 512]
        self.pool = nn.MaxPool2d(2, 2)
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)

        self.conv0_0 = VGGBlock(input_channels, nb_filter[0], nb_filter[0])
        self.conv1_0 = VGGBlock(nb_filter[0], nb_filter[1], nb_filter[1])
        self.conv2_0 = VGGBlock(nb_filter[1], nb_filter[2], nb_filter[2])
        self.conv3_0 = VGGBlock(nb_filter[2], nb_filter[3], nb_filter[3])
        self.conv4_0 = VGGBlock(nb_filter[3], nb_filter[4], nb_filter[4])

        self.conv3_1 = VGGBlock(nb_filter[3] + nb_filter[4], nb_filter[3], nb_filter[3])
        self.conv2_2 = VGGBlock(nb_filter[2] + nb_filter[3], nb_filter[2], nb_filter[2])
        self.conv1_3 = VGGBlock(nb_filter[1] + nb_filter[2], nb_filter[1], nb_filter[1])
        self.conv0_4 = VGGBlock(nb_filter[0] + nb_filter[1], nb_filter[0], nb_filter[0])

        self.final = nn.Conv2d(nb_filter[0], num_classes, kernel_size=1)

        self.input_channels = input_channels
        self.num_classes = num_classes
        self._graph = engine.unet_graph()
        self._heads = [("final", "0_4")]
        self._returns_list = False


class NestedUNet(_EngineModel):
    """finished/archs1.py:74-143."""

    def __init__(self, num_classes, input_channels=3, deep_supervision=False, **kwargs):
        super().__init__()
        nb_filter = [32, 64, 128, 256, 512]
        self.deep_supervision = deep_supervision
        self.pool = nn.MaxPool2d(2, 2)
        self.up = nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)

        self.conv0_0 = VGGBlock(input_channels, nb_filter[0], nb_filter[0])
        self.conv1_0 = VGGBlock(nb_filter[0], nb_filter[1], nb_filter[1])
        self.conv2_0 = VGGBlock(nb_filter[1], nb_filter[2], nb_filter[2])
        self.conv3_0 = VGGBlock(nb_filter[2], nb_filter[3], nb_filter[3])
        self.conv4_0 = VGGBlock(nb_filter[3], nb_filter[4], nb_filter[4])

        self.conv0_1 = VGGBlock(nb_filter[0] + nb_filter[1], nb_filter[0], nb_filter[0])
        self.conv1_1 = VGGBlock(nb_filter[1] + nb_filter[2], nb_filter[1], nb_filter[1])
        self.conv2_1 = VGGBlock(nb_filter[2] + nb_filter[3], nb_filter[2], nb_filter[2])
        self.conv3_1 = VGGBlock(nb_filter[3] + nb_filter[4], nb_filter[3], nb_filter[3])

        self.conv0_2 = VGGBlock(nb_filter[0] * 2 + nb_filter[1], nb_filter[0], nb_filter[0])
        self.conv1_2 = VGGBlock(nb_filter[1] * 2 + nb_filter[2], nb_filter[1], nb_filter[1])
        self.conv2_2 = VGGBlock(nb_filter[2] * 2 + nb_filter[3], nb_filter[2], nb_filter[2])

        self.conv0_3 = VGGBlock(nb_filter[0] * 3 + nb_filter[1], nb_filter[0], nb_filter[0])
        self.conv1_3 = VGGBlock(nb_filter[1] * 3 + nb_filter[2], nb_filter[1], nb_filter[1])

        self.conv0_4 = VGGBlock(nb_filter[0] * 4 + nb_filter[1], nb_filter[0], nb_filter[0])

        if self.deep_supervision:
            self.final1 = nn.Conv2d(nb_filter[0], num_classes, kernel_size=1)
            self.final2 = nn.Conv2d(nb_filter[0], num_classes, kernel_size=1)
            self.final3 = nn.Conv2d(nb_filter[0], num_classes, kernel_size=1)
            self.final4 = nn.Conv2d(nb_filter[0], num_classes, kernel_size=1)
        else:
            self.final = nn.Conv2d(nb_filter[0], num_classes, kernel_size=1)

        self.input_channels = input_channels
        self.num_classes = num_classes
        self._graph = engine.nested_graph()
        if self.deep_supervision:
            self._heads = [("final1", "0_1"), ("final2", "0_2"), ("final3", "0_3"), ("final4", "0_4")]
        else:
            self._heads = [("final", "0_4")]
        self._returns_list = bool(self.deep_supervision)
