"""Physical (scene) attention parameter containers (reference: mggan/model/modules/cnn.py).
Module tree and parameter names follow the reference (`CNN.encoder.ConvBlock_i.Block.{Conv_1,BN_1}`,
`cnn_attention.{0,2}`); forward = the `mggan_scene_*` kernels."""
import torch
from torch import nn

from mggan import kernels as K


class Conv_Blocks(nn.Module):
    def __init__(self, input_dim, output_dim):
        super().__init__()
        self.Block = nn.Sequential()
        self.Block.add_module("Conv_1", nn.Conv2d(input_dim, output_dim, 3, 1, 1))
        self.Block.add_module("BN_1", nn.BatchNorm2d(output_dim))
        self.Block.add_module("NonLin_1", nn.ReLU())
        self.Block.add_module("Pool", nn.MaxPool2d(kernel_size=(2, 2), stride=(2, 2)))


class CNN(nn.Module):
    """2 conv blocks, in_channels=4 (reference cnn.py:178-282 with the arguments AttentionGlobal uses)."""

    def __init__(self, channels_cnn=4, in_channels=4):
        super().__init__()
        self.encoder = nn.Sequential()
        self.encoder.add_module("ConvBlock_1", Conv_Blocks(in_channels, channels_cnn))
        self.encoder.add_module("ConvBlock_2", Conv_Blocks(channels_cnn, channels_cnn))
        self.bootleneck_channel = channels_cnn
        self.bottleneck_dim = 64

        def init_kaiming(m):          # reference cnn.py:257-262 (non_lin='relu')
            if type(m) in [nn.Conv2d, nn.ConvTranspose2d]:
                torch.nn.init.kaiming_normal_(m.weight, mode="fan_in")
                m.bias.data.fill_(0.01)

        self.apply(init_kaiming)


class AttentionGlobal(nn.Module):
    """reference cnn.py:101-116 (+ VisualNetwork/AttentionNetwork set-up :28-98)."""

    def __init__(self, noise_attention_dim=0, PhysFeature=True, num_layers=2, channels_cnn=4, mlp_dim=32):
        super().__init__()
        if num_layers != 2 or channels_cnn not in (8, 16) or mlp_dim != 32:
            raise NotImplementedError("B200 path: 2 conv blocks with 8 or 16 channels, attention MLP width 32")
        self.noise_attention_dim, self.mlp_dim = noise_attention_dim, mlp_dim
        self.CNN = CNN(channels_cnn=channels_cnn, in_channels=4)
        self.cnn_attention = nn.Sequential(nn.Linear(channels_cnn, mlp_dim), nn.LeakyReLU(),
                                           nn.Linear(mlp_dim, channels_cnn))
        self.stat_group = None          # process group sharing BatchNorm statistics (data-parallel runs)
        self.memo = None                # dict while a caller guarantees constant weights (kernels.scene_attention)

    def forward(self, features, rows=None, rows_key=None):
        """features (N,4,33,33) -> (N,64).  `rows` (int32) gathers image rows without copying them;
        `rows_key` identifies that selection for the per-batch patch-statistics cache."""
        if rows is not None and rows_key is None:
            rows_key = ("rows", rows.data_ptr(), rows._version, int(rows.numel()))
        return K.scene_attention(features, self, rows, self.stat_group, rows_key)
