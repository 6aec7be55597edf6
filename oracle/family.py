"""Plain ``torch.nn`` model family (oracle; test infrastructure only).

Restates the topology of the reference ``model.py`` with the depth made a function
of the image size:

    n_down = log2(S) - 2 stride-2 4x4 layers, channels ch[i] = 64 * 2**min(i, 5)

At S=512 this is layer-for-layer the reference network (``model.py:8-35`` for the
Discriminator, ``model.py:80-143`` for the Generator): same module order, same
state-dict keys, same parameter count.  The reference itself only runs at S=512
(SURVEY.md F1); smaller sizes truncate the same pattern.
"""
import math

import torch.nn as nn


def family_channels(image_size: int):
    """Trunk channel list for an image size: 512 -> [64,128,256,512,1024,2048,2048]."""
    n_down = int(round(math.log2(image_size))) - 2
    if 2 ** (n_down + 2) != image_size or n_down < 2:
        raise ValueError(f"image_size must be a power of two >= 16, got {image_size}")
    return [64 * 2 ** min(i, 5) for i in range(n_down)]


class Discriminator(nn.Module):
    """model.py:5-69 -- conv1..convN (4x4 s2 p1), BN on 2..n_down, LeakyReLU(0.2),
    final 4x4 valid conv + Sigmoid; returns (prob, [post-activation feats 2..n_down])."""

    def __init__(self, image_size: int = 512):
        super().__init__()
        ch = family_channels(image_size)
        self.image_size = image_size
        self.n_down = len(ch)
        self.conv1 = nn.Conv2d(3, ch[0], 4, 2, 1, bias=False)          # model.py:8
        self.relu1 = nn.LeakyReLU(0.2, inplace=True)
        for i in range(1, self.n_down):                                   # model.py:11-33
            k = i + 1
            setattr(self, f"conv{k}", nn.Conv2d(ch[i - 1], ch[i], 4, 2, 1, bias=False))
            setattr(self, f"bn{k}", nn.BatchNorm2d(ch[i]))
            setattr(self, f"relu{k}", nn.LeakyReLU(0.2, inplace=True))
        setattr(self, f"conv{self.n_down + 1}", nn.Conv2d(ch[-1], 1, 4, 1, 0, bias=False))  # model.py:35
        self.sigmoid = nn.Sigmoid()

    def forward(self, x):                                                 # model.py:38-69
        h = self.relu1(self.conv1(x))
        feats = []
        for k in range(2, self.n_down + 1):
            h = getattr(self, f"relu{k}")(getattr(self, f"bn{k}")(getattr(self, f"conv{k}")(h)))
            feats.append(h)
        out = self.sigmoid(getattr(self, f"conv{self.n_down + 1}")(h))
        return out, feats


class Generator(nn.Module):
    """model.py:72-225 -- encoder (n_down s2 convs + 4x4 valid conv to 100 ch) and the
    mirrored ConvTranspose2d decoder.  ``extra_layers`` is accepted and ignored, as in
    the reference where both branches build the same network."""

    def __init__(self, extra_layers: bool = False, image_size: int = 512):
        super().__init__()
        ch = family_channels(image_size)
        self.image_size = image_size
        self.main = None                                                  # model.py:215
        enc = [nn.Conv2d(3, ch[0], 4, 2, 1, bias=False), nn.LeakyReLU(0.2, inplace=True)]
        for i in range(1, len(ch)):
            enc += [nn.Conv2d(ch[i - 1], ch[i], 4, 2, 1, bias=False), nn.BatchNorm2d(ch[i]),
                    nn.LeakyReLU(0.2, inplace=True)]
        enc += [nn.Conv2d(ch[-1], 100, 4, 1, 0, bias=False), nn.BatchNorm2d(100),
                nn.LeakyReLU(0.2, inplace=True)]
        dec = [nn.ConvTranspose2d(100, ch[-1], 4, 1, 0, bias=False), nn.BatchNorm2d(ch[-1]), nn.ReLU(True)]
        for i in range(len(ch) - 1, 0, -1):
            dec += [nn.ConvTranspose2d(ch[i], ch[i - 1], 4, 2, 1, bias=False), nn.BatchNorm2d(ch[i - 1]),
                    nn.ReLU(True)]
        dec += [nn.ConvTranspose2d(ch[0], 3, 4, 2, 1, bias=False), nn.Sigmoid()]
        self.encoder = nn.Sequential(*enc)
        self.decoder = nn.Sequential(*dec)

    def forward(self, x):                                                 # model.py:217-225
        return self.decoder(self.encoder(x))
