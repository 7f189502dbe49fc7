"""Plain-PyTorch segmentation networks used as the *consumer* around the hot path in
``bench.py`` and the GPU tests.  NOT part of the accelerated path and not a rebuild of the
reference's model zoo (out of scope): the attack only needs
``model(x) -> logits [B,C,H,W]`` at input resolution and autograd back to ``x``; these run on
stock cuDNN / cuBLAS.

``upernet_convnext(variant, n_cls)`` builds the architecture BASELINE.json's configs name
(UperNet decode head over ConvNeXt-T/S with the two-conv "CvSt" stem, bilinear up-sampling of
the logits to the input size), random-initialised -- there is no network for checkpoints.
Shapes follow the published ConvNeXt / UperNet designs (depths 3-3-9-3 or 3-3-27-3, widths
96-192-384-768, PSP pool scales 1-2-3-6, 512 decoder channels, 256-channel FCN aux head).
``segmenter_vit(variant, n_cls)`` is config 3's Segmenter (ViT-S/16: 384 wide, 6 heads, 12 layers;
mask-transformer decoder with 2 layers; x16 bilinear up-sampling of the class masks).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

CONVNEXT = {"T": (3, 3, 9, 3), "S": (3, 3, 27, 3)}
WIDTHS = (96, 192, 384, 768)


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over the channel axis of an NCHW tensor."""

    def forward(self, x):
        return F.layer_norm(x.permute(0, 2, 3, 1), self.normalized_shape, self.weight, self.bias,
                            self.eps).permute(0, 3, 1, 2)


class NeXtBlock(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.dw = nn.Conv2d(dim, dim, 7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.fc2 = nn.Linear(4 * dim, dim)
        self.scale = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        y = self.dw(x).permute(0, 2, 3, 1)
        y = self.fc2(F.gelu(self.fc1(self.norm(y)))) * self.scale
        return x + y.permute(0, 3, 1, 2)


class ConvNeXtCvSt(nn.Module):
    def __init__(self, depths):
        super().__init__()
        w = WIDTHS
        stem = nn.Sequential(nn.Conv2d(3, w[0] // 2, 3, 2, 1), LayerNorm2d(w[0] // 2, eps=1e-6), nn.GELU(),
                             nn.Conv2d(w[0] // 2, w[0], 3, 2, 1), LayerNorm2d(w[0], eps=1e-6), nn.GELU())
        self.down = nn.ModuleList([stem] + [
            nn.Sequential(LayerNorm2d(w[i], eps=1e-6), nn.Conv2d(w[i], w[i + 1], 2, 2)) for i in range(3)])
        self.stages = nn.ModuleList([nn.Sequential(*[NeXtBlock(w[i]) for _ in range(d)])
                                     for i, d in enumerate(depths)])
        self.out_norm = nn.ModuleList([LayerNorm2d(c, eps=1e-6) for c in w])
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        feats = []
        for d, s, n in zip(self.down, self.stages, self.out_norm):
            x = s(d(x))
            feats.append(n(x))
        return feats


def conv_bn_relu(cin, cout, k):
    return nn.Sequential(nn.Conv2d(cin, cout, k, padding=k // 2, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class UperHead(nn.Module):
    def __init__(self, n_cls, ch=512, scales=(1, 2, 3, 6)):
        super().__init__()
        w = WIDTHS
        self.psp = nn.ModuleList([nn.Sequential(nn.AdaptiveAvgPool2d(s), conv_bn_relu(w[-1], ch, 1)) for s in scales])
        self.bottleneck = conv_bn_relu(w[-1] + len(scales) * ch, ch, 3)
        self.lateral = nn.ModuleList([conv_bn_relu(c, ch, 1) for c in w[:-1]])
        self.fpn = nn.ModuleList([conv_bn_relu(ch, ch, 3) for _ in w[:-1]])
        self.fuse = conv_bn_relu(len(w) * ch, ch, 3)
        self.classifier = nn.Conv2d(ch, n_cls, 1)

    def forward(self, feats, interp=None):
        def up(t, size):
            if interp is not None and t.is_cuda and t.dtype == torch.float32:
                return interp(t, size)
            return F.interpolate(t, size=size, mode="bilinear", align_corners=False)

        top = feats[-1]
        size = top.shape[2:]
        pooled = [top] + [up(p(top), size) for p in self.psp]
        lat = [l(f) for l, f in zip(self.lateral, feats)] + [self.bottleneck(torch.cat(pooled, 1))]
        for i in range(len(lat) - 1, 0, -1):
            lat[i - 1] = lat[i - 1] + up(lat[i], lat[i - 1].shape[2:])
        outs = [f(l) for f, l in zip(self.fpn, lat)] + [lat[-1]]
        size0 = outs[0].shape[2:]
        outs = [outs[0]] + [up(o, size0) for o in outs[1:]]
        return self.classifier(self.fuse(torch.cat(outs, 1)))


class UperNetConvNeXt(nn.Module):
    """``fast_upsample=True`` routes the final logit up-sampling (the only [B,C,H,W]-sized op of
    the model) through robseg's kernels instead of ``F.interpolate``;
    ``fast_upsample="all"`` also routes the bilinear up-samplings inside the decode head."""

    def __init__(self, variant="T", n_cls=150, fast_upsample=False):
        super().__init__()
        self.fast_upsample = fast_upsample
        self.backbone = ConvNeXtCvSt(CONVNEXT[variant])
        self.decode_head = UperHead(n_cls)
        self.aux_head = nn.Sequential(conv_bn_relu(WIDTHS[2], 256, 3), nn.Conv2d(256, n_cls, 1))

    def forward(self, x, lbl=None):
        feats = self.backbone(x)
        fast = self.fast_upsample and x.is_cuda
        if fast:
            from . import ops
        low = self.decode_head(feats, ops.upsample_bilinear if fast and self.fast_upsample == "all" else None)
        if fast and low.dtype == torch.float32:
            logits = ops.upsample_bilinear(low, x.shape[2:])
        else:
            logits = F.interpolate(low, size=x.shape[2:], mode="bilinear", align_corners=False)
        if lbl is None:
            return logits
        aux = F.interpolate(self.aux_head(feats[2]), size=x.shape[2:], mode="bilinear", align_corners=False)
        loss = F.cross_entropy(logits, lbl, ignore_index=-1) + 0.4 * F.cross_entropy(aux, lbl, ignore_index=-1)
        return (loss, logits) if self.training else logits


def upernet_convnext(variant="T", n_cls=150, fast_upsample=False):
    return UperNetConvNeXt(variant, n_cls, fast_upsample)


VIT = {"S": (384, 6, 12), "B": (768, 12, 12), "L": (1024, 16, 24)}  # width, heads, layers (patch 16)


class TokenBlock(nn.Module):
    """Pre-norm transformer block (attention through F.scaled_dot_product_attention)."""

    def __init__(self, dim, heads, hidden):
        super().__init__()
        self.heads = heads
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.qkv, self.proj = nn.Linear(dim, 3 * dim), nn.Linear(dim, dim)
        self.fc1, self.fc2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        B, N, D = x.shape
        q, k, v = self.qkv(self.norm1(x)).view(B, N, 3, self.heads, D // self.heads).permute(2, 0, 3, 1, 4)
        x = x + self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, D))
        return x + self.fc2(F.gelu(self.fc1(self.norm2(x))))


class SegmenterViT(nn.Module):
    """Segmenter (ViT encoder, patch 16, + 2-layer mask-transformer decoder, bilinear x16 up-sampling
    of the class masks to the input size): the architecture BASELINE.json's config 3 names, random
    init.  ``fast_upsample`` routes the up-sampling through robseg's kernels."""

    def __init__(self, variant="S", n_cls=150, image_size=512, dec_layers=2, fast_upsample=False):
        super().__init__()
        dim, heads, layers = VIT[variant]
        self.patch, self.n_cls, self.fast_upsample = 16, n_cls, fast_upsample
        self.embed = nn.Conv2d(3, dim, 16, 16)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos = nn.Parameter(torch.zeros(1, 1 + (image_size // 16) ** 2, dim))
        self.encoder = nn.Sequential(*[TokenBlock(dim, heads, 4 * dim) for _ in range(layers)])
        self.enc_norm = nn.LayerNorm(dim)
        self.proj_dec = nn.Linear(dim, dim)
        self.cls_emb = nn.Parameter(torch.zeros(1, n_cls, dim))
        self.decoder = nn.Sequential(*[TokenBlock(dim, heads, 4 * dim) for _ in range(dec_layers)])
        self.dec_norm = nn.LayerNorm(dim)
        self.proj_patch = nn.Parameter(dim ** -0.5 * torch.randn(dim, dim))
        self.proj_classes = nn.Parameter(dim ** -0.5 * torch.randn(dim, dim))
        self.mask_norm = nn.LayerNorm(n_cls)
        for t in (self.cls_token, self.pos, self.cls_emb):
            nn.init.trunc_normal_(t, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)

    def forward(self, x):
        B, _, H, W = x.shape
        gh, gw = H // self.patch, W // self.patch
        if (gh * gw + 1) != self.pos.shape[1] or H % self.patch or W % self.patch:
            raise ValueError("SegmenterViT was built for a different image size")
        t = self.embed(x).flatten(2).transpose(1, 2)
        t = torch.cat([self.cls_token.expand(B, -1, -1), t], 1) + self.pos
        t = self.enc_norm(self.encoder(t))[:, 1:]
        t = torch.cat([self.proj_dec(t), self.cls_emb.expand(B, -1, -1)], 1)
        t = self.dec_norm(self.decoder(t))
        patches, classes = t[:, :-self.n_cls] @ self.proj_patch, t[:, -self.n_cls:] @ self.proj_classes
        patches = patches / patches.norm(dim=-1, keepdim=True)
        classes = classes / classes.norm(dim=-1, keepdim=True)
        masks = self.mask_norm(patches @ classes.transpose(1, 2))
        low = masks.transpose(1, 2).reshape(B, self.n_cls, gh, gw)
        if self.fast_upsample and low.is_cuda and low.dtype == torch.float32:
            from . import ops

            return ops.upsample_bilinear(low, (H, W))
        return F.interpolate(low, size=(H, W), mode="bilinear", align_corners=False)


def segmenter_vit(variant="S", n_cls=150, image_size=512, fast_upsample=False):
    return SegmenterViT(variant, n_cls, image_size, fast_upsample=fast_upsample)


class TinySegNet(nn.Module):
    """3 -> hidden -> C two-conv net for smoke tests (milliseconds on any device)."""

    def __init__(self, n_cls=21, hidden=8, seed=0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.c1 = nn.Conv2d(3, hidden, 3, padding=1)
        self.c2 = nn.Conv2d(hidden, n_cls, 3, padding=1)
        with torch.no_grad():
            for p in self.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * (0.6 if p.dim() > 1 else 0.1))

    def forward(self, x):
        return self.c2(torch.tanh(self.c1(x - 0.5)))
