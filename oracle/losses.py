"""Loss helpers restated from the reference (oracle; test infrastructure only).

Same call signatures as the reference helpers so tests can swap them in.
"""
import torch


def get_fm_loss(real_feats, fake_feats, criterion, device):
    """image_translation.py:136-144 -- per D layer: criterion((mean_b r - mean_b f)^2, ones).
    With ``nn.HingeEmbeddingLoss`` and target +1 the criterion is the plain mean."""
    losses = 0
    for real_feat, fake_feat in zip(real_feats, fake_feats):
        d = real_feat.mean(0) - fake_feat.mean(0)
        l2 = d * d
        losses = losses + criterion(l2, torch.ones(l2.size()).to(device))
    return losses


def get_fm_loss_angle(real_feats, fake_feats, criterion, device):
    """angle_pairing.py:111-120 -- same, skipping the first returned feature map."""
    return get_fm_loss(real_feats[1:], fake_feats[1:], criterion, device)


def get_gan_loss(dis_real, dis_fake, criterion, device):
    """image_translation.py:146-168 -- BCE against ones/zeros on [B,1] views.
    dis = 0.5*(BCE(Dr,1)+BCE(Df,0)); gen = BCE(Df,1)."""
    batch_size = dis_real.size(0)
    if dis_real.dim() > 2:
        dis_real = dis_real.view(batch_size, -1)
    if dis_fake.dim() > 2:
        dis_fake = dis_fake.view(batch_size, -1)
    ones = torch.ones(batch_size, 1).to(device)
    zeros = torch.zeros(batch_size, 1).to(device)
    dis_loss = (criterion(dis_real, ones) + criterion(dis_fake, zeros)) * 0.5
    gen_loss = criterion(dis_fake, ones)
    return dis_loss, gen_loss
