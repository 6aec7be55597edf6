"""The reference train step, restated (oracle; test infrastructure only).

The reference keeps the step inline in ``main()`` (``image_translation.py:335-390``,
``distributed_image_translation.py:465-518``, ``angle_pairing.py:291-346``), so it is
restated here around the oracle's nets and loss helpers.  fp32, stock torch ops.
"""
from itertools import chain

import torch
import torch.nn as nn
import torch.optim as optim

from . import family
from .losses import get_fm_loss, get_fm_loss_angle, get_gan_loss


def synthetic_batch(batch, image_size, step=0, rank=0, device="cpu"):
    """A,B = uniform [0,1) fp32 from seed 1000*rank+step
    (precedent for random batches: batch_size_optimization.py:62-63)."""
    g = torch.Generator().manual_seed(1000 * rank + step)
    A = torch.rand(batch, 3, image_size, image_size, generator=g)
    B = torch.rand(batch, 3, image_size, image_size, generator=g)
    return A.to(device), B.to(device)


def build_nets(image_size, seed=1234, device="cpu", gen_cls=None, dis_cls=None):
    """Seed then construct G_A, G_B, D_A, D_B in that order
    (distributed_image_translation.py:372-376)."""
    gen_cls = gen_cls or (lambda: family.Generator(extra_layers=True, image_size=image_size))
    dis_cls = dis_cls or (lambda: family.Discriminator(image_size=image_size))
    torch.manual_seed(seed)
    G_A, G_B = gen_cls(), gen_cls()
    D_A, D_B = dis_cls(), dis_cls()
    return [n.to(device) for n in (G_A, G_B, D_A, D_B)]


class OracleStep:
    """One iteration of the reference loop; call ``step(A, B)`` repeatedly.

    variant: 'image_translation' (default rates 0.01/0.5) or 'angle_pairing'
    (FM skips feat 0; rates 0.9/0.9 -- angle_pairing.py:55-57,115).
    """

    def __init__(self, nets, model_arch="discogan", lr=2e-4, beta1=0.5, beta2=0.999,
                 update_interval=3, gan_curriculum=10000, starting_rate=None, default_rate=None,
                 variant="image_translation", device="cpu"):
        self.G_A, self.G_B, self.D_A, self.D_B = nets
        self.device = device
        self.model_arch = model_arch
        self.update_interval = update_interval
        self.gan_curriculum = gan_curriculum
        angle = variant == "angle_pairing"
        self.starting_rate = (0.9 if angle else 0.01) if starting_rate is None else starting_rate
        self.default_rate = (0.9 if angle else 0.5) if default_rate is None else default_rate
        self.fm = get_fm_loss_angle if angle else get_fm_loss
        self.recon_criterion = nn.MSELoss()                 # image_translation.py:267
        self.gan_criterion = nn.BCELoss()                   # :268
        self.feat_criterion = nn.HingeEmbeddingLoss()       # :269
        self.optim_gen = optim.Adam(chain(self.G_A.parameters(), self.G_B.parameters()),
                                    lr=lr, betas=(beta1, beta2), weight_decay=0.00001)   # :275-280
        self.optim_dis = optim.Adam(chain(self.D_A.parameters(), self.D_B.parameters()),
                                    lr=lr, betas=(beta1, beta2), weight_decay=0.00001)   # :282-287
        self.iters = 0

    def nets(self):
        return self.G_A, self.G_B, self.D_A, self.D_B

    def losses(self, A, B):
        """Forward graph of image_translation.py:342-382; returns dict of tensors."""
        G_A, G_B, D_A, D_B, dev = self.G_A, self.G_B, self.D_A, self.D_B, self.device
        AB = G_B(A)
        BA = G_A(B)
        ABA = G_A(AB)
        BAB = G_B(BA)
        recon_A = self.recon_criterion(ABA, A)
        recon_B = self.recon_criterion(BAB, B)
        A_real, A_feats_real = D_A(A)
        A_fake, A_feats_fake = D_A(BA)
        dis_A, gen_A = get_gan_loss(A_real, A_fake, self.gan_criterion, dev)
        fm_A = self.fm(A_feats_real, A_feats_fake, self.feat_criterion, dev)
        B_real, B_feats_real = D_B(B)
        B_fake, B_feats_fake = D_B(AB)
        dis_B, gen_B = get_gan_loss(B_real, B_fake, self.gan_criterion, dev)
        fm_B = self.fm(B_feats_real, B_feats_fake, self.feat_criterion, dev)
        rate = self.starting_rate if self.iters < self.gan_curriculum else self.default_rate
        gen_A_total = (fm_B * 0.9 + gen_B * 0.1) * (1 - rate) + recon_A * rate
        gen_B_total = (fm_A * 0.9 + gen_A * 0.1) * (1 - rate) + recon_B * rate
        if self.model_arch == "discogan":
            gen_loss = gen_A_total + gen_B_total
            dis_loss = dis_A + dis_B
        elif self.model_arch == "recongan":
            gen_loss = gen_A_total
            dis_loss = dis_B
        elif self.model_arch == "gan":
            gen_loss = gen_B * 0.1 + fm_B * 0.9
            dis_loss = dis_B
        else:
            raise ValueError(self.model_arch)
        return dict(gen_loss=gen_loss, dis_loss=dis_loss, gen_loss_A=gen_A, gen_loss_B=gen_B,
                    fm_loss_A=fm_A, fm_loss_B=fm_B, recon_loss_A=recon_A, recon_loss_B=recon_B,
                    dis_loss_A=dis_A, dis_loss_B=dis_B, AB=AB, BA=BA, ABA=ABA, BAB=BAB)

    def backward(self, A, B):
        """zero_grad + forward + backward of the loss this iteration steps on (image_translation.py:336-364,385-390
        without the optimiser call).  Returns the logged scalars."""
        for n in (self.G_A, self.G_B, self.D_A, self.D_B):
            n.zero_grad()
        out = self.losses(A, B)
        is_dis = self.iters % self.update_interval == 0
        (out["dis_loss"] if is_dis else out["gen_loss"]).backward()
        logged = {k: float(v.detach()) for k, v in out.items() if k.endswith(("_A", "_B")) and v.dim() == 0}
        logged["is_dis_step"] = is_dis
        return logged

    def apply(self):
        """optimizer.step() of the stepped group (image_translation.py:385-390) and the iteration counter."""
        is_dis = self.iters % self.update_interval == 0
        (self.optim_dis if is_dis else self.optim_gen).step()
        self.iters += 1

    def step(self, A, B, grad_hook=None):
        """image_translation.py:336-390.  ``grad_hook(nets)`` runs between backward and
        optimizer.step() (used to emulate the DDP gradient average)."""
        logged = self.backward(A, B)
        if grad_hook is not None:
            grad_hook((self.G_A, self.G_B, self.D_A, self.D_B))
        self.apply()
        return logged


class OracleDataParallel:
    """Single-process emulation of the reference's R-rank DDP step (distributed_image_translation.py:396-404,
    465-518 with the broadcast_buffers crash of SURVEY.md F4 avoided): R replicas start from the same weights, each
    runs forward/backward on its own shard with its own BatchNorm statistics and feature-matching means, parameter
    gradients are averaged over replicas (DDP's all-reduce mean), every replica takes the same Adam step."""

    def __init__(self, make_step, world):
        self.replicas = [make_step(r) for r in range(world)]
        ref = [list(n.parameters()) for n in self.replicas[0].nets()]
        for st in self.replicas[1:]:                    # DDP constructor: rank 0's weights everywhere
            for net, ps in zip(st.nets(), ref):
                for p, q in zip(net.parameters(), ps):
                    p.data.copy_(q.data)

    def step(self, shards):
        """shards[r] = (A_r, B_r).  Returns the per-replica logged scalars."""
        logs = [st.backward(A, B) for st, (A, B) in zip(self.replicas, shards)]
        R = len(self.replicas)
        plists = [[p for n in st.nets() for p in n.parameters()] for st in self.replicas]
        for ps in zip(*plists):
            if ps[0].grad is None:
                continue
            mean = sum(p.grad for p in ps) / R
            for p in ps:
                p.grad = mean.clone()
        for st in self.replicas:
            st.apply()
        return logs
