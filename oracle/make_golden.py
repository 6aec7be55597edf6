"""Mint golden vectors from the reference's own Python (oracle tooling; test infrastructure).

Run in the build container only (needs /root/reference, which does not travel to the
GPU box):   python -m oracle.make_golden

Writes small fixtures to tests/golden/:
  ref512_structure.json  state-dict keys/shapes/param counts of reference model.py
  ref512_forward.pt      reference G and D forward at 512^2, B=2 (seeded), sampled outputs
  ref_losses.pt          reference get_gan_loss / get_fm_loss (both variants) on seeded inputs
  ref512_step.pt         2 iterations (D step, G step) of the restated loop driven with the
                         REFERENCE model classes and REFERENCE loss helpers, B=2
  family64_step.pt       3 iterations of the oracle family at 64^2, B=8 (self-minted: the
                         reference cannot run below 512^2)
"""
import importlib
import json
import sys
import types
from pathlib import Path

import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def import_reference():
    """Import model.py and image_translation.py from the reference (matplotlib stubbed)."""
    if not REF.exists():
        raise RuntimeError("reference tree not present; goldens can only be minted in the build container")
    sys.path.insert(0, str(REF))
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    ref_model = importlib.import_module("model")
    ref_it = importlib.import_module("image_translation")
    # angle_pairing.py does not import as shipped: inject the missing names
    ds = importlib.import_module("dataset")
    for n in ("get_cars", "get_chairs", "get_faces_3d"):
        if not hasattr(ds, n):
            setattr(ds, n, lambda *a, **k: None)
    ref_ap = importlib.import_module("angle_pairing")
    return ref_model, ref_it, ref_ap


def sample_idx(numel, n=256, seed=7):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (n,), generator=g)


def main():
    from oracle.step import OracleStep, build_nets, synthetic_batch
    torch.set_num_threads(8)
    OUT.mkdir(parents=True, exist_ok=True)
    ref_model, ref_it, ref_ap = import_reference()

    # ---- structure ------------------------------------------------------------------
    torch.manual_seed(1234)
    G = ref_model.Generator(extra_layers=True)
    D = ref_model.Discriminator()
    structure = {
        "generator": {k: list(v.shape) for k, v in G.state_dict().items()},
        "discriminator": {k: list(v.shape) for k, v in D.state_dict().items()},
        "generator_params": sum(p.numel() for p in G.parameters()),
        "discriminator_params": sum(p.numel() for p in D.parameters()),
        "generator_param_order": [n for n, _ in G.named_parameters()],
        "discriminator_param_order": [n for n, _ in D.named_parameters()],
    }
    (OUT / "ref512_structure.json").write_text(json.dumps(structure, indent=0))

    # ---- forward at 512^2 -----------------------------------------------------------
    A, B = synthetic_batch(2, 512, step=0)
    G.train(); D.train()
    with torch.no_grad():
        y = G(A)
        p, feats = D(B)
    fwd = {
        "seed": 1234, "batch": 2,
        "G_out_idx": sample_idx(y.numel()), "D_prob": p.flatten().clone(),
        "G_out_mean": y.mean(), "G_out_std": y.std(),
        "feat_means": torch.stack([f.mean() for f in feats]),
        "feat_stds": torch.stack([f.std() for f in feats]),
        "feat_shapes": [list(f.shape) for f in feats],
    }
    fwd["G_out_samples"] = y.flatten()[fwd["G_out_idx"]].clone()
    fwd["G_running_mean_3"] = G.state_dict()["encoder.3.running_mean"].clone()
    G.eval()
    with torch.no_grad():
        ye = G(A)
    fwd["G_eval_samples"] = ye.flatten()[fwd["G_out_idx"]].clone()
    torch.save(fwd, OUT / "ref512_forward.pt")
    del G, D

    # ---- loss helpers ---------------------------------------------------------------
    g = torch.Generator().manual_seed(99)
    dr = torch.rand(6, 1, 1, 1, generator=g) * 0.98 + 0.01
    df = torch.rand(6, 1, 1, 1, generator=g) * 0.98 + 0.01
    df[0] = 0.0  # exercises the -100 log clamp
    rf = [torch.randn(6, c, s, s, generator=g) for c, s in ((8, 8), (16, 4), (32, 2))]
    ff = [torch.randn(6, c, s, s, generator=g) for c, s in ((8, 8), (16, 4), (32, 2))]
    bce, hinge = torch.nn.BCELoss(), torch.nn.HingeEmbeddingLoss()
    dl, gl = ref_it.get_gan_loss(dr, df, bce, "cpu")
    losses = {"dr": dr, "df": df, "rf": rf, "ff": ff, "dis_loss": dl, "gen_loss": gl,
              "fm": ref_it.get_fm_loss(rf, ff, hinge, "cpu"),
              "fm_angle": ref_ap.get_fm_loss(rf, ff, hinge, "cpu")}
    torch.save(losses, OUT / "ref_losses.pt")

    # ---- restated step driven with reference classes + reference loss helpers ---------
    import oracle.step as ostep
    nets = build_nets(512, gen_cls=lambda: ref_model.Generator(extra_layers=True),
                      dis_cls=lambda: ref_model.Discriminator())
    saved = (ostep.get_gan_loss, ostep.get_fm_loss)
    ostep.get_gan_loss, ostep.get_fm_loss = ref_it.get_gan_loss, ref_it.get_fm_loss
    try:
        st = OracleStep(nets)
        st.fm = ref_it.get_fm_loss
        logs = []
        for it in range(2):
            A, B = synthetic_batch(2, 512, step=it)
            logs.append(st.step(A, B))
    finally:
        ostep.get_gan_loss, ostep.get_fm_loss = saved
    w = nets[2].state_dict()["conv2.weight"]
    torch.save({"logs": logs, "batch": 2,
                "D_A_conv2_idx": sample_idx(w.numel()),
                "D_A_conv2_samples": w.flatten()[sample_idx(w.numel())].clone(),
                "G_A_enc0_samples": nets[0].state_dict()["encoder.0.weight"].flatten()[:64].clone()},
               OUT / "ref512_step.pt")
    del nets, st

    # ---- family at 64^2 (self-minted) -----------------------------------------------
    for variant, arch in (("image_translation", "discogan"), ("angle_pairing", "discogan"),
                          ("image_translation", "recongan"), ("image_translation", "gan")):
        nets = build_nets(64)
        st = OracleStep(nets, model_arch=arch, variant=variant)
        logs = []
        for it in range(3):
            A, B = synthetic_batch(8, 64, step=it)
            logs.append(st.step(A, B))
        torch.save({"logs": logs, "batch": 8,
                    "G_A_enc0": nets[0].state_dict()["encoder.0.weight"].flatten()[:64].clone(),
                    "D_B_conv1": nets[3].state_dict()["conv1.weight"].flatten()[:64].clone(),
                    "G_B_rm": nets[1].state_dict()["encoder.3.running_mean"][:16].clone()},
                   OUT / f"family64_step_{variant}_{arch}.pt")
    print("goldens written to", OUT)


if __name__ == "__main__":
    main()
