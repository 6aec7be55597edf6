"""Markdown summary of gpurun_out/parity_report.jsonl (written by tests/test_parity_gpu.py):
    python tools/parity_summary.py gpurun_out/parity_report.jsonl > profiles/r02_parity.md"""
import json
import statistics
import sys

L = ("dis_loss_A", "gen_loss_A", "dis_loss_B", "gen_loss_B", "fm_loss_A", "fm_loss_B", "recon_loss_A", "recon_loss_B")


def main(path):
    recs = [json.loads(l) for l in open(path)]
    out = ["# Parity report (round 2) — measured by `tests/test_parity_gpu.py` on a B200", "",
           "Oracles run on the same device with TF32 off: `fp32` = the reference restated in stock PyTorch; `bf16e` = the same "
           "with bf16 rounding at the kernels' storage points; `floor` = distance between `bf16e` and its 1e-6-perturbed twin "
           "(two correct bf16 computations with different fp32 summation order).  rel = relative L2.", ""]
    for r in recs:
        if r["test"] == "layers_teacher_forced":
            out += [f"## Teacher-forced per-layer parity, {r['S']}x{r['S']} B={r['B']} (each layer fed the emulation's own tensors)", "",
                    "| net | layer | kind | input | fprop | dgrad | wgrad | BN fwd | BN dz | dgamma | dbeta | fused stats: mean err (sigma) / invstd rel |",
                    "|---|---|---|---|---|---|---|---|---|---|---|---|"]
            for x in r["rows"]:
                fs = f"{x['fused_mean_err_sigma']:.1e} / {x['fused_invstd_rel']:.1e}" if x.get("stats_fused") else "split-K (separate pass)"  # noqa: E501
                out.append(f"| {x['net']} | {x['layer']} | {x['kind']} | {'x'.join(map(str, x['x']))} | {x['fprop']:.2e} | {x['dgrad']:.2e} | "
                           f"{x['wgrad']:.1e} | {x['bn_fwd']:.2e} | {x['bn_bwd_dz']:.2e} | {x['bn_dgamma']:.1e} | {x['bn_dbeta']:.1e} | {fs} |")
            out.append("")
        if r["test"] == "per_layer_gradient_table":
            rows = r["rows"]
            out += [f"## End-to-end gradients after one backward, {r['S']}x{r['S']} B={r['B']}, {r['kind']} step (identical weights)", "",
                    f"median rel(kernel,bf16e)/floor = {statistics.median(x['rel_bf16e'] / x['floor'] for x in rows):.3f}, "
                    f"max = {max(x['rel_bf16e'] / x['floor'] for x in rows):.3f}; "
                    f"max rel(kernel,fp32)/rel(bf16e,fp32) = {max(x['rel_fp32'] / x['rel_bf16e_vs_fp32'] for x in rows):.3f}", "",
                    "| net | parameter | rel(kernel, bf16e) | floor | rel(kernel, fp32) | rel(bf16e, fp32) | cos(kernel, fp32) |", "|---|---|---|---|---|---|---|"]
            for x in rows:
                out.append(f"| {x['net']} | {x['param']} | {x['rel_bf16e']:.4f} | {x['floor']:.4f} | {x['rel_fp32']:.4f} | "
                           f"{x['rel_bf16e_vs_fp32']:.4f} | {x['cos_fp32']:.4f} |")
            out.append("")
        if r["test"] == "step_parity":
            c = r["curves"]
            out += [f"## Losses over the first steps, {r['S']}x{r['S']} B={r['B']}, {r['variant']} / {r['arch']} (kernel / bf16e / fp32)", "",
                    "| it | " + " | ".join(L) + " |", "|---|" + "---|" * len(L)]
            for it in range(len(c["kernel"])):
                out.append(f"| {it} | " + " | ".join(f"{c['kernel'][it][k]:.4f} / {c['bf16e'][it][k]:.4f} / {c['fp32'][it][k]:.4f}" for k in L) + " |")
            out.append("")
        if r["test"] == "accumulated_update":
            rows = [x for x in r["rows"] if x.get("floor")]
            if not rows:
                out += [f"Accumulated update after {r['steps']} steps ({r['S']}x{r['S']}): cos(kernel,bf16e) min "
                        f"{min(x['cos_bf16e'] for x in r['rows']):.3f}, cos(kernel,fp32) min {min(x['cos_fp32'] for x in r['rows']):.3f}", ""]
                continue
            out += [f"Accumulated update after {r['steps']} steps ({r['S']}x{r['S']}, {r['variant']} / {r['arch']}): rel(kernel,bf16e)/floor median "
                    f"{statistics.median(x['rel_bf16e'] / x['floor'] for x in rows):.3f}, max {max(x['rel_bf16e'] / x['floor'] for x in rows):.3f}; "
                    f"cos(kernel,fp32) min {min(x['cos_fp32'] for x in rows):.3f}", ""]
        if r["test"] == "loss_curve_100_summary":
            out += ["## 100-step loss curves, 64x64 B=64 (deviation relative to the loss's mean over the run)", "",
                    "| loss | mean (fp32) | last-50 mean: kernel vs fp32 | bf16e vs fp32 | worst step: kernel vs fp32 | bf16e vs fp32 |", "|---|---|---|---|---|---|"]
            for k, s in r["summary"].items():
                out.append(f"| {k} | {s['mean_fp32']:.4f} | {s['tail_kernel_vs_fp32']:.4f} | {s['tail_bf16e_vs_fp32']:.4f} | "
                           f"{s['max_kernel_vs_fp32']:.3f} | {s['max_bf16e_vs_fp32']:.3f} |")
            out.append("")
        if r["test"] == "dp_semantics_one_gpu":
            g = r["grad_rows"]
            out += ["## Data-parallel semantics, two in-process ranks vs the oracle's R=2 DDP emulation (64x64, B=32 per rank)", "",
                    f"step-0 averaged gradients ({len(g)} parameters): norm ratio {min(x['norm_ratio'] for x in g):.4f} .. "
                    f"{max(x['norm_ratio'] for x in g):.4f} (a sum instead of a mean would read 2.0); "
                    f"max rel(kernel,fp32)/rel(bf16e,fp32) = {max(x['rel_fp32'] / x['rel_bf16e_vs_fp32'] for x in g):.3f}; "
                    f"per-rank losses within 2 % + 0.01 of bf16e over {len(r['rows']) // 16} steps.", ""]
    print("\n".join(out))


if __name__ == "__main__":
    main(sys.argv[1])
