"""Developer experiment: time the headline kernel variants (v3 = per-tile CTAs, v4 = persistent with deep
prefetch, ring depth 2..4, CTAs/SM cap) on C3 / C2 / C4 / C5-on-one-GPU and check v4 against v3."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import mav_trajectory_generation_b200 as m
from tools.quick_bench import synth

OPT_VARIANT, OPT_RD, OPT_CTAS, OPT_STAG, OPT_DYN = 1, 2, 3, 4, 5


def main():
    dev = torch.device("cuda:0")
    s = m.Solver(0)
    cfgs = [("C3", 10, 4, 16, 3, 262144), ("C2", 10, 4, 8, 3, 65536), ("C4", 8, 3, 4, 3, 1048576),
            ("C5x1", 10, 4, 16, 3, 1048576), ("odd", 10, 4, 7, 3, 100001)]
    # (variant, ring depth, CTA cap [9 = one CTA per tile], stagger us, dynamic tiles)
    variants = [(3, 0, 0, 0, 0), (4, 3, 0, 0, 0), (4, 3, 9, 0, 0), (4, 3, 0, 0, 1), (4, 3, 0, 8, 1), (4, 3, 0, 16, 1),
                (4, 3, 0, 24, 1), (4, 3, 0, 16, 0), (4, 2, 0, 16, 1)]
    rows = []
    for name, N, r, K, D, B in cfgs:
        prob = m.Problem(N, r, K, D)
        times, dfix = synth(N, K, D, B, dev)
        ref = None
        for variant, rd, cap, stag, dyn in variants:
            s.set_option(OPT_VARIANT, variant)
            if rd:
                s.set_option(OPT_RD, rd)
            s.set_option(OPT_CTAS, cap)
            s.set_option(OPT_STAG, stag)
            s.set_option(OPT_DYN, dyn)
            out = torch.zeros((B, K, D, N), device=dev, dtype=torch.float64)
            st = torch.full((B,), -1, dtype=torch.int32, device=dev)
            dfree = torch.zeros((B, D, prob.n_free), device=dev, dtype=torch.float64)
            try:
                for _ in range(3):
                    s.solve_linear(prob, times, dfix, coeffs=out, status=st, d_free=dfree)
                torch.cuda.synchronize()
            except Exception as e:
                print(json.dumps(dict(cfg=name, variant=variant, rd=rd, cap=cap, stag=stag, dyn=dyn, error=str(e))))
                continue
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
            ev[0].record()
            for i in range(20):
                s.solve_linear(prob, times, dfix, coeffs=out)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
            med = ms[len(ms) // 2]
            rate = B / (med * 1e-3)
            gbs = rate * prob.bytes_per_trajectory / 1e9
            if ref is None:
                ref = (out.clone(), dfree.clone())
                diff = dfd = 0.0
            else:
                den = ref[0].abs().reshape(B, -1).max(dim=1).values
                diff = float(((out - ref[0]).abs().reshape(B, -1).max(dim=1).values / den).max())
                dfd = float((dfree - ref[1]).abs().max() / ref[1].abs().max())
            row = dict(cfg=name, variant=variant, rd=rd, cap=cap, stag=stag, dyn=dyn, ms=round(med, 4), min_ms=round(ms[0], 4),
                       traj_per_s=round(rate), frac_hbm=round(gbs / 6575.4, 4), status_ok=bool((st == 0).all().item()),
                       finite=bool(torch.isfinite(out).all().item()), max_rel_diff_vs_v3=diff, dfree_diff=dfd)
            rows.append(row)
            print(json.dumps(row))
    s.set_option(OPT_VARIANT, 0)
    s.set_option(OPT_STAG, 0)
    s.set_option(OPT_DYN, 0)
    s.set_option(OPT_CTAS, 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/k1_variants.json", "w"), indent=1)


if __name__ == "__main__":
    main()
