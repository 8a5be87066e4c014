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
            ("C5x1", 10, 4, 16, 3, 1048576), ("K12", 10, 4, 12, 3, 262144)]
    # (variant, ring depth, CTA cap [9 = one CTA per tile], stagger us, dynamic tiles)
    # ring depth "2" selects the ring-depth-3 kernel WITHOUT the hoisted outward-sweep work (A/B of the hoist)
    variants = [(3, 0, 0, 0, 0), (4, 3, 0, 0, 1), (4, 3, 0, 0, 2), (6, 3, 0, 0, 1), (6, 3, 0, 0, 2)]
    rows = []
    if "--large-only" in sys.argv:
        cfgs = []
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
    # ---- large K: the chunked kernel (default routing), and its recompute overhead on the headline shape
    s.set_option(OPT_DYN, 0)
    for name, N, r, K, D, B, variant, chunk in (("K50", 10, 4, 50, 3, 65536, 0, 0), ("K50c4", 10, 4, 50, 3, 65536, 5, 4), ("K50c3", 10, 4, 50, 3, 65536, 5, 3), ("K100c4", 10, 4, 100, 3, 32768, 5, 4),
                                                ("K100", 10, 4, 100, 3, 32768, 0, 0), ("K34", 10, 4, 34, 3, 65536, 0, 0),
                                                ("K36", 10, 4, 36, 3, 65536, 0, 0),
                                                ("C3chunk7", 10, 4, 16, 3, 262144, 5, 0), ("C3chunk4", 10, 4, 16, 3, 262144, 5, 4)):
        prob = m.Problem(N, r, K, D)
        times, dfix = synth(N, K, D, B, dev)
        out = torch.zeros((B, K, D, N), device=dev, dtype=torch.float64)
        st = torch.full((B,), -1, dtype=torch.int32, device=dev)
        s.set_option(OPT_VARIANT, variant)
        s.set_option(6, chunk)
        try:
            for _ in range(3):
                s.solve_linear(prob, times, dfix, coeffs=out, status=st)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
            ev[0].record()
            for i in range(10):
                s.solve_linear(prob, times, dfix, coeffs=out)
                ev[i + 1].record()
            torch.cuda.synchronize()
            ms = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
            med = ms[len(ms) // 2]
            gbs = B / (med * 1e-3) * prob.bytes_per_trajectory / 1e9
            row = dict(cfg=name, variant=variant, chunk=chunk, ms=round(med, 4), traj_per_s=round(B / (med * 1e-3)),
                       frac_hbm=round(gbs / 6575.4, 4), status_ok=bool((st == 0).all().item()),
                       finite=bool(torch.isfinite(out).all().item()))
        except Exception as e:
            row = dict(cfg=name, variant=variant, chunk=chunk, error=str(e))
        rows.append(row)
        print(json.dumps(row))
    s.set_option(6, 0)
    s.set_option(OPT_VARIANT, 0)
    # ---- arbitrary masks: masked block kernel (0) vs the banded global-scratch kernel (1)
    import numpy as np
    for name, N, r, K, D, B in (("gmask16", 10, 4, 16, 3, 65536), ("gmask6_D5", 10, 4, 6, 5, 65536), ("gmaskN12", 12, 5, 8, 3, 32768)):
        h = N // 2
        mask = np.zeros((K + 1, h), dtype=np.uint8)
        mask[:, :2] = 1
        mask[0, :] = 1
        mask[-1, :] = 1
        prob = m.Problem(N, r, K, D, fixed_mask=mask)
        g = torch.Generator(device=dev)
        g.manual_seed(5)
        times = torch.rand((B, K), generator=g, device=dev, dtype=torch.float64) * 4 + 2
        dfix = torch.rand((B, D, prob.n_fixed), generator=g, device=dev, dtype=torch.float64) * 4 - 2
        ref = None
        for gv in (1, 0):
            s.set_option(7, gv)
            out = torch.zeros((B, K, D, N), device=dev, dtype=torch.float64)
            st = torch.full((B,), -1, dtype=torch.int32, device=dev)
            try:
                for _ in range(2):
                    s.solve_linear(prob, times, dfix, coeffs=out, status=st)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    s.solve_linear(prob, times, dfix, coeffs=out)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 5
                gbs = B / (ms * 1e-3) * prob.bytes_per_trajectory / 1e9
                diff = 0.0
                if ref is None:
                    ref = out.clone()
                else:
                    den = ref.abs().reshape(B, -1).max(dim=1).values
                    diff = float(((out - ref).abs().reshape(B, -1).max(dim=1).values / den).max())
                row = dict(cfg=name, generic_variant=gv, ms=round(ms, 4), traj_per_s=round(B / (ms * 1e-3)),
                           frac_hbm=round(gbs / 6575.4, 4), status_ok=bool((st == 0).all().item()),
                           max_rel_diff_vs_banded=diff)
            except Exception as e:
                row = dict(cfg=name, generic_variant=gv, error=str(e))
            rows.append(row)
            print(json.dumps(row))
    s.set_option(7, 0)
    s.set_option(OPT_STAG, 0)
    s.set_option(OPT_DYN, 0)
    s.set_option(OPT_CTAS, 0)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/k1_variants.json", "w"), indent=1)


if __name__ == "__main__":
    main()
