"""Per-shape roofline table of the convolution calls of one train256 iteration, from tools/call_times.py output:
measured device time per call (taken BEFORE the second MMA-issuing warp: the halo-kernel rows are ~12 % faster now)
against (1) the HBM bound (activations read once + written once, 2 bytes each, at the
measured copy bandwidth), (2) the tensor bound (2*9*Cin*Cout flop per pixel at the sustained bf16 peak) and (3) the
shared-memory operand bound of the 1-CTA tcgen05.mma formulation used by conv_halo.cu (profiles/r1_mma_issue_probe.txt:
max(32 + N/4, N/2) clk per M=128, K=16 instruction, N = min(Cout, 128)).

    python tools/layer_rooflines.py profiles/r1_call_times_train256_b32.txt > profiles/r1_layer_rooflines_train256.txt
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EVENT_US = 5.0          # what the event pair around a call adds (the smallest calls of the file measure 6.4 us)
SMS, CLK = 148, 1.965e9


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p.get("hbm_gbs_sustained", p.get("hbm_gbs", 7000.0))), float(p.get("bf16_tflops_sustained", 1389.2))
    except Exception:
        return 7000.0, 1389.2


def main(path):
    hbm_gbs, tf = peaks()
    rows, name = [], None
    for line in open(path):
        m = re.match(r"== (\w+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+(\d+) x\s+([\d.]+) us\s+\((.*)\)", line)
        if not m or name not in ("bg_conv_fprop", "bg_conv_fprop_stats", "bg_conv_style_fprop", "bg_conv_pool4_fprop",
                                 "bg_conv_pool4_dgrad", "bg_conv_pool_fprop"):
            continue
        cnt, us, a = int(m.group(1)), float(m.group(2)) - EVENT_US, [int(v) for v in m.group(3).split(",") if v.strip()]
        n, h, w, ci, co = a[:5]
        hin, win, hout, wout, macs_px = h, w, h, w, 9
        kind = name.replace("bg_conv_", "")
        if name == "bg_conv_pool4_fprop":          # args are the FULL-resolution input; output pooled; 16 taps per pooled pixel
            hout, wout = h // 2, w // 2
        elif name == "bg_conv_pool_fprop":
            hout, wout = h // 2, w // 2
        elif name == "bg_conv_pool4_dgrad":        # args: pooled gradient map (N, Hp, Wp, Cout) -> full map (Cin)
            hin, win, hout, wout = h, w, 2 * h, 2 * w
            ci, co = a[3], a[4]
        elif name == "bg_conv_style_fprop" and a[5] == 1:   # upsample fused: input is the low-resolution map
            hin, win = h // 2, w // 2
        full_px = n * max(hin * win, hout * wout)
        flops = 2.0 * full_px * 9 * ci * co          # the reference's 3x3 formulation
        byts = 2.0 * n * (hin * win * ci + hout * wout * co)
        t_hbm = byts / (hbm_gbs * 1e9) * 1e6
        t_ref = flops / (tf * 1e12) * 1e6
        # conv+pool folded into one 4x4 stride-2 conv executes 16 MACs per pooled pixel = 4/9 of the reference's count
        t_tc = t_ref * (4.0 / 9.0 if "pool4" in name else 1.0)
        # instructions actually issued by the halo kernel: per 128 output pixels, per tap (16 taps per pooled pixel = 4 per
        # full-resolution pixel for the folded conv+pool), per 16 input channels, per n-block
        nb = min(co, 128)
        taps = 9 if "pool4" not in name else (16 if name == "bg_conv_pool4_fprop" else 4)
        out_px = n * hout * wout if name != "bg_conv_pool4_dgrad" else n * hout * wout
        instr = out_px / 128.0 * taps * (ci / 16.0) * (co / nb)
        t_smem = instr * max(32 + nb / 4.0, nb / 2.0) / SMS / CLK * 1e6 if min(hout, wout) >= 16 else float("nan")
        bound = max(t_hbm, t_tc, t_smem if t_smem == t_smem else 0.0)
        rows.append((cnt * us, kind, (n, h, w, a[3], a[4]), cnt, us, t_hbm, t_tc, t_ref, t_smem, bound / us))
    rows.sort(reverse=True)
    print(f"HBM {hbm_gbs:.0f} GB/s, tensor {tf:.0f} TFLOP/s (MEASURED_PEAKS.json or the stated fallback); times in us per call, event pair "
          f"({EVENT_US} us) subtracted; 'frac' = largest bound / measured")
    print("tensor = executed MACs at the sustained peak; ref = the reference's conv-then-pool count (differs for the folded conv+pool)")
    print(f"{'call':18s} {'(N, H, W, C, C)':28s} {'n':>3s} {'meas':>8s} {'hbm':>7s} {'tensor':>7s} {'ref':>7s} {'smem-op':>8s} {'frac':>6s}")
    tot_meas = tot_bound = 0.0
    for _, kind, shape, cnt, us, t_hbm, t_tc, t_ref, t_smem, frac in rows:
        print(f"{kind:18s} {str(shape):28s} {cnt:3d} {us:8.1f} {t_hbm:7.1f} {t_tc:7.1f} {t_ref:7.1f} {t_smem:8.1f} {frac:6.2f}")
        tot_meas += cnt * us
        tot_bound += cnt * us * frac
    print(f"total measured {tot_meas / 1e3:.2f} ms, sum of per-call bounds {tot_bound / 1e3:.2f} ms ({tot_bound / tot_meas:.2f})")


if __name__ == "__main__":
    main(sys.argv[1])
