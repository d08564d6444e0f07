#!/usr/bin/env python
"""Benchmark of the BYO-GAN hot path on B200: G+D training iterations at 256x256 (BASELINE.json configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload train256|...]

One "step" is one full reference iteration (train.py:135-219): critic step with the R1 gradient penalty
(double-backward) and its Adam update, then generator step and its Adam update, driven through the public
drop-in API of byo-gan_b200/gan.py.  Prints ONE JSON line (see the task contract): `value` is img/s with inputs
resident in HBM, `e2e` the same iteration fed from pinned HOST buffers with the losses read back every step,
`roofline` the tcgen05 implicit-GEMM conv kernel against the measured bf16 peak, `cpu_baseline` the oracle
port of the reference timed on this box's host cores on a bounded sample.

--impl reference runs only that CPU leg (the reference is pure PyTorch; /root/reference is not on the GPU box,
so its arithmetic is timed through oracle/gan_oracle.py, which is pinned to the real gan.py by tests/golden).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "byo-gan_b200"))
sys.path.insert(0, ROOT)

# the step allocates and frees GB-sized activations in a pattern that takes the caching allocator several iterations
# to settle with fixed-size segments (cudaMalloc stalls inside the first timed steps); expandable segments settle at once
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import torch  # noqa: E402

# step FLOPs per image in the reference's formulation: 4*G_fwd + 11*D_fwd
WORKLOADS = {
    # name: (steps, alpha, per-GPU batch, GFLOP/img/iteration, description)
    "train4": (1, None, 16, 1.258, "4x4 stage, batch 16 (BASELINE configs[0], the reference's CPU-runnable case)"),
    "train64": (5, 0.5, 64, 235.09, "64x64 stage with fade-in alpha=0.5, batch 64 (BASELINE configs[1])"),
    "train256": (7, None, 32, 423.65, "256x256 stage with style mixing, batch 32 per GPU, R1 every step (BASELINE configs[2])"),
    "train512": (8, None, 16, 518.06, "512x512 stage, batch 16 per GPU, R1 every step (BASELINE configs[3])"),
}
BOUND = {"train4": "tensor", "train64": "tensor", "train256": "tensor", "train512": "hbm"}
STYLE_MIXING_DEFAULT = {"train256"}     # BASELINE configs[2] is "256x256 stage WITH style mixing"
LAMBDA = 10.0          # config.txt gradient_lambda
LR, BETAS = 0.002, (0.0, 0.99)   # config.txt lr / beta_1 / beta_2
# train.py:76-80 builds plain torch.optim.Adam; fused=True is the same update in one multi-tensor kernel per optimizer
# (BG_FUSED_ADAM=0 restores the for-each implementation, ~40 small launches per iteration more)
FUSED_ADAM = os.environ.get("BG_FUSED_ADAM", "1") != "0"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "src": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe).  Sampled through NVML from a
    thread of this process every 40 ms — the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*`
    prints; a polling nvidia-smi PROCESS was observed to stall kernel launches for hundreds of milliseconds on these
    hosts (timed regions 60 % long while it ran, never after it was stopped).  Falls back to nvidia-smi at a 500 ms
    period if pynvml is unavailable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self._stop, self._thread = index, [], None, False, None

    def _nvml_loop(self, nv, h):
        names = [("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)]
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.rows.append((time.time(), [str(sm), str(mx)] + ["Active" if mask & bit else "Not Active" for _, bit in names]))
            except Exception:
                pass
            time.sleep(0.04)

    def start(self):
        """Started BEFORE the warm-up; only rows that arrive inside the window() are reported."""
        try:
            import pynvml as nv

            nv.nvmlInit()
            # the process may see a subset of the GPUs (CUDA_VISIBLE_DEVICES): map through the PCI bus id
            bus = torch.cuda.get_device_properties(self.index).pci_bus_id if hasattr(
                torch.cuda.get_device_properties(self.index), "pci_bus_id") else None
            h = None
            if bus is not None:
                for i in range(nv.nvmlDeviceGetCount()):
                    hi = nv.nvmlDeviceGetHandleByIndex(i)
                    if nv.nvmlDeviceGetPciInfo(hi).bus == bus:
                        h = hi
                        break
            if h is None:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self._thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self._thread.start()
            return
        except Exception:
            self._thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "500", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        self._stop = True
        if self._thread is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        if self.proc is not None:
            self.proc.terminate()
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        rows = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.05]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": "nvml" if self._thread is not None else "nvidia-smi"}


# ------------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------------
def make_trainer(steps, alpha, batch, device, style_mixing, graph=False):
    """The package's Trainer (byo-gan_b200/trainer.py): train.py:58-80 + one iteration of train.py:135-219.
    graph: replay the iteration from CUDA graphs (single process only; Adam with device-side step counters)."""
    import trainer

    tr = trainer.Trainer(steps, alpha, batch, device, lr=LR, betas=BETAS, c_lambda=LAMBDA, fused_adam=FUSED_ADAM,
                         style_mixing=style_mixing, perturb_init=True, capturable=graph)
    if graph:
        tr.enable_graphs()
    return tr


def conv_bytes(name, args):
    """Algorithmic HBM bytes of one conv launch (bf16 maps read once and written once; weights excluded):
    2 * N * (Cin * Hin * Win + Cout * Hout * Wout); + the gate map for a gated pass is not counted."""
    n, h, w, ci, co = args[:5]
    if name == "bg_conv_pool4_dgrad":                      # reads the pooled gradient, writes (and gates) the full map
        return 2.0 * n * (ci * h * w + co * 4 * h * w)
    hin = win = None
    if name == "bg_conv_style_fprop" and args[5]:          # upsample flag: the input is the quarter-size map
        hin, win = h // 2, w // 2
    hin, win = hin or h, win or w
    ho, wo = (h // 2, w // 2) if name in ("bg_conv_pool_fprop", "bg_conv_pool4_fprop") else (h, w)
    return 2.0 * n * (ci * hin * win + co * ho * wo)


def conv_flops(args):
    # bg_conv_fprop scalar args: N,H,W,Cin,Cout,ksize,act,slope ; bg_conv_wgrad: N,H,W,Cin,Cout,accumulate
    n, h, w, ci, co = args[:5]
    ks = args[5] if len(args) >= 8 else 3
    return 2.0 * n * h * w * ks * ks * ci * co


FPROP_CALLS = ("bg_conv_fprop", "bg_conv_fprop_stats", "bg_conv_pool_fprop", "bg_conv_pool4_fprop", "bg_conv_pool4_dgrad",
               "bg_conv_style_fprop")
WGRAD_CALLS = ("bg_conv_wgrad", "bg_conv_pool4_wgrad")


def conv_call_work(name, a):
    """(credited flops, executed flops, algorithmic bytes) of one conv C-ABI call from its scalar arguments.
    Credited = the reference formulation (a 3x3 conv at full resolution); executed = what the kernel
    really multiplies: the folded conv+pool calls run 16 taps per POOLED pixel = 16/36 of the credited MACs."""
    if name == "bg_conv_pool4_wgrad":          # N, Hp, Wp, Cin, Cout, accumulate
        fl = conv_flops((a[0], 2 * a[1], 2 * a[2], a[3], a[4], 0))
        return fl, fl * 16.0 / 36.0, 2.0 * a[0] * (a[3] * 4 * a[1] * a[2] + a[4] * a[1] * a[2])
    if name == "bg_conv_pool4_dgrad":          # N, Hp, Wp, Cout, Cin -> the 3x3 dgrad at full resolution
        fl = conv_flops((a[0], 2 * a[1], 2 * a[2], a[3], a[4], 3, 0, 0.0))
        return fl, fl * 16.0 / 36.0, conv_bytes(name, a)
    if name == "bg_conv_wgrad":
        fl = conv_flops(a)
        return fl, fl, 2.0 * a[0] * a[1] * a[2] * (a[3] + a[4])
    fl = conv_flops(a if name not in ("bg_conv_pool_fprop", "bg_conv_pool4_fprop", "bg_conv_style_fprop") else a[:5] + (3, 0, 0.0))
    return fl, (fl * 16.0 / 36.0 if name == "bg_conv_pool4_fprop" else fl), conv_bytes(name, a)


def aux_call_bytes(name, a):
    """Algorithmic HBM bytes of the fused norm / noise / resample / 1x1 helpers (one read of every input map, one write of
    every output map; bf16 feature maps, fp32 image planes), or None for calls that are not map-sized streams."""
    try:
        if name in ("bg_adain_apply",):                       # N, HW, C, eps: read a, write x
            return 2.0 * 2 * a[0] * a[1] * a[2]
        if name == "bg_adain_bwd_reduce":                     # N, HW, C: read g, a
            return 2.0 * 2 * a[0] * a[1] * a[2]
        if name == "bg_adain_bwd_apply":                      # read g, a; write gpre
            return 2.0 * 3 * a[0] * a[1] * a[2]
        if name == "bg_in_stats":
            return 2.0 * a[0] * a[1] * a[2]
        if name in ("bg_upsample2x_fwd", "bg_upsample2x_bwd"):   # N, H, W, C of the SMALL map: small + 4x large
            return 2.0 * 5 * a[0] * a[1] * a[2] * a[3]
        if name == "bg_planes3_to_nhwc":                      # P, HW, C, ...: 3 fp32 planes in (or C planes), C bf16 out
            return a[0] * (12.0 + 2.0 * a[2])
        if name == "bg_nhwc_to_planes3":
            return a[0] * (12.0 + 2.0 * a[2])
        if name == "bg_to_rgb_adain":                         # N, HW, C
            return a[0] * a[1] * (12.0 + 2.0 * a[2])
        if name == "bg_channel_wsum":                         # P, C, HW, img_stride, plane_stride, nplanes
            return a[0] * (2.0 * a[1] + 4.0 * a[5])
        if name in ("bg_act_gate", "bg_axpby"):               # n elements: two maps in (gate: g, y), one out
            return 2.0 * 3 * a[0]
        if name == "bg_axpby_f32":
            return 4.0 * 3 * a[0]
        if name == "bg_pool_act_bwd":                         # N, Ho, Wo, C: read gy, y (pooled), write gu (4x)
            return 2.0 * 6 * a[0] * a[1] * a[2] * a[3]
        if name == "bg_pool_act_fwd":
            return 2.0 * 5 * a[0] * a[1] * a[2] * a[3]
        if name == "bg_style_modulate":                       # N, Cin, Cout, HW: fp32 W in, N bf16 packs out
            return 9.0 * a[1] * a[2] * (4.0 + 2.0 * a[0])
    except Exception:
        return None
    return None


def roofline_from_calls(rec, workload, peaks, img_per_s_per_gpu, gflop_img, use_traffic_file):
    """rec: [(C-ABI call, scalar args, ms)] of ONE iteration, each timed with a CUDA-event pair on the launching stream."""
    fam = {}
    dom = {"n": 0, "ms": 0.0, "flops": 0.0, "exec": 0.0, "bytes": 0.0}
    wg = {"ms": 0.0, "flops": 0.0, "exec": 0.0}
    aux = {}
    for name, a, t in rec:
        f = fam.setdefault(name, [0, 0.0])
        f[0] += 1
        f[1] += t
        if name in FPROP_CALLS or name in WGRAD_CALLS:
            fl, ex, by = conv_call_work(name, a)
            tgt = dom if name in FPROP_CALLS else wg
            tgt["ms"] += t
            tgt["flops"] += fl
            tgt["exec"] += ex
            if name in FPROP_CALLS:
                dom["n"] += 1
                dom["bytes"] += by
        else:
            by = aux_call_bytes(name, a)
            if by is not None and t >= 0.015:                 # below ~15 us an event pair mostly times the launch itself
                e = aux.setdefault((name, a), [0, 0.0, by])
                e[0] += 1
                e[1] += t
    tot_ms = sum(v[1] for v in fam.values())
    dom["ms"] = max(dom["ms"], 1e-9)
    tf = dom["flops"] / (dom["ms"] / 1e3) / 1e12 if dom["n"] else 0.0
    tf_exec = dom["exec"] / (dom["ms"] / 1e3) / 1e12 if dom["n"] else 0.0
    gbs = dom["bytes"] / (dom["ms"] / 1e3) / 1e9 if dom["n"] else 0.0
    shares = {k: round(v[1] / tot_ms, 4) for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])[:8]}
    traffic = None
    if use_traffic_file:
        for fn in ("r2_dram_traffic_fprop_family_train256.json", "r1_dram_traffic_fprop_family_train256.json"):
            try:
                with open(os.path.join(ROOT, "profiles", fn)) as f:
                    tj = json.load(f)
                traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
                break
            except Exception:
                continue
    bound = BOUND.get(workload, "tensor")
    kernel = ("conv_halo_kernel / conv_fprop_kernel (tcgen05 implicit GEMM: fprop, dgrad, R1 tangent pass, fused pool / "
              "IN-stats / bias-grad epilogues)")
    common = {"kernel": kernel, "bound": bound, "traffic": traffic,
              "traffic_note": "DRAM read+write bytes of ALL launches of this kernel family in one iteration (ncu capture under "
                              "profiles/); algorithmic_bytes is the same sum of 2*N*(Cin*Hin*Win+Cout*Hout*Wout)",
              "algorithmic_bytes": dom["bytes"], "algorithmic_flops": dom["flops"], "executed_flops": dom["exec"],
              "launches_per_step": dom["n"], "share_of_step": round(dom["ms"] / tot_ms, 4)}
    if bound == "hbm":
        # the dominant layers of this workload have <= 32 channels: below the 216 flop/B ridge
        roofline = dict(common, achieved=round(gbs, 1), peak=peaks["hbm"], unit="GB/s", frac=round(gbs / peaks["hbm"], 4),
                        peak_source=peaks["src"] + ", copy bandwidth",
                        tensor_tflops=round(tf, 1), tensor_frac=round(tf / peaks["tf_sustained"], 4))
    else:
        roofline = dict(common, achieved=round(tf, 1), peak=peaks["tf_sustained"], unit="TFLOP/s",
                        frac=round(tf / peaks["tf_sustained"], 4),
                        frac_note="reference-formulation FLOPs (the folded conv+pool launches are credited with the 3x3 "
                                  "count they replace, SURVEY.md §8d); executed_frac counts the MACs really issued",
                        executed_tflops=round(tf_exec, 1), executed_frac=round(tf_exec / peaks["tf_sustained"], 4),
                        hbm_gbs=round(gbs, 1),
                        peak_source=peaks["src"] + ", sustained bf16 (kernel timed inside a long step)")
    roofline["wgrad_tflops"] = round(wg["flops"] / (wg["ms"] / 1e3) / 1e12, 1) if wg["ms"] > 0 else None
    roofline["wgrad_executed_tflops"] = round(wg["exec"] / (wg["ms"] / 1e3) / 1e12, 1) if wg["ms"] > 0 else None
    roofline["wgrad_share_of_step"] = round(wg["ms"] / tot_ms, 4)
    roofline["step_share_by_call"] = shares
    roofline["step_model_flops_frac"] = round(img_per_s_per_gpu * gflop_img * 1e9 / (peaks["tf_sustained"] * 1e12), 4)
    # achieved GB/s of the HBM-bound helpers (the fused norm / noise / resample / 1x1 paths), worst first by lost time
    rows = []
    for (name, a), (cnt, ms_sum, by) in aux.items():
        g = by * cnt / (ms_sum / 1e3) / 1e9
        rows.append({"call": name, "args": list(a[:6]), "launches": cnt, "us": round(ms_sum / cnt * 1e3, 1),
                     "GB/s": round(g, 1), "frac_of_hbm": round(g / peaks["hbm"], 3),
                     "lost_us": round(max(0.0, ms_sum * 1e3 - by * cnt / (peaks["hbm"] * 1e9) * 1e6), 1)})
    rows.sort(key=lambda r: -r["lost_us"])
    aux_table = {"peak_GB/s": peaks["hbm"], "share_of_step": round(sum(v[1] for v in aux.values()) / tot_ms, 4),
                 "worst": rows[0] if rows else None, "top": rows[:10],
                 "note": "per C-ABI call, CUDA events on the launching stream inside the step; bytes = one read of each "
                         "input map + one write of each output map"}
    return roofline, aux_table


def workload_config(workload, world, style_mixing, batch=None):
    """The `config` object of the JSON line — built by ONE function so both arms print the same one."""
    steps, alpha, b, _, desc = WORKLOADS[workload]
    b = batch or b
    return {"workload": f"{workload}: {desc}", "resolution": 4 * 2 ** (steps - 1), "progressive_steps": steps, "alpha": alpha,
            "batch_per_gpu": b, "global_batch": b * world, "parallelism": f"dp{world}",
            "style_mixing": bool(style_mixing),
            "optimizer": "Adam(lr=0.002, betas=(0,0.99)), both updates inside the step",
            "loss": "non-saturating logistic + R1 (lambda=10) with double-backward every step",
            "l2_policy": "inputs rotate over 4 batches; per-step working set (GBs of activations) >> 126 MB L2"}


def metric_name(workload):
    return "G+D train img/s at 256x256" if workload == "train256" else f"G+D train img/s ({workload})"


def torch_gpu_leg(workload, batch, device, iters=3):
    """The reference's arithmetic on THIS GPU: the oracle (plain torch ops, i.e. cuDNN / cuBLAS kernels, fp32 with torch's
    default TF32 convolution setting, autograd incl. create_graph double-backward) running the same iteration at the
    workload's real batch.  The checker timed as a side note: it answers 'what would unmodified PyTorch code
    reach on the same device', which the CPU arm cannot."""
    from oracle import gan_oracle as O

    steps, alpha, _, gflop_img, _ = WORKLOADS[workload]
    try:
        G, D = O.make_state("gen", 0), O.make_state("critic", 0)
        args = (O.make_latents(batch, 1), O.make_latents(batch, 2), O.make_images(batch, steps, 3),
                O.make_noise(batch, steps, 4), O.make_noise(batch, steps, 5))
        O.train_iteration(G, D, *args, steps, alpha, LAMBDA, device=device)           # warm-up (cuDNN autotune)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            O.train_iteration(G, D, *args, steps, alpha, LAMBDA, device=device)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
        out = {"value": round(batch / dt, 1), "unit": "img/s", "ms_per_step": round(dt * 1e3, 1), "batch": batch,
               "kind": "oracle port on the GPU (torch/cuDNN fp32, TF32 convolutions at torch's default, Adam excluded, "
                       "state re-uploaded every iteration)"}
    except Exception as e:  # noqa: BLE001 - a side note must never fail the bench
        out = {"unavailable": f"{type(e).__name__}: {str(e)[:120]}"}
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import bg_native as bgn
    import dist as bdist

    rank, world, local = bdist.init_from_env()
    if args.gpus != world:
        if rank == 0 and world > 1:
            print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    steps, alpha, batch, gflop_img, desc = WORKLOADS[args.workload]
    if args.batch:
        batch = args.batch
    R = 4 * 2 ** (steps - 1)
    style_mixing = (args.workload in STYLE_MIXING_DEFAULT) if args.style_mixing is None else bool(args.style_mixing)
    use_graph = bool(args.graph) and world == 1
    tr = make_trainer(steps, alpha, batch, device, style_mixing, graph=use_graph)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    POOL = 4
    host_real = [torch.rand(batch, 3, R, R, generator=g).mul_(2).sub_(1).pin_memory() for _ in range(POOL)]
    host_z = [torch.randn(2, batch, 512, generator=g).clamp_(-0.75, 0.75).pin_memory() for _ in range(POOL)]
    dev_real = [t.to(device) for t in host_real]
    dev_z = [t.to(device) for t in host_z]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    copy_stream = torch.cuda.Stream(device=device)
    RUNAHEAD = int(os.environ.get("BG_BENCH_RUNAHEAD", "2"))   # 0: unlimited

    def timed(n_steps, from_host):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        if from_host:
            # every step's inputs come from pinned host memory INSIDE the timed region; the copy of step i+1 is issued on
            # a copy stream while step i computes (what a prefetching data loader does), step 0's copy is not hidden
            main = torch.cuda.current_stream(device)

            def upload(k):
                with torch.cuda.stream(copy_stream):
                    r = host_real[k % POOL].to(device, non_blocking=True)
                    z = host_z[k % POOL].to(device, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                return r, z, ev

            nxt = upload(0)
        # Device-resident leg: the host never reads anything back, so it would queue all K steps at once.  With several
        # ranks per node that makes the step time noisy (launch queues fill, NCCL's kernels of step i + 1 are launched at
        # different moments on different ranks): like a training loop that reads its losses one step late, the host stays
        # at most RUNAHEAD steps in front of the device (an event wait; no effect on a GPU-bound single process).
        marks = []
        for i in range(n_steps):
            j = i % POOL
            if not from_host and RUNAHEAD > 0 and len(marks) >= RUNAHEAD:
                marks.pop(0).synchronize()
            if from_host:
                real, zz, ev = nxt
                main.wait_event(ev)
                real.record_stream(main)
                zz.record_stream(main)
                if i + 1 < n_steps:
                    nxt = upload(i + 1)
                tr.iteration(real, zz[0], zz[1], read_losses=True)
            else:
                tr.iteration(dev_real[j].clone(), dev_z[j][0].clone(), dev_z[j][1].clone(), read_losses=False)
                if RUNAHEAD > 0:
                    ev = torch.cuda.Event()
                    ev.record()
                    marks.append(ev)
        if from_host:
            tr.flush_reads()                                 # the last iteration's losses, inside the timed region
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=device)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return ms.item()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    timed(args.warmup, from_host=False)                      # warm-up (also builds the weight packs)
    # settle: the caching allocator (GB-sized activations, freed in a different order by backward) and Python's
    # cyclic GC sometimes need a few more iterations before the step time is stationary; keep warming up (untimed,
    # at most 8 more iterations) until two consecutive iterations are within 4 % of the fastest one seen
    import gc
    gc.collect()
    gc.disable()                                             # no collector pauses inside the timed regions
    extra, best, calm = 0, float("inf"), 0
    while extra < 8 and calm < 2:
        t = timed(1, from_host=False)
        extra += 1
        best = min(best, t)
        calm = calm + 1 if t <= 1.04 * best else 0
        if rank == 0:
            st = torch.cuda.memory_stats(device)
            print(f"[settle] iteration {extra}: {t:.2f} ms  reserved {st['reserved_bytes.all.current'] / 2**30:.2f} GiB "
                  f"cudaMalloc retries {st.get('num_alloc_retries', 0)} segments {st.get('segment.all.current', 0)}",
                  file=sys.stderr)
    # The GPU boxes are shared hosts: a neighbour's burst on the host cores now and then slows the Python thread that
    # feeds ~500 C-ABI calls (~700 launches) per iteration, and one timed region in four or five comes out 10-80 % long with identical
    # clocks and allocator state.  Each leg is therefore timed REPEATS (5) times (each region = exactly K steps between
    # barrier + synchronize, max over ranks) and the MEDIAN region is reported; all regions are listed in the line.
    REPEATS = 5
    n0 = bgn.launch_count
    bytes0 = tr.sync.bytes_reduced
    w0 = time.time()
    reps = [timed(args.steps, from_host=False) for _ in range(REPEATS)]
    allreduce_bytes_per_step = (tr.sync.bytes_reduced - bytes0) // (REPEATS * args.steps)
    clocks.window(w0, time.time())
    launches = (bgn.launch_count - n0) // REPEATS
    clk = clocks.stop() if rank == 0 else None
    parity_variant = None
    if style_mixing:
        tr.style_mixing = False
        timed(2, from_host=False)
        pv = sorted(timed(args.steps, from_host=False) for _ in range(3))[1]
        parity_variant = {"style_mixing": False, "value": round(batch * world * args.steps / (pv / 1e3), 2), "unit": "img/s",
                          "ms_per_step": round(pv / args.steps, 3),
                          "note": "same workload exactly as the reference's Generator.forward runs it (one latent); this is "
                                  "the variant the parity tests and the reference arm cover"}
        tr.style_mixing = True
    timed(1, from_host=True)
    reps_e2e = [timed(args.steps, from_host=True) for _ in range(REPEATS)]
    ms, ms_e2e = sorted(reps)[REPEATS // 2], sorted(reps_e2e)[REPEATS // 2]
    imgs = batch * world * args.steps
    value = imgs / (ms / 1e3)
    e2e_value = imgs / (ms_e2e / 1e3)

    # ---- roofline leg: CUDA-event pair around every C-ABI call of one more iteration (rank 0's numbers)
    peaks = load_peaks()
    barrier()
    saved_graphs, tr.graphs = tr.graphs, None                # per-call timing needs the eager call sequence
    bgn.start_timing()
    tr.iteration(dev_real[0].clone(), dev_z[0][0].clone(), dev_z[0][1].clone(), read_losses=False)
    rec = bgn.stop_timing()
    tr.graphs = saved_graphs
    roofline, aux_table = roofline_from_calls(rec, args.workload, peaks, value / world, gflop_img,
                                              use_traffic_file=(args.workload == "train256" and not args.batch))

    # replicas must still hold identical parameters after all those averaged updates (outside every timed region)
    replica_diff = None
    if world > 1:
        with torch.no_grad():
            mine = torch.cat([p.detach().reshape(-1) for m in (tr.gen, tr.critic) for p in m.parameters()])
            ref0 = mine.clone()
            torch.distributed.broadcast(ref0, 0)
            d = (mine - ref0).abs().max().reshape(1)
            torch.distributed.all_reduce(d, op=torch.distributed.ReduceOp.MAX)
            replica_diff = float(d)
            del mine, ref0

    sampling = None
    if rank == 0 and world == 1 and not args.no_sampling:
        del tr.critic
        torch.cuda.empty_cache()
        sampling = sampling_leg(device, gen=tr.gen)
    cpu = torch_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        del tr
        torch.cuda.empty_cache()
        torch_gpu = torch_gpu_leg(args.workload, batch, device)
        cpu = cpu_leg(args.workload, max_seconds=40.0, steps_cap=3, warmup=1)

    if rank == 0:
        h2d = host_real[0].numel() * 4 + host_z[0].numel() * 4
        line = {
            "metric": metric_name(args.workload),
            "value": round(value, 2), "unit": "img/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_extra_settle_steps": extra,
            "timed_regions_ms_per_step": {"value": [round(t / args.steps, 3) for t in reps],
                                          "e2e": [round(t / args.steps, 3) for t in reps_e2e], "reported": "median"},
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.workload, world, style_mixing, args.batch or None),
            "host_runahead_steps": RUNAHEAD,
            "parity_variant": parity_variant,
            "clocks": clk,
            "e2e": {"value": round(e2e_value, 2), "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8, "loss_readback": "both losses copied to pinned host memory every step, consumed one step later (no queue drain)",
                    "input_feed": "each step's images + latents copied from pinned host memory inside the timed region, "
                                  "step i+1 on a copy stream while step i computes",
                    "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": launches,
            "cuda_graph": use_graph,
            "roofline": roofline,
            "aux_kernels": aux_table,
            "cpu_baseline": cpu,
            "torch_gpu_baseline": torch_gpu,
            "sampling_512": sampling,
            "replicas_max_abs_param_diff": replica_diff,
            "grad_allreduce_bytes_per_step": allreduce_bytes_per_step if world > 1 else 0,
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# sampling leg: generate_samples.py's hot call, Generator.forward(z, steps=8) under no_grad (BASELINE configs[4])
# ------------------------------------------------------------------------------------------------------
G_FWD_GFLOP_512 = 21.25      # per image, reference formulation


def sampling_leg(device, batch=256, steps=8, iters=3, warmup=2, gen=None):
    """512x512 sampling, batch 256, one GPU.  `value`: latents resident in HBM, images left in HBM.  `e2e`: latents
    from pinned host memory and the finished images copied back to pinned host memory inside the timed region
    (what generate_samples.py does before utils.save_image).  Per-layer noise is drawn on the device by the model
    exactly like the reference (gan.py:189-197)."""
    import gan

    if gen is None:
        torch.manual_seed(0)
        gen = gan.Generator().to(device)
        with torch.no_grad():
            for n, p in gen.named_parameters():
                if n.endswith("bias") or n.endswith("inject_noise.weights"):
                    p.add_(0.05 * torch.randn_like(p))
    gen.eval()
    R = 4 * 2 ** (steps - 1)
    g = torch.Generator(device="cpu").manual_seed(7)
    host_z = [torch.randn(batch, 512, generator=g).clamp_(-0.75, 0.75).pin_memory() for _ in range(2)]
    dev_z = [t.to(device) for t in host_z]
    host_img = [torch.empty(batch, 3, R, R).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)

    def run(n, from_host):
        """from_host: the 805 MB image copy of batch i runs on a copy stream (double-buffered pinned buffers) while
        batch i+1 is generated; the timed region ends when the last copy has landed."""
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        held = [None, None]
        with torch.no_grad():
            for i in range(n):
                if from_host:
                    img = gen(host_z[i % 2].to(device, non_blocking=True), steps=steps, alpha=None)
                    ready = torch.cuda.Event()
                    ready.record(main)
                    with torch.cuda.stream(copy_stream):
                        copy_stream.wait_event(ready)
                        host_img[i % 2].copy_(img, non_blocking=True)
                    img.record_stream(copy_stream)
                    held[i % 2] = img
                else:
                    img = gen(dev_z[i % 2], steps=steps, alpha=None)
                del img
        main.wait_stream(copy_stream)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1)

    run(warmup, False)
    ms = sorted(run(iters, False) for _ in range(3))[1]
    run(2, True)                  # touches BOTH pinned image buffers: the first copy into fresh pinned pages is ~10x slower
    # median of 3 regions: on the shared hosts the 805 MB device->host copies now and then run at a fraction of the PCIe rate
    ms_e2e = sorted(run(iters, True) for _ in range(3))[1]
    peaks = load_peaks()
    v = batch * iters / (ms / 1e3)
    # roofline of the sampling batch: every C-ABI call of one more batch timed with CUDA events on the launching stream
    import bg_native as bgn

    torch.cuda.synchronize()
    bgn.start_timing()
    with torch.no_grad():
        gen(dev_z[0], steps=steps, alpha=None)
    rec = bgn.stop_timing()
    conv_ms = sum(t for n, a, t in rec if n in FPROP_CALLS)
    conv_by = sum(conv_call_work(n, a)[2] for n, a, t in rec if n in FPROP_CALLS)
    conv_fl = sum(conv_call_work(n, a)[0] for n, a, t in rec if n in FPROP_CALLS)
    tot = sum(t for _, _, t in rec)
    gbs = conv_by / (max(conv_ms, 1e-9) / 1e3) / 1e9
    s_roof = {"kernel": "conv_halo_kernel fused style / upsample forms (bg_conv_style_fprop) + small-map convs",
              "bound": "hbm", "achieved": round(gbs, 1), "peak": peaks["hbm"], "unit": "GB/s",
              "frac": round(gbs / peaks["hbm"], 4), "traffic": None,
              "algorithmic_bytes": conv_by, "share_of_batch": round(conv_ms / max(tot, 1e-9), 4),
              "tensor_tflops": round(conv_fl / (max(conv_ms, 1e-9) / 1e3) / 1e12, 1),
              "note": "the 256x256 / 512x512 layers (16-64 channels) that dominate the batch are below the 216 flop/B ridge "
                      "(SURVEY.md §8a); bytes = 2*N*(Cin*Hin*Win + Cout*H*W) per conv, upsampled inputs counted at low "
                      "resolution",
              "step_share_by_call": {k: round(sum(t for n, _, t in rec if n == k) / max(tot, 1e-9), 4)
                                     for k in sorted({n for n, _, _ in rec},
                                                     key=lambda k: -sum(t for n, _, t in rec if n == k))[:6]}}
    return {"metric": "generate img/s at 512x512", "roofline": s_roof, "value": round(v, 1), "unit": "img/s", "batch": batch, "iters": iters,
            "ms_per_batch": round(ms / iters, 2),
            "e2e": {"value": round(batch * iters / (ms_e2e / 1e3), 1), "unit": "img/s",
                    "h2d_bytes_per_step": batch * 512 * 4, "d2h_bytes_per_step": batch * 3 * R * R * 4},
            "model_flops_frac": round(v * G_FWD_GFLOP_512 * 1e9 / (peaks["tf_sustained"] * 1e12), 4),
            "config": f"Generator.forward(z, steps={steps}, alpha=None) under no_grad, batch {batch}, internal randn noise"}


# ------------------------------------------------------------------------------------------------------
# the CPU leg (oracle port of the reference on the host cores)
# ------------------------------------------------------------------------------------------------------
def cpu_leg(workload, max_seconds, steps_cap, warmup=0):
    from oracle import gan_oracle as O

    steps, alpha, _, gflop_img, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # config 1 (4x4) runs in full on the CPU; the others at the smallest batch the minibatch-stddev groups allow
    # (gan.py:269; img/s is batch-insensitive on the CPU)
    batch = WORKLOADS[workload][2] if workload == "train4" else 4
    G, D = O.make_state("gen", 0), O.make_state("critic", 0)
    done, t_total = 0, 0.0
    for i in range(warmup + steps_cap):
        args = (O.make_latents(batch, i), O.make_latents(batch, 100 + i), O.make_images(batch, steps, i),
                O.make_noise(batch, steps, i), O.make_noise(batch, steps, 100 + i))
        t0 = time.perf_counter()
        O.train_iteration(G, D, *args, steps, alpha, LAMBDA)
        dt = time.perf_counter() - t0
        if i >= warmup:
            done += 1
            t_total += dt
        if t_total > max_seconds:
            break
    return {"value": round(batch * done / t_total, 3), "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
            "batch": batch, "iterations": done, "warmup_iterations": warmup, "ms_per_iteration": round(t_total / done * 1e3, 1),
            "sample": f"{done} timed iteration(s) after {warmup} warm-up of the same workload at batch {batch} (fp32, torch "
                      f"CPU, {cores} host threads; forward+R1 double-backward+G backward, Adam excluded)",
            "gflops": round(batch * done * gflop_img / t_total, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = cpu_leg(args.workload, max_seconds=240.0, steps_cap=max(1, args.steps), warmup=max(1, min(args.warmup, 1)))
    style_mixing = (args.workload in STYLE_MIXING_DEFAULT) if args.style_mixing is None else bool(args.style_mixing)
    line = {"impl": "reference", "metric": metric_name(args.workload),
            "value": cpu["value"], "unit": "img/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            # one "step" of this arm is one iteration of the bounded sample: the SAME workload at batch `sample_batch`
            "ms_per_step": cpu["ms_per_iteration"], "sample_batch": cpu["batch"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, max(1, args.gpus), style_mixing, args.batch or None),
            "note": "the reference's CPU path (oracle port of gan.py / train.py:135-217, pinned to the real gan.py by "
                    "tests/golden) on this box's host cores, one process whatever --gpus says; a bounded sample of the "
                    "configured workload at batch `sample_batch` (img/s is batch-insensitive on the CPU); the reference "
                    "has no style mixing, so the sample runs Generator.forward as the reference defines it",
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def _protect_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


_REAL_STDOUT = None


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train256", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (debugging only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--style-mixing", dest="style_mixing", action="store_true", default=None,
                    help="train with the style-mixing extension (two latents, per-step crossover); default: on for train256 "
                         "(BASELINE configs[2] names it), off elsewhere")
    ap.add_argument("--no-style-mixing", dest="style_mixing", action="store_false")
    ap.add_argument("--no-sampling", action="store_true", help="skip the 512x512 sampling leg")
    ap.add_argument("--graph", dest="graph", action="store_true", default=False,
                    help="replay each iteration from a CUDA graph (Trainer.enable_graphs; single GPU only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_b200(args)


if __name__ == "__main__":
    main()
