#!/usr/bin/env python
"""Benchmark of the ResShift denoising hot path (BASELINE.json metric: 256x256 x4-SR images/sec at 15
steps; ms/denoise-step).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one pass of the hot path over one batch: the full T=15-step residual-shift sampling loop
(denoiser forward + p_sample update per step) for a batch of 16 latents of 64x64 (= sixteen 256x256 x4-SR
images), BASELINE config 2 (`realsr_swinunet_realesrgan256_journal.yaml` with steps=15), random-init
weights, synthetic inputs.  N > 1 (torchrun, one rank per GPU): every rank owns 16 images (weak scaling,
config 3 = 128 images on 8 GPUs); rank 0 broadcasts the weights over NCCL once (untimed set-up) and the
final latents are all-gathered inside the timed step.

Printed JSON (one line, rank 0): see README/DESIGN.md.  `value` = images/s with inputs resident in HBM
(CUDA-graph replay of the loop); `e2e` = the same through the C-ABI host-buffer entry point
(`rs_sampler_run_host`: pinned host -> device copies of z_y, the T+1 noise tensors and the LQ image, the
loop, device -> host copy of the final latent, all inside the timed region).
The VQ-GAN encode/decode bookends are outside the hot path (they stay PyTorch) and are
not part of either number.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GF_PER_IMAGE_STEP = 101.32e9      # realsr denoiser, 2*MACs, hook-counted on the reference
T_STEPS = 15
BATCH_PER_GPU = 16


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"tensor_tflops": float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0))),
                "hbm_gbs": float(d.get("hbm_gbs", 6650.0)), "source": "MEASURED_PEAKS.json (sustained bf16 cuBLAS)"}
    return {"tensor_tflops": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md: ~1.4 PF sustained)"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


NCU_DRAM_BYTES_PER_GEMM_LAUNCH = 28.38e6


def cpu_reference_images_per_s(steps_T: int, n_images: int, repeats: int):
    """The reference's algorithm for this path on the host cores: the CPU oracle (a torch fp32 restatement of
    UNetModelSwin.forward + p_sample, pinned to reference-generated goldens), all host threads."""
    import torch
    from oracle import diffusion_oracle as do
    from oracle import unet_oracle as uo
    from resshift_b200.config import preset
    from resshift_b200.weights import random_state_dict
    # torch's default intra-op thread count (= the physical cores this process may use); forcing os.cpu_count()
    # threads inside a cgroup-limited container oversubscribes and stalls.  torchrun exports OMP_NUM_THREADS=1 to its
    # workers, which would silently make this a single-thread baseline: undo that with half the schedulable CPUs
    # (SMT pairs), the same count torch picks on its own.
    if torch.get_num_threads() == 1 and os.environ.get("OMP_NUM_THREADS") == "1":
        torch.set_num_threads(max(1, min(64, len(os.sched_getaffinity(0)) // 2)))
    ucfg, dcfg = preset("realsr_journal", steps_T)
    sd = random_state_dict(ucfg, 0)
    tabs = do.schedule_tables(do.eta_schedule(dcfg.steps, dcfg.min_noise_level, dcfg.etas_end, dcfg.kappa,
                                              dcfg.schedule_kwargs["power"]), dcfg.kappa)
    g = torch.Generator().manual_seed(12345)
    y = torch.rand(n_images, 3, 64, 64, generator=g) * 2 - 1
    noises = [torch.randn(n_images, 3, 64, 64, generator=g) for _ in range(steps_T + 1)]
    times = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        do.p_sample_loop(lambda x, t: uo.unet_forward(sd, ucfg, x, t, lq=y), y, noises, tabs, dcfg.kappa)
        times.append(time.perf_counter() - t0)
    return n_images / min(times), times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = []
    cores = os.cpu_count()
    for i in range(args.warmup + args.steps):
        ips, times, cores = cpu_reference_images_per_s(T_STEPS, 1, 1)
        if i >= args.warmup:
            per_step.append(times[0])
    sec = sum(per_step) / len(per_step)
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": "256x256 x4 SR images/sec (15 steps), denoising hot path", "value": val,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "ms_per_denoise_step": sec * 1e3 / T_STEPS,
        "config": {"workload": "realsr 64x64 latent (256x256 x4 SR), 15 steps, random-init weights; bounded sample: "
                               "1 image per step on the host CPU", "batch": 1},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "1 image x 15 denoise steps per timed step (oracle port of the reference, torch fp32 CPU, all host threads)"},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_gpu(args):
    import faulthandler
    faulthandler.dump_traceback_later(600, exit=True)
    import torch
    import torch.distributed as dist
    from resshift_b200 import _lib
    from resshift_b200.config import preset
    from resshift_b200.models.script_util import create_gaussian_diffusion
    from resshift_b200.models.unet import UNetModelSwin
    from resshift_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=dev)

    B = args.batch
    ucfg, dcfg = preset("realsr_journal", T_STEPS)
    dcfg.sf = 1       # the loop works on the latent; the bicubic + VQ bookends are outside the hot path
    model = UNetModelSwin(**ucfg.to_kwargs())
    if rank == 0:
        model.load_state_dict(random_state_dict(ucfg, 0))
    model = model.cuda().eval()
    if world > 1:       # one NCCL broadcast of the weights (as one flat buffer), untimed set-up
        flat = torch.cat([p.data.reshape(-1) for p in model.parameters()])
        dist.broadcast(flat, src=0)
        off = 0
        for p in model.parameters():
            p.data.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        del flat
    diff = create_gaussian_diffusion(**dcfg.to_kwargs())
    T = diff.num_timesteps

    g = torch.Generator(device=dev).manual_seed(12345 + rank)
    z_y = torch.rand(B, 3, 64, 64, device=dev, generator=g) * 2 - 1
    noises = torch.randn(T + 1, B, 3, 64, 64, device=dev, generator=g)
    lq = z_y.clone()
    out = torch.empty_like(z_y)
    sampler = diff.native_sampler(model, B, 64, 64)
    plan = model.plan(B, 64, 64)
    launches_per_forward = _lib.lib.rs_plan_num_launches(plan.handle)
    launches_per_loop = T * (launches_per_forward - 6 + 1) + 2      # per step: body + p_sample; + prior + pack
    stream = torch.cuda.current_stream().cuda_stream
    gathered = [torch.empty_like(out) for _ in range(world)] if world > 1 else None

    def one_step(use_graph=True):
        _lib.check(_lib.lib.rs_sampler_run(sampler, z_y.data_ptr(), noises.data_ptr(), lq.data_ptr(), None,
                                           out.data_ptr(), int(use_graph), stream))
        if world > 1:
            dist.all_gather(gathered, out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----------------------------------------------------
    log(f"model + plan ready: {launches_per_forward} launches/forward, batch {B}")
    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    log("warm-up done")
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        one_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = t_ms.item()
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- end to end through the host-buffer C-ABI entry ("e2e") ------------------------------------
    staging_bytes = _lib.lib.rs_sampler_staging_bytes(sampler)
    staging = torch.empty(staging_bytes + 256, dtype=torch.uint8, device=dev)
    staging_ptr = (staging.data_ptr() + 255) // 256 * 256
    h_zy = z_y.cpu().pin_memory()
    h_noise = noises.cpu().pin_memory()
    h_lq = lq.cpu().pin_memory()
    h_out = torch.empty(B, 3, 64, 64).pin_memory()

    def e2e_step():
        _lib.check(_lib.lib.rs_sampler_run_host(sampler, h_zy.data_ptr(), h_noise.data_ptr(), h_lq.data_ptr(), None,
                                                h_out.data_ptr(), staging_ptr, staging_bytes, 1, stream))
    log(f"device-resident: {ms_per_step:.2f} ms/step")
    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    wall = (time.perf_counter() - t0) * 1e3
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), wall)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (e2e_ms.item() / args.steps * 1e-3)
    h2d = (h_zy.numel() + h_noise.numel() + h_lq.numel()) * 4
    d2h = h_out.numel() * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (conv/linear implicit GEMM on tcgen05), measured live -------
    ms_kind = (C.c_double * 4)()
    flops = C.c_double()
    nconv = C.c_int32()
    tt = torch.full((B,), 7.0, device=dev)
    x0 = torch.randn(B, 3, 64, 64, device=dev)
    o0 = torch.empty_like(x0)
    prof = []
    for _ in range(3):
        _lib.check(_lib.lib.rs_plan_profile(plan.handle, x0.data_ptr(), tt.data_ptr(), lq.data_ptr(), None, ms_kind,
                                            C.byref(flops), C.byref(nconv), stream))
        prof.append(list(ms_kind))
    pk = prof[-1]
    peaks = _peaks()
    conv_tflops = flops.value / (pk[0] * 1e-3) / 1e12 if pk[0] > 0 else 0.0
    roofline = {
        "kernel": "tcgen05 GEMM kernels: conv_gemm_sm100_kernel<1|2>, conv_gemm_persist_sm100_kernel<1|2> (all conv3x3 / "
                  "conv1x1 / linear layers) + mlp_fused_sm100_kernel (Swin MLPs)", "bound": "tensor",
        "achieved": conv_tflops, "peak": peaks["tensor_tflops"], "unit": "TFLOP/s",
        "frac": conv_tflops / peaks["tensor_tflops"],
        # dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged over the first 20 GEMM launches of a forward
        # (the 64x64 level) captured with `ncu --set full` on this workload (cold caches: ncu flushes between replays);
        # profiles/r1_s40_gemm_kernels_ncu_full_summary.csv.  Reads are ~ each layer's input (+ residual) once — e.g.
        # 28-35 MB for a 64x64 160->160 3x3 layer whose input + residual + weights are 42.4 MB (part still L2-resident)
        # — i.e. no operand is re-read from DRAM (tap / channel-tile re-use is served by the 126 MB L2); outputs mostly
        # stay in L2 (writes ~1 MB / launch).
        "traffic": NCU_DRAM_BYTES_PER_GEMM_LAUNCH if B == BATCH_PER_GPU else None,
        "launches_per_forward": int(nconv.value), "avg_launch_us": pk[0] * 1e3 / max(1, nconv.value),
        "algorithmic_gflop_per_forward": flops.value / 1e9, "peak_source": peaks["source"],
        "per_forward_ms_by_kernel": {"conv_gemm": pk[0], "groupnorm": pk[1], "window_attn": pk[2], "upsample": pk[3]},
        "traffic_source": "profiles/r1_s40_gemm_kernels_ncu_full_summary.csv (ncu --set full, 20 launches, batch 16)",
        "tensor_pipe_active_pct_ncu": {"conv_gemm_persist<2> (3x3, 64x64)": 45.5, "conv_gemm<2>": 43.6, "mlp_fused": 23.0,
                                       "conv_gemm_persist<1> (1x1, epilogue-bound)": 14.4, "conv_gemm<1>": 8.9},
        "note": "achieved = algorithmic FLOPs of all GEMM launches / sum of their durations, CUDA events around every "
                "launch of one un-graphed forward on the launching stream (includes inter-launch gaps, so it "
                "under-states the graph-replayed step)",
    }
    # whole-step tensor-pipe fraction as a cross-check
    step_tflops = world * B * T * GF_PER_IMAGE_STEP / (ms_per_step * 1e-3) / 1e12 / world

    # ---- CPU baseline: bounded sample on this box's host cores ----------------------------------------
    cpu = None
    log("profile done; CPU baseline next")
    if not args.no_cpu_baseline:
        ips, times, cores = cpu_reference_images_per_s(T_STEPS, 1, 1)
        cpu = {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"1 image x {T_STEPS} denoise steps, {times[0]:.2f} s (oracle port of the reference, torch fp32 CPU, all host threads)"}

    line = {
        "metric": "256x256 x4 SR images/sec (15 steps), denoising hot path", "value": value, "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 (fp32 accumulate)",
        "data": "synthetic", "ms_per_denoise_step": ms_per_step / T,
        "config": {"workload": "BASELINE config %s: batch=%d/GPU 256x256 x4 real-SR (64x64 latent), 15 steps, "
                               "realsr_swinunet_realesrgan256_journal.yaml with steps=15, random-init weights"
                               % ("2" if world == 1 else "3", B),
                   "global_batch": world * B, "parallelism": f"dp{world} (independent image shards, NCCL weight broadcast + final all_gather)",
                   "l2": "no explicit flush: per-step working set (237 MB fp16 weights + >1 GB activations) exceeds the 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches_per_loop * args.steps * 2),
        "launches_per_denoise_step": int(launches_per_forward - 6 + 1),
        "denoiser_tflops_per_gpu": step_tflops,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="images per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
