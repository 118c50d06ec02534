#!/usr/bin/env python
"""bench.py — CLIP-guidance throughput (loss forward + backward to the image), cutouts/s.

    python bench.py --gpus N --steps K --warmup W            # native sm_100a path
    python bench.py --impl reference --gpus N ...             # the reference algorithm on the host CPU cores
    torchrun --nproc-per-node N bench.py --gpus N ...         # N > 1: one rank per GPU over NCCL

Workload (BASELINE.json metric: "ViT-L/14 ... at 1/2/4/8 B200", configs[2]): ViT-L/14 @224, random-init weights,
N_gpu synthetic 3x512x512 images, 128 random cutouts per image, 2 random targets; the cutout table is sharded over
the ranks (one image's worth of cutouts per GPU => weak scaling) and the per-image gradient is summed with one NCCL
all-reduce.  One step = sample cutouts -> loss forward -> backward to images.grad.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (architecture, image H=W, cutouts per image, min cutout size, images per GPU)
    "vit_l14_224_128cut_512px": ("ViT-L-14", 512, 128, 128, 1),
    # BASELINE configs[2] as SURVEY 8(d) fixes it (B = 1): ONE 512x512 image, its 128 cutouts sharded over the ranks
    "vit_l14_224_128cut_512px_1image": ("ViT-L-14", 512, 128, 128, 0),
    "vit_b32_224_64cut_4x512px": ("ViT-B-32", 512, 64, 64, 4),
    "vit_l14_336_256cut_768px": ("ViT-L-14-336", 768, 256, 192, 1),
    # BASELINE configs[3] as written: ONE 768x768 image, 256 cutouts in total, sharded over the ranks (strong scaling)
    "vit_l14_336_256cut_768px_1image": ("ViT-L-14-336", 768, 256, 192, 0),
    "vit_b32_224_16cut_256px": ("ViT-B-32", 256, 16, 64, 1),
    # one rank's share of the strong-scaling record at 8 GPUs (16 of the 128 cutouts), as a single-GPU proxy for tuning
    "vit_l14_224_16cut_512px": ("ViT-L-14", 512, 16, 128, 1),
}
DEFAULT_WORKLOAD = "vit_l14_224_128cut_512px"
# roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per gemm_tcgen05_kernel launch, mean over the eight GEMM
# shapes of one transformer layer (forward qkv / out / fc / proj, backward dproj / dfc / dout / dqkv), READ from the
# committed `ncu --set full` summaries (tools/ncu_summ.py output) rather than typed in.
# (file, first GEMM row, rows): tools/profile_r02e.sh captures four consecutive GEMM launches of the timed step's forward
# (layer 10: qkv, out, fc, proj) and four of its backward (layer 17: dproj, dfc, dout, dqkv)
NCU_GEMM_TRAFFIC_FILES = {"vit_l14_224_128cut_512px": (("profiles/r02e_gemm_fwd_ncu_full.csv", 0, 4),
                                                       ("profiles/r02e_gemm_bwd_ncu_full.csv", 0, 4))}


def ncu_gemm_traffic(workload: str):
    """(bytes per launch, source) from the committed ncu summaries, or (None, reason)."""
    import csv
    files = NCU_GEMM_TRAFFIC_FILES.get(workload)
    if not files:
        return None, "no ncu capture committed for this workload"
    vals = []
    for rel, first, count in files:
        path = os.path.join(ROOT, rel)
        if not os.path.exists(path):
            return None, f"{rel} missing"
        with open(path) as f:
            rows = [r for r in csv.reader(line for line in f if not line.startswith("#"))]
        hdr = rows[0]
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        units = rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        mine = []
        for r in rows[2:]:
            if len(r) > max(ri, wi) and "gemm_tcgen05_kernel" in r[ki]:
                mine.append(float(r[ri].replace(",", "")) * scale.get(units[ri], 1.0)
                            + float(r[wi].replace(",", "")) * scale.get(units[wi], 1.0))
        vals += mine[first:first + count]
    if not vals:
        return None, "no gemm_tcgen05_kernel rows in the ncu summaries"
    return sum(vals) / len(vals), f"mean of {len(vals)} captured launches (one layer's eight shapes) in {', '.join(f[0] for f in files)}"
METRIC = "CLIP-guidance cutouts/sec (loss + image grad), ViT-L/14, at 1/2/4/8 B200"


_RESULT_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version banner at communicator
    creation), so file descriptor 1 is pointed at stderr for the run and the result goes out through a saved copy."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines: list[str] = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        # "under load": samples drawing noticeably more than idle power
        hot = [s for s, p in zip(sm, power) if p > 300.0] or sm
        return {"sm_mhz": statistics.median(hot) if hot else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_step_fn(arch: str, hw: int, min_size: int, sample_cutouts: int):
    """Returns a closure running one loss+backward over `sample_cutouts` cutouts with the CPU oracle."""
    from oracle import guidance as guidance_oracle
    from oracle import loss as loss_oracle
    from perceptor_b200 import cutouts
    from perceptor_b200.vit import SHAPES, random_state_dict

    shape = SHAPES[arch]
    sd = random_state_dict(shape, 0)
    g = torch.Generator().manual_seed(0)
    images = torch.rand(1, 3, hw, hw, generator=g)
    targets = torch.nn.functional.normalize(torch.randn(2, shape.embed, generator=g))
    weights = torch.ones(2)
    gen = torch.Generator().manual_seed(0)
    mult = loss_oracle.default_multiplier(arch)

    def step():
        rows = cutouts.sample_cutouts(gen, 1, hw, hw, sample_cutouts, 1.0, min_size, hw)
        img = images.clone().requires_grad_()
        loss = guidance_oracle.guidance_loss(img, rows.tolist(), sd, shape.image_size, shape.patch, shape.layers,
                                             shape.heads, targets, weights, mult)
        loss.backward()
        return float(loss.detach())

    return step


def time_cpu(arch, hw, min_size, sample_cutouts, steps, warmup, budget_s=25.0):
    torch.set_num_threads(os.cpu_count() or 1)
    step = cpu_step_fn(arch, hw, min_size, sample_cutouts)
    for _ in range(warmup):
        step()
    times = []
    t_start = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    med = statistics.median(times)
    return sample_cutouts / med, med, len(times)


def run_reference(args, rank: int):
    if rank != 0:
        return
    arch, hw, n_cut, min_size, imgs = WORKLOADS[args.workload]
    sample = args.cpu_sample
    value, med, done = time_cpu(arch, hw, min_size, sample, max(args.steps, 1), min(args.warmup, 1), budget_s=120.0)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "cutouts/s", "n_gpus": args.gpus,
        "steps": done, "warmup": min(args.warmup, 1), "ms_per_step": med * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "model": arch, "image": f"{hw}x{hw}", "cutouts_per_image": n_cut,
                   "note": f"each step is a bounded sample of {sample} cutouts of this workload on the host CPU"},
        "cpu_baseline": {"value": value, "unit": "cutouts/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} cutouts/step, {done} steps, median; fp32 torch CPU oracle "
                                   f"(reference resize_right + ruclip ViT restated), {os.cpu_count()} host cpus"},
        "e2e": {"value": value, "unit": "cutouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def parity_check(loss_mod, arch, shape, hw, min_size, device, n_check=4):
    """The timed configuration against the CPU oracle, outside the timed region: the SAME encoder (weights, targets,
    graph path) on a 4-cutout shard of the workload's cutout distribution.  Returns the loss relative error and the
    image-gradient cosine (north_star's bars: 1e-2 and 0.999)."""
    import numpy as np

    from oracle import guidance as guidance_oracle
    from perceptor_b200 import cutouts
    from perceptor_b200.guidance import GuidanceLossFn

    eng = loss_mod.model.engine()
    g = torch.Generator().manual_seed(1234)
    image = torch.rand(1, 3, hw, hw, generator=g)
    rows = cutouts.sample_cutouts(torch.Generator().manual_seed(99), 1, hw, hw, n_check, 1.0, min_size, hw)
    targets = loss_mod.encodings.detach().float().cpu()
    weights = loss_mod.weights.detach().float().cpu()
    img = image.to(device).requires_grad_()
    loss = GuidanceLossFn.apply(img, eng, eng.plan_cutouts(rows), loss_mod.encodings.detach().float().contiguous(),
                                loss_mod.weights.detach().float().contiguous(), float(loss_mod.multiplier), None)
    loss.backward()
    sd = {k: v.detach().float().cpu() for k, v in loss_mod.model.state_dict_openai().items()}
    torch.set_num_threads(os.cpu_count() or 1)
    ref_img = image.clone().requires_grad_()
    ref = guidance_oracle.guidance_loss(ref_img, rows.tolist(), sd, shape.image_size, shape.patch, shape.layers,
                                        shape.heads, targets, weights, float(loss_mod.multiplier))
    ref.backward()
    ga, gb = img.grad.detach().float().cpu().flatten().double(), ref_img.grad.flatten().double()
    cos = float((ga @ gb) / (ga.norm() * gb.norm() + 1e-300))
    rel = abs(float(loss.detach()) - float(ref.detach())) / max(abs(float(ref.detach())), 1e-30)
    return {"loss_rel_err": rel, "grad_cosine": cos, "cutouts": int(n_check), "tolerance": {"loss_rel_err": 1e-2,
            "grad_cosine": 0.999}, "ok": bool(rel <= 1e-2 and cos >= 0.999),
            "what": f"{arch}, {n_check} cutouts of one {hw}x{hw} image through the timed encoder vs the fp32 CPU oracle"}


# ------------------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------------------
def run_native(args, rank: int, world: int, local_rank: int):
    import torch.distributed as dist

    from perceptor_b200 import losses, native
    from perceptor_b200.vit import SHAPES

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the native guidance path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD

    arch, hw, n_cut, min_size, imgs_per_gpu = WORKLOADS[args.workload]
    shape = SHAPES[arch]
    n_images = imgs_per_gpu * world if imgs_per_gpu > 0 else 1  # 0: one image in total, its cutouts sharded
    cutouts_per_step = n_images * n_cut
    # weak scaling: every rank owns its images (shard="images": no gradient collective, only the scalar loss);
    # strong scaling: one image, its cutouts split over the ranks, image gradient summed with one all-reduce
    shard = "images" if imgs_per_gpu > 0 else "cutouts"

    loss_mod = losses.CLIP(arch, n_cutouts=n_cut, min_size=min_size, max_size=hw, seed=0, process_group=group, shard=shard)
    g = torch.Generator().manual_seed(0)
    target_enc = torch.randn(2, shape.embed, generator=g)
    loss_mod.add_encodings_(target_enc)
    eng = loss_mod.model.engine()
    eng.prebuild_tables(min_size, hw)
    all_images = torch.rand(n_images, 3, hw, hw, generator=g)
    if shard == "images":  # this rank's images only: what it uploads and what it gets a gradient for
        all_images = all_images[rank * imgs_per_gpu:(rank + 1) * imgs_per_gpu]
    host_images = all_images.contiguous().pin_memory()
    dev_images = host_images.to(device)
    host_grad = torch.empty_like(host_images).pin_memory()

    def step_device():
        img = dev_images.detach().requires_grad_()
        loss = loss_mod(img)
        loss.backward()
        return loss, img.grad

    def step_e2e():
        img = host_images.to(device, non_blocking=True).requires_grad_()
        loss = loss_mod(img)
        loss.backward()
        host_grad.copy_(img.grad, non_blocking=True)
        return float(loss.detach())  # D2H read of the scalar; synchronises the step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(args.warmup):
        step_device()
    barrier()

    lib = native.lib()
    profile = rank == 0
    graphs_on = eng.use_graphs
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    # timed region: K steps exactly as a user runs them (CUDA-graph replay of the ~390 launches, no instrumentation)
    total_ms = timed(step_device, args.steps)
    clock_info = clocks.stop() if rank == 0 else None
    # second pass over the same K steps for the per-kernel numbers: eager launches with a CUDA event pair around
    # every launch on the launch stream (the events cost 1-3 % of a step, which is why they stay out of `value`)
    # (every rank runs it: the steps contain the all-reduces; only rank 0 records events)
    eng.use_graphs = False
    step_device()
    if profile:
        lib.pcg_profile_enable(1)
        native.profile_collect()
    prof_ms = timed(step_device, args.steps)
    prof = native.profile_collect() if profile else None
    lib.pcg_profile_enable(0)
    eng.use_graphs = graphs_on
    ms_per_step = total_ms / args.steps
    value = cutouts_per_step / (ms_per_step * 1e-3)
    launches = args.steps * (eng.launches_fwd + eng.launches_bwd)

    # end to end through the public module API with HOST buffers (H2D of the images, D2H of loss + gradient)
    for _ in range(2):
        step_e2e()
    e2e_steps = max(2, min(args.steps, 10))
    e2e_ms = timed(step_e2e, e2e_steps) / e2e_steps
    e2e_value = cutouts_per_step / (e2e_ms * 1e-3)

    # The same workload with the LAST block run in full (every token through its out-projection / MLP, as the reference's
    # dense computation does) instead of on the class-token rows the loss reads: outputs are the same, 3.3 % more FLOPs.
    pooled = bool(lib.pcg_set_pooled_last_block(1))
    lib.pcg_set_pooled_last_block(int(pooled))
    full_last = None
    if pooled and world == 1 and not args.no_full_last_block:
        lib.pcg_set_pooled_last_block(0)
        eng._slot = eng._slot_key = None  # captured graphs keep the sequence they were captured with
        try:
            for _ in range(3):
                step_device()
            f_ms = timed(step_device, args.steps) / args.steps
        finally:
            lib.pcg_set_pooled_last_block(1)
            eng._slot = eng._slot_key = None
        full_last = {"ms_per_step": f_ms, "value": cutouts_per_step / (f_ms * 1e-3), "unit": "cutouts/s",
                     "steps": args.steps,
                     "note": "PCG_FULL_LAST_BLOCK=1: all n*T rows through the last block's attention, out-projection, "
                             "ln_2 and MLP (what the reference computes and then discards); same loss and gradient"}

    # strong-scaling sub-record of the headline workload (C3 at B = 1): ONE image, the same 128 cutouts
    # split over the ranks, gradient all-reduced.  Every rank runs it; at N = 1 it is the main measurement itself.
    strong = None
    if world > 1 and args.workload == DEFAULT_WORKLOAD and not args.no_strong:
        s_arch, s_hw, s_cut, s_min, _ = WORKLOADS["vit_l14_224_128cut_512px_1image"]
        strong_mod = losses.CLIP(s_arch, n_cutouts=s_cut, min_size=s_min, max_size=s_hw, seed=0, process_group=group,
                                 shard="cutouts")
        strong_mod.add_encodings_(target_enc)
        one_image = torch.rand(1, 3, s_hw, s_hw, generator=torch.Generator().manual_seed(1)).to(device)

        def step_strong():
            img = one_image.detach().requires_grad_()
            loss = strong_mod(img)
            loss.backward()
            return loss

        for _ in range(max(args.warmup, 3)):
            step_strong()
        s_ms = timed(step_strong, args.steps) / args.steps
        strong = {"workload": "vit_l14_224_128cut_512px_1image", "scaling": "strong", "n_gpus": world,
                  "cutouts_per_step": s_cut, "cutouts_per_gpu": s_cut // world, "ms_per_step": s_ms,
                  "value": s_cut / (s_ms * 1e-3), "unit": "cutouts/s",
                  "note": "one 512x512 image, 128 cutouts sharded over the ranks, image gradient + loss all-reduced; "
                          "device-timed, max over ranks, CUDA-graph replay"}

    # The other BASELINE.json configs on ONE GPU, as short sub-records beside the headline (they are parity-test cases
    # first; these lines only put a driver-run number next to each): configs[0] shape, configs[1], configs[3] shape.
    others = None
    if world == 1 and args.workload == DEFAULT_WORKLOAD and not args.no_other_configs:
        others = {}
        del loss_mod, eng  # frees the 19 GB activation slot of the headline workload (the encoder itself is memoised)
        torch.cuda.empty_cache()
        for tag, wl, o_steps in (("configs[0] shape on the GPU", "vit_b32_224_16cut_256px", 20),
                                 ("configs[1]", "vit_b32_224_64cut_4x512px", 20),
                                 ("configs[3] shape on one GPU", "vit_l14_336_256cut_768px", 3)):
            o_arch, o_hw, o_cut, o_min, o_imgs = WORKLOADS[wl]
            o_mod = losses.CLIP(o_arch, n_cutouts=o_cut, min_size=o_min, max_size=o_hw, seed=0)
            o_mod.add_encodings_(torch.randn(2, SHAPES[o_arch].embed, generator=torch.Generator().manual_seed(0)))
            o_mod.model.engine().prebuild_tables(o_min, o_hw)
            o_images = torch.rand(o_imgs, 3, o_hw, o_hw, generator=torch.Generator().manual_seed(0)).to(device)

            def step_other():
                img = o_images.detach().requires_grad_()
                loss = o_mod(img)
                loss.backward()
                return loss

            for _ in range(3):
                step_other()
            o_ms = timed(step_other, o_steps) / o_steps
            others[wl] = {"baseline_config": tag, "model": o_arch, "images": o_imgs, "image": f"{o_hw}x{o_hw}",
                          "cutouts_per_step": o_imgs * o_cut, "steps": o_steps, "ms_per_step": o_ms,
                          "value": o_imgs * o_cut / (o_ms * 1e-3), "unit": "cutouts/s",
                          "frac_of_sustained_peak": o_imgs * o_cut / (o_ms * 1e-3)
                          * SHAPES[o_arch].flops_per_cutout(pooled_last_block=pooled) / 1e12 / measured_peaks()["bf16_tflops_sustained"]}
            del o_mod, o_images
            torch.cuda.empty_cache()
        loss_mod = losses.CLIP(arch, n_cutouts=n_cut, min_size=min_size, max_size=hw, seed=0)  # for the parity field
        loss_mod.add_encodings_(target_enc)

    if rank != 0:
        return
    parity = parity_check(loss_mod, arch, shape, hw, min_size, device) if not args.no_parity else None
    peaks = measured_peaks()
    flops_per_cutout = shape.flops_per_cutout(pooled_last_block=pooled)  # what the step executes
    step_tflops = value / world * flops_per_cutout / 1e12
    model_tflops = value / world * shape.flops_per_cutout() / 1e12       # the reference's dense op count at this speed
    gemm = prof["gemm"]
    gemm_tflops = gemm["work"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    peak = peaks["bf16_tflops_sustained"]
    families = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["count"] / args.steps,
                    "rate": (v["work"] / (v["ms"] * 1e-3) / (1e12 if k in ("gemm", "attn_fwd", "attn_bwd") else 1e9))
                    if v["ms"] > 0 else 0.0} for k, v in prof.items()}

    traffic, traffic_src = ncu_gemm_traffic(args.workload)
    cpu_value, cpu_med, cpu_done = time_cpu(arch, hw, min_size, args.cpu_sample, 2, 1) if not args.no_cpu else (None, None, 0)
    line = {
        "metric": METRIC, "value": value, "unit": "cutouts/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak" if imgs_per_gpu > 0 else "strong",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": args.workload, "model": arch, "images": n_images, "image": f"{hw}x{hw}",
                   "cutouts_per_image": n_cut, "cutouts_per_step": cutouts_per_step, "targets": 2,
                   "parallelism": (f"image-sharded dp{world} (each rank its own image, no gradient collective)"
                                   if shard == "images" else f"cutout-sharded dp{world} (gradient all-reduce)")
                   if world > 1 else "single GPU",
                   "weights": "random-init (no network for checkpoints)",
                   "l2": "per-step working set (activation stash >= 19 GB for ViT-L/14 x128) >> 126 MB L2; no flush needed",
                   "launch": "CUDA-graph replay (forward graph + backward graph)" if graphs_on else "eager launches",
                   "last_block": ("class-token rows only after the K/V projection (the loss reads nothing else); "
                                  "FLOPs below are the EXECUTED ones" if pooled else "full"),
                   "step_tflops_per_gpu": step_tflops, "step_frac_of_peak": step_tflops / peak,
                   "step_dense_model_tflops_per_gpu": model_tflops,
                   "kernel_families": families},
        "clocks": clock_info,
        "e2e": {"value": e2e_value, "unit": "cutouts/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": host_images.numel() * 4 + cutouts_per_step // world * 32,  # per rank
                "d2h_bytes_per_step": host_grad.numel() * 4 + 4},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": gemm_tflops, "peak": peak,
                     "unit": "TFLOP/s", "frac": gemm_tflops / peak,
                     "traffic": traffic,
                     "traffic_unit": f"bytes per launch (ncu dram read + write; {traffic_src})",
                     "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                     "launches_timed": gemm["count"], "share_of_step": gemm["ms"] / prof_ms,
                     "timing": "CUDA event pair around every launch, second pass over the same K steps "
                               f"({prof_ms / args.steps:.2f} ms/step with the events, eager launches)"},
    }
    if full_last is not None:
        line["full_last_block"] = full_last
    if strong is not None:
        line["strong"] = strong
    if others:
        line["other_configs"] = others
    if parity is not None:
        line["parity"] = parity
    if cpu_value is not None:
        line["cpu_baseline"] = {"value": cpu_value, "unit": "cutouts/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{args.cpu_sample} cutouts/step of the same workload, {cpu_done} timed steps "
                                          f"(median {cpu_med:.2f} s), fp32 torch CPU oracle, {os.cpu_count()} host cpus"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default=DEFAULT_WORKLOAD)
    ap.add_argument("--cpu-sample", type=int, default=8, help="cutouts per CPU-baseline step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling sub-record (N > 1)")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the short sub-records of the other BASELINE configs (N = 1, default workload)")
    ap.add_argument("--no-full-last-block", action="store_true",
                    help="skip the sub-record with the last block run on all tokens (N = 1)")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity check against the CPU oracle")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("bench.py: --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
