#!/usr/bin/env python
"""Benchmark of the diffusion-planning hot path (BASELINE.json metric: trajectories/sec of the full DDIM/DDPM loop).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one complete plan: T denoising iterations (denoiser + scheduler step + waypoint overwrite) plus the
final clamp/scale for a batch of B trajectories per GPU.  Default workload = BASELINE.json configs[1]:
configs/default.yaml (NO_GUIDANCE), GuidanceDDIMScheduler, EVAL.SAMPLE_STEPS = 100, B = 256 per GPU, precomputed
image feature [B,64] (the encoder is hoisted out of the loop, SURVEY.md §8d "loop-only").  Weak scaling: every rank
plans its own 256 trajectories, no collective on the data path; torch.distributed is only used for the timing barrier
and the max-over-ranks reduction.

`--impl reference` times the reference's algorithm on the host CPU cores (oracle port of the reference PyTorch code:
/root/reference does not exist on the GPU box and diffusers is not installed) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOPS_PER_EVAL = {"NO_GUIDANCE": 78_874_624, "FREE_GUIDANCE": 78_883_072, "CLASSIFIER_GUIDANCE": 80_845_952}
SCHED = {"ddim": "guidance_ddim", "ddpm": "guidance_ddpm", "inpainting_ddim": "inpainting_ddim", "inpainting_ddpm": "inpainting_ddpm"}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="trajectories per GPU")
    ap.add_argument("--timesteps", type=int, default=100, help="EVAL.SAMPLE_STEPS")
    ap.add_argument("--sched", default="ddim", choices=list(SCHED))
    ap.add_argument("--mode", default="NO_GUIDANCE", choices=list(FLOPS_PER_EVAL))
    ap.add_argument("--precision", default="bf16x3", choices=["fp32", "bf16x3", "bf16"],
                    help="bf16x3 = tcgen05 with bf16 hi/lo split operands (3 MMAs, fp32 accumulate): meets the fp32 parity bound (<=1e-3)")
    ap.add_argument("--no-other-precisions", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the per-mode report (SURVEY.md 8d configs 1, 3, 4) and the encoder figure")
    ap.add_argument("--cpu-sample-iters", type=int, default=5, help="denoising iterations per timed CPU sample")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d["bf16_tflops_sustained"], tensor_burst=d["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor=1400.0, tensor_burst=1590.0, source="fallback (B200_PROFILING.md)")


def b1_roofline(p50_ms: float, nfe: int, weight_bytes: int, pk: dict) -> dict:
    """Batch-1 bound: every denoiser evaluation streams the fp32 weights once (the GEMV path computes in
    exact fp32 whatever the precision mode); achieved = weight bytes x evaluations / p50 plan latency."""
    achieved = weight_bytes * nfe / (p50_ms * 1e-3) / 1e9
    return {"bound": "hbm", "achieved": achieved, "peak": pk["hbm"], "unit": "GB/s", "frac": achieved / pk["hbm"],
            "algorithmic_bytes_per_evaluation": weight_bytes, "evaluations_per_plan": nfe, "peak_source": pk["source"] + ", copy bandwidth",
            "note": "weights (fp32, K-major) are L2-resident across the iterations, so the HBM copy peak is the conservative denominator; "
                    "the plan is a chain of ~43 dependent launches per iteration at ~3.9 us each (profiles/r01_gemv_stage_trace_b1.txt), "
                    "latency-bound, not bandwidth-bound"}


def workload_name(a):
    return f"configs/default.yaml {a.mode} {SCHED[a.sched]} T={a.timesteps} B={a.batch}/GPU precomputed-feature (BASELINE.json configs[1])"


# --------------------------------------------------------------------------------------------------------------
# CPU side (oracle port of the reference): used for cpu_baseline and for --impl reference
# --------------------------------------------------------------------------------------------------------------
def cpu_iterations(a, n_iters: int, B: int):
    """Time n_iters denoising iterations of the reference algorithm on the host cores; returns seconds per iteration."""
    from oracle import schedulers as S
    from oracle import unet as U
    from oracle import weights as W

    if not hasattr(cpu_iterations, "_state"):
        sd = W.make_state_dict(a.mode, seed=0, with_perception=False)
        inp = W.synth_inputs(B, a.timesteps if a.sched != "ddim" else 0, seed=1)
        cpu_iterations._state = (sd, inp, S.alphas_cumprod(100), S.SchedCfg(num_inference_steps=a.timesteps))
    sd, inp, ac, cfg = cpu_iterations._state
    x = inp["x"].clone()
    ts = S.leading_timesteps(100, a.timesteps)[:n_iters]
    t0 = time.perf_counter()
    with torch.no_grad():
        for i, t in enumerate(ts):
            t = int(t)
            tt = torch.full((B,), t, dtype=torch.long)
            if a.mode == "FREE_GUIDANCE":
                cond = torch.cat([inp["target"], torch.zeros_like(inp["target"])], 0)
                c, u = U.unet_forward(sd, torch.cat([x, x], 0), inp["feat"], torch.tensor([t]), cond, a.mode).chunk(2, 0)
                mo = u + 7.5 * (c - u)
            else:
                mo = U.unet_forward(sd, x, inp["feat"], tt, None, a.mode)
            if a.sched.endswith("ddim"):
                x, _ = S.ddim_step(cfg, ac, mo, t, x, inpainting=a.sched.startswith("inpainting"))
            else:
                x, _ = S.ddpm_step(cfg, ac, mo, t, x, variance_noise=inp["noise"][i], inpainting=a.sched.startswith("inpainting"))
            x[:, 0, :3] = 0.0
    return (time.perf_counter() - t0) / len(ts)


def cpu_config0(reps: int = 3):
    """BASELINE.json configs[0]: default.yaml, no guidance, DDPM, T=100, batch 1, on the host cores (whole plans, loop only:
    the image feature is precomputed as on the GPU arm)."""
    from oracle import plan as OP
    from oracle import weights as W

    sd = W.make_state_dict("NO_GUIDANCE", seed=0, with_perception=False)
    inp = W.synth_inputs(1, 100, seed=1)
    run = lambda: OP.plan(sd, "NO_GUIDANCE", "guidance_ddpm", inp["x"], inp["feat"], 100, noise=inp["noise"])  # noqa: E731
    run()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    ms = statistics.median(ts) * 1e3
    return {"ms_per_plan_p50": ms, "traj_per_s": 1e3 / ms, "batch": 1, "T": 100, "sched": "guidance_ddpm", "cores": torch.get_num_threads(),
            "sample": f"{reps} whole plans (oracle port, fp32, precomputed feature)"}


def run_reference_arm(a, rank: int):
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    B, T, n = a.batch, a.timesteps, max(1, min(a.cpu_sample_iters, a.timesteps))
    for _ in range(a.warmup):
        cpu_iterations(a, 1, B)
    per_iter = [cpu_iterations(a, n, B) for _ in range(a.steps)]
    plan_s = statistics.mean(per_iter) * T
    value = B / plan_s
    sample = (f"each step = {n} of the {T} denoising iterations at B={B} (every iteration runs the same denoiser + scheduler step), "
              f"extrapolated x{T}/{n}; oracle port of the reference PyTorch code, fp32, torch CPU {torch.get_num_threads()} threads")
    line = {"impl": "reference", "metric": "trajectories_per_sec_full_sampling_loop", "value": value, "unit": "trajectories/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": plan_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "l2": "n/a (CPU)"},
            "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------
# GPU side
# --------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.f, self.p = index, None, None

    def start(self):
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return out
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        out.update(sm_mhz=statistics.median(sm), sm_max_mhz=float(rows[0][2]), reasons=sorted(reasons), samples=len(rows),
                   power_w_max=max(float(r[3]) for r in rows))
        return out


def mode_report(P, W, dev, B: int, precision: str):
    """throughput at B trajectories and batch-1 p50 plan latency for the other BASELINE.json configs, plus the
    end-to-end figure with the image encoder in front of the loop (rank 0 only, a few plans each; report-only numbers)."""
    cases = {
        "config1_noguidance_ddpm100": ("NO_GUIDANCE", "guidance_ddpm", 100),
        "config2_noguidance_ddim10": ("NO_GUIDANCE", "guidance_ddim", 10),
        "config3_cfg_ddim10_scale7.5": ("FREE_GUIDANCE", "guidance_ddim", 10),
        "config4a_classifier_ddim2_scale15": ("CLASSIFIER_GUIDANCE", "guidance_ddim", 2),
        "config4b_classifier_inpainting_ddim2": ("CLASSIFIER_GUIDANCE", "inpainting_ddim", 2),
    }
    classes = {"guidance_ddim": P.GuidanceDDIMScheduler, "guidance_ddpm": P.GuidanceDDPMScheduler,
               "inpainting_ddim": P.InpaintingDDIMScheduler, "inpainting_ddpm": P.InpaintingDDPMScheduler}
    models, out = {}, {}
    for name, (mode, kind, T) in cases.items():
        cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION=precision),
                         GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0,
                                       LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
        if mode not in models:
            m = P.build_model(cfg)
            m.load_state_dict(W.make_state_dict(mode, seed=0))
            models[mode] = m.to(dev).eval()
        sched = classes[kind](cfg=cfg, **P.scheduler_kwargs(cfg)) if kind.startswith("guidance") else classes[kind](**P.scheduler_kwargs(cfg))
        planner = P.DiffusionPlanner(models[mode], sched, cfg)
        inp = W.synth_inputs(B, T, seed=2)
        needs_noise, inpaint = kind != "guidance_ddim", kind.startswith("inpainting")
        dd = dict(target=inp["target"].to(dev) if mode != "NO_GUIDANCE" and not inpaint else None,
                  noise=inp["noise"].to(dev) if needs_noise else None,
                  target_traj=inp["target_traj"].to(dev) if inpaint else None, target_mask=inp["mask"].to(dev) if inpaint else None)
        x, f = inp["x"].to(dev), inp["feat"].to(dev)
        one = {k: (None if v is None else (v[:, :1] if k == "noise" else v[:1]).contiguous()) for k, v in dd.items()}
        for _ in range(2):
            planner.plan(x, f, **dd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            planner.plan(x, f, **dd)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        for _ in range(3):
            planner.plan(x[:1], f[:1], **one)
        lat = []
        for _ in range(20):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            planner.plan(x[:1], f[:1], **one)
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
        out[name] = {"traj_per_s": B / (ms * 1e-3), "ms_per_plan": ms, "batch": B, "latency_b1_p50_ms": statistics.median(lat),
                     "launches_per_plan": planner.last_launch_count()}
    # large-batch regime: the same kernels when a layer has hundreds of row tiles instead of 4..32
    try:
        mode, kind, T, BL = "NO_GUIDANCE", "guidance_ddim", 10, 4096
        cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION=precision))
        planner = P.DiffusionPlanner(models[mode], P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
        g = torch.Generator(device=dev).manual_seed(5)
        x, f = torch.randn(BL, 16, 7, device=dev, generator=g), torch.randn(BL, 64, device=dev, generator=g)
        for _ in range(2):
            planner.plan(x, f)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            planner.plan(x, f)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tf = FLOPS_PER_EVAL[mode] * BL * T / (ms * 1e-3) / 1e12
        out["large_batch_ddim10"] = {"batch": BL, "traj_per_s": BL / (ms * 1e-3), "ms_per_plan": ms, "nominal_tflops": tf,
                                     "frac_of_bf16_sustained_peak": tf / peaks()["tensor"]}
        del x, f
    except Exception as exc:
        out["large_batch_error"] = repr(exc)[:200]
    # end to end with the image encoder: one ResNet-34 pass per distinct scene (hoisted out of the loop), then the DDIM-100 loop
    try:
        mode, kind, T = "NO_GUIDANCE", "guidance_ddim", 100
        cfg = P.load_cfg(EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION=precision))
        m = P.build_model(cfg)
        m.load_state_dict(W.make_state_dict(mode, seed=0, with_perception=True))
        m = m.to(dev).eval()
        planner = P.DiffusionPlanner(m, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
        for scenes in sorted({min(16, B), B}):
            img = torch.randn(scenes, 3, 256, 900, device=dev, generator=torch.Generator(device=dev).manual_seed(2))   # ImageNet-normalised frames
            x = W.synth_inputs(B, 0, seed=2)["x"].to(dev)
            rep = B // scenes

            def run():
                with torch.no_grad():
                    feat = m.perception(img)
                return planner.plan(x, feat.repeat_interleave(rep, 0) if rep > 1 else feat)

            def enc_only():
                with torch.no_grad():
                    return m.perception(img)

            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            for _ in range(3):
                enc_only()
            e1.record()
            for _ in range(3):
                run()
            e2.record()
            torch.cuda.synchronize()
            out[f"with_encoder_{scenes}_scenes"] = {"traj_per_s": B / (e1.elapsed_time(e2) / 3 * 1e-3), "ms_per_plan": e1.elapsed_time(e2) / 3,
                                                    "encoder_ms": e0.elapsed_time(e1) / 3, "batch": B, "image": "3x256x900 fp32 per scene",
                                                    "encoder": "ResNet-34 on torch/cuDNN (library code, SURVEY 8f rank 1), one pass per scene, hoisted"}
            del img
        # closed-loop tick at batch 1: a NEW camera frame every tick (interact.py:170-176 -> generate_traj), encoder included
        frames = [torch.randn(1, 3, 256, 900, device=dev) for _ in range(4)]
        for name, (mode, kind, T) in (("tick_noguidance_ddim100", ("NO_GUIDANCE", "guidance_ddim", 100)), ("tick_cfg_ddim10", cases["config3_cfg_ddim10_scale7.5"]),
                                      ("tick_classifier_ddim2", cases["config4a_classifier_ddim2_scale15"])):
            cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION=precision),
                             GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0,
                                           LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
            mm = P.build_model(cfg)
            mm.load_state_dict(W.make_state_dict(mode, seed=0, with_perception=True))
            mm = mm.to(dev).eval()
            agent = P.DiffusionPlanner(mm, P.GuidanceDDIMScheduler(cfg=cfg, **P.scheduler_kwargs(cfg)), cfg)
            tg = torch.tensor([[0.1, 0.3]], device=dev) if mode != "NO_GUIDANCE" else None
            for i in range(4):
                agent.generate_traj(frames[i % 4], tg)
            lat = []
            for i in range(20):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                agent.generate_traj(frames[i % 4], tg)
                torch.cuda.synchronize()
                lat.append((time.perf_counter() - t0) * 1e3)
            out[name] = {"closed_loop_tick_p50_ms": statistics.median(lat), "batch": 1, "T": T,
                         "includes": "ResNet-34 encoder on the new frame (cuDNN, folded BN, one CUDA graph) + the whole sampling loop + host call"}
    except Exception as exc:  # report-only: never fail the bench line on the 'next' row
        out["with_encoder_error"] = repr(exc)[:200]
    return out


def run_b200_arm(a, rank: int, world: int, local_rank: int):
    import autonomous_driving_with_diffusion_model_b200 as P
    from oracle import weights as W  # deterministic synthetic weights/inputs only (not the checker, not timed)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the product path has no CPU fallback")
    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    B, T = a.batch, a.timesteps
    mode = a.mode
    cfg = P.load_cfg(TRAIN=dict(USE_COND=mode), EVAL=dict(SAMPLE_STEPS=T), B200=dict(PRECISION=a.precision),
                     GUIDANCE=dict(USE_COND=mode, FREE_SCALE=7.5, CLASSIFIER_SCALE=15.0,
                                   LOSS_LIST=[["TargetGuidance", []]] if mode == "CLASSIFIER_GUIDANCE" else None))
    model = P.build_model(cfg)
    sd = W.make_state_dict(mode, seed=0)
    weight_bytes = 4 * sum(int(v.numel()) for k, v in sd.items() if not k.startswith("perception.") and "num_batches_tracked" not in k)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    kind = SCHED[a.sched]
    cls = {"guidance_ddim": P.GuidanceDDIMScheduler, "guidance_ddpm": P.GuidanceDDPMScheduler,
           "inpainting_ddim": P.InpaintingDDIMScheduler, "inpainting_ddpm": P.InpaintingDDPMScheduler}[kind]
    sched = cls(cfg=cfg, **P.scheduler_kwargs(cfg)) if kind.startswith("guidance") else cls(**P.scheduler_kwargs(cfg))
    planner = P.DiffusionPlanner(model, sched, cfg)

    # global synthetic batch, sharded by rank (weak scaling: B per GPU)
    inp = W.synth_inputs(B * world, T if kind != "guidance_ddim" else 0, seed=1)
    sh = lambda t, d=0: None if t is None else P.shard(t, rank, world, d).contiguous()  # noqa: E731
    needs_noise = kind != "guidance_ddim"
    inpaint = kind.startswith("inpainting")
    host = dict(x=sh(inp["x"]).pin_memory(), feat=sh(inp["feat"]).pin_memory(),
                target=sh(inp["target"]).pin_memory() if mode != "NO_GUIDANCE" else None,
                noise=sh(inp["noise"], 1).pin_memory() if needs_noise else None,
                traj=sh(inp["target_traj"]).pin_memory() if inpaint else None, mask=sh(inp["mask"]).pin_memory() if inpaint else None)
    d = {k: (None if v is None else v.to(dev)) for k, v in host.items()}
    call = lambda: planner.plan(d["x"], d["feat"], target=d["target"], noise=d["noise"], target_traj=d["traj"], target_mask=d["mask"])  # noqa: E731
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 3)):
        out = call()
    barrier()
    launches_per_step = planner.last_launch_count()
    stream = torch.cuda.current_stream()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    for e0, e1 in evs:
        flush.fill_(1)          # L2 flush between timed iterations (not timed)
        e0.record(stream)
        out = call()
        e1.record(stream)
    barrier()
    clocks = sampler.stop()
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = sum(step_ms)

    # end to end through the public API with HOST buffers: H2D of the inputs and D2H of the trajectories every step
    e2e_out = torch.empty_like(host["x"]).pin_memory()
    e2e_call = lambda: planner.plan_host(host["x"], host["feat"], target=host["target"], noise=host["noise"], target_traj=host["traj"],  # noqa: E731
                                         target_mask=host["mask"], out=e2e_out, device=dev)
    for _ in range(3):
        e2e_call()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_call()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert torch.equal(e2e_out, out.cpu()), "host-entry result differs from the device-entry result"
    h2d = sum(v.numel() * 4 for v in host.values() if v is not None)
    d2h = e2e_out.numel() * 4

    # dominant kernel (fused conv block) timed live: one eager denoiser evaluation = its 41 conv launches + 2 small ones
    tt = torch.full((B,), 50, dtype=torch.long, device=dev)
    n_eval = 20
    xin = d["x"] if mode != "FREE_GUIDANCE" else torch.cat([d["x"], d["x"]], 0)
    fwd = (lambda: model(xin, d["feat"], tt[:1], cond=torch.cat([d["target"], torch.zeros_like(d["target"])], 0))) if mode == "FREE_GUIDANCE" else \
          (lambda: model(xin, d["feat"], tt, return_action_and_time_only=(mode == "CLASSIFIER_GUIDANCE")))
    for _ in range(3):
        fwd()
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record(stream)
    for _ in range(n_eval):
        fwd()
    k1.record(stream)
    torch.cuda.synchronize()
    eval_ms = k0.elapsed_time(k1) / n_eval
    eval_launches = model.last_launch_count()

    # batch-1 plan latency (p50), same scheduler / T
    one = {k: (None if v is None else (v[:, :1] if k == "noise" else v[:1]).contiguous()) for k, v in d.items()}
    lat_call = lambda: planner.plan(one["x"], one["feat"], target=one["target"], noise=one["noise"], target_traj=one["traj"], target_mask=one["mask"])  # noqa: E731
    for _ in range(3):
        lat_call()
    lat = []
    for _ in range(30):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lat_call()
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)

    # the other arithmetic modes on the same workload (3 plans each, device-resident inputs), for the report only
    others = {}
    if not a.no_other_precisions:
        for prec in ("fp32", "bf16x3", "bf16"):
            if prec == a.precision:
                continue
            model.set_precision(prec)
            for _ in range(2):
                call()
            torch.cuda.synchronize()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record(stream)
            for _ in range(3):
                call()
            o1.record(stream)
            torch.cuda.synchronize()
            others[prec] = B * world * 3 / (o0.elapsed_time(o1) * 1e-3)
        model.set_precision(a.precision)

    modes = None
    if rank == 0 and not a.no_modes:
        modes = mode_report(P, W, dev, B, a.precision)

    # max over ranks
    tot = torch.tensor([total_ms, e2e_s], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(tot[0]), float(tot[1])

    cpu = None
    if rank == 0 and not a.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        n = max(1, min(a.cpu_sample_iters, T))
        cpu_iterations(a, 1, B)
        reps, t_begin, per = 0, time.perf_counter(), []
        while reps < 3 or (time.perf_counter() - t_begin < 12 and reps < 40):
            per.append(cpu_iterations(a, n, B))
            reps += 1
        cpu = {"value": B / (statistics.mean(per) * T), "unit": "trajectories/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{reps} x {n} of the {T} denoising iterations at B={B} (oracle port of the reference PyTorch code, fp32), extrapolated x{T}/{n}"}
        try:
            cpu["config0_b1_ddpm100"] = cpu_config0()
        except Exception as exc:  # report-only figure: never lose the bench line over it
            cpu["config0_b1_ddpm100"] = {"error": repr(exc)}

    if rank == 0:
        pk = peaks()
        nfe = T * (2 if mode == "FREE_GUIDANCE" else 1)
        rows = B * (2 if mode == "FREE_GUIDANCE" else 1)
        flops_eval = FLOPS_PER_EVAL[mode] * rows
        # dominant kernel = the fused conv layer kernel (>95 % of the step, profiles/r01_launches_*): its algorithmic FLOPs over the
        # timed region divided by the timed region itself (conservative: the scheduler launches are inside the denominator)
        achieved = flops_eval * T / (total_ms / a.steps * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(f"{a.precision}:{mode}:B{B}")
        line = {
            "metric": "trajectories_per_sec_full_sampling_loop", "value": B * world * a.steps / (total_ms * 1e-3), "unit": "trajectories/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": total_ms / a.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16x3": "bf16x3(f32-accum)", "bf16": "bf16(f32-accum)"}[a.precision],
            "data": "synthetic",
            "config": {"workload": workload_name(a), "global_batch": B * world, "sharding": f"batch/{world} contiguous, no collective on the data path",
                       "l2": "256 MiB flush between timed iterations", "weights": "random-init (hash RNG), reference state_dict layout",
                       "nfe_per_plan": nfe, "cuda_graph": True},
            "e2e": {"value": B * world * a.steps / e2e_s, "unit": "trajectories/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "entry": "DiffusionPlanner.plan_host -> b2p_plan_host (pinned host buffers, wall clock incl. sync)"},
            "gpu_launches": int(launches_per_step * a.steps),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tensor"], "unit": "TFLOP/s", "frac": achieved / pk["tensor"],
                         "traffic": traffic, "peak_source": pk["source"] + ", bf16 sustained",
                         "kernel": "fused conv block (conv_ffma_kernel, CUDA cores)" if a.precision == "fp32" else "fused conv block (conv_tc_kernel: TMA + tcgen05.mma + TMEM epilogue)",
                         "note": "B=256/GPU: ~40 dependent layer launches per denoising iteration, each <=128 CTAs of 128 rows x 16 channels; per layer ~1.15 us dependency release + 0.8 us first-operand latency + a main loop of narrow (N<=80) MMAs paced by shared-memory operand reads + a TMEM-read-bound epilogue (profiles/r01_tc_stage_trace_b256.txt); see scripts/sweep.py for the large-batch regime",
                         "how": f"algorithmic FLOPs ({FLOPS_PER_EVAL[mode]} nominal 2*MAC x {rows} rows x {T} evaluations per plan) / CUDA-event time of the plan "
                                f"(timed region, CUDA graph replay)",
                         "eager_eval": {"tflops": flops_eval / (eval_ms * 1e-3) / 1e12, "us": eval_ms * 1e3, "launches": eval_launches,
                                        "note": "one denoiser evaluation launched eagerly (no graph), CUDA events, avg of %d" % n_eval}},
            "latency_b1": {"p50_ms": statistics.median(lat), "p95_ms": sorted(lat)[int(0.95 * len(lat)) - 1], "T": T, "sched": kind,
                           "roofline": b1_roofline(statistics.median(lat), nfe, weight_bytes, pk)},
            "precision": {"mode": a.precision, "parity_bound_max_abs": {"fp32": 1e-3, "bf16x3": 1e-3, "bf16": 0.3}[a.precision],
                          "other_modes_traj_per_s_rank0_x_world": others},
        }
        if modes is not None:
            line["modes"] = modes
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank)
        return
    if world != a.gpus and world == 1 and a.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={a.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29541"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200_arm(a, rank, world, local_rank)


if __name__ == "__main__":
    main()
