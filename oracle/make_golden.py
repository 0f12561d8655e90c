"""Generate tests/golden/*.npz by running the REAL reference code.  BUILD-CONTAINER ONLY.

    python -m oracle.make_golden            # writes tests/golden/, prints oracle-vs-reference differences

What runs: the reference's own ``TemporalMapUnet`` (``modeling/temporal.py``), its four scheduler classes
(``scheduler/*.py`` step() bodies verbatim, on the diffusers base-class shim), ``GuidanceLoss``/``TargetGuidance``
(``control/``) driven by a loop that follows ``interact.py:115-168`` statement by statement.  Inputs and weights come
from ``oracle.weights`` (hash-based, reproducible anywhere), so the golden files only hold OUTPUTS plus a digest of
the weights they were made with.  The same cases are replayed through ``oracle.plan`` and the max-abs difference is
stored in the file (``oracle_vs_reference``) — this is the pin of the restatement against the reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import plan as P  # noqa: E402
from oracle import reference_loader as RL  # noqa: E402
from oracle import schedulers as S  # noqa: E402
from oracle import unet as U  # noqa: E402
from oracle import weights as W  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (mode, scheduler kind, T, B, seed)
PLAN_CASES = {
    "cfg1_noguid_ddpm100_b1": ("NO_GUIDANCE", "guidance_ddpm", 100, 1, 11),
    "cfg2_noguid_ddim100_b1": ("NO_GUIDANCE", "guidance_ddim", 100, 1, 12),
    "cfg2_noguid_ddim10_b4": ("NO_GUIDANCE", "guidance_ddim", 10, 4, 13),
    "cfg3_free_ddim10_b1": ("FREE_GUIDANCE", "guidance_ddim", 10, 1, 14),
    "cfg3_free_ddim10_b3": ("FREE_GUIDANCE", "guidance_ddim", 10, 3, 15),
    "cfg3_free_ddpm10_b2": ("FREE_GUIDANCE", "guidance_ddpm", 10, 2, 16),
    "cfg4_classifier_ddim2_b3": ("CLASSIFIER_GUIDANCE", "guidance_ddim", 2, 3, 17),
    "cfg4_classifier_ddpm2_b2": ("CLASSIFIER_GUIDANCE", "guidance_ddpm", 2, 2, 18),
    "cfg4_classifier_ddim10_b2": ("CLASSIFIER_GUIDANCE", "guidance_ddim", 10, 2, 19),
    "cfg4b_inpaint_ddim10_b2": ("NO_GUIDANCE", "inpainting_ddim", 10, 2, 20),
    "cfg4b_inpaint_ddpm10_b2": ("NO_GUIDANCE", "inpainting_ddpm", 10, 2, 21),
    "cfg4b_classifier_inpaint_ddim2_b2": ("CLASSIFIER_GUIDANCE", "inpainting_ddim", 2, 2, 22),
}


class _NoiseFeeder:
    def __init__(self, noise):
        self.noise, self.i = noise, 0

    def __call__(self, shape, generator=None, device=None, dtype=None, layout=None):
        n = self.noise[self.i]
        assert tuple(n.shape) == tuple(shape)
        return n.clone()


def reference_generate_traj(model, sched, mode, x_init, feat, T, target=None, noise=None, free_scale=7.5,
                            target_traj=None, target_mask=None, trace=None):
    """interact.py:115-168 driven on the reference objects.  ``feat`` is fed through an Identity ``perception``."""
    ref = RL.load()
    GT = ref.GuidanceType
    use = GT[mode]
    inpaint = type(sched).__name__.startswith("Inpainting")
    trajs = x_init.clone().detach()
    B = trajs.shape[0]
    if target is not None and use == GT.FREE_GUIDANCE:
        target = torch.cat([target, torch.zeros_like(target)], dim=0)
    trajs[:, 0, :3] = 0.0
    sched.set_timesteps(T, device="cpu")
    feeder = _NoiseFeeder(noise) if noise is not None else None
    import scheduler.guidance_ddpm_scheduler as m1
    import scheduler.inpainting_ddim_scheduler as m2
    import scheduler.inpainting_ddpm_scheduler as m3
    import scheduler.guidance_ddim_scheduler as m4
    for m in (m1, m2, m3, m4):
        if feeder is not None:
            m.randn_tensor = feeder
    action = None
    for i, t in enumerate(sched.timesteps):
        if feeder is not None:
            feeder.i = i
        if use == GT.FREE_GUIDANCE:
            with torch.no_grad():
                c, u = model(torch.cat([trajs, trajs], 0), feat, t.reshape(-1), cond=target).chunk(2, dim=0)
            mo = u + free_scale * (c - u)
        else:
            tt = t.reshape(-1).repeat(B)  # train.py:85 (the reference's batched call)
            with torch.no_grad():
                mo = model(trajs, feat, tt, return_action_and_time_only=(use == GT.CLASSIFIER_GUIDANCE))
        if use == GT.CLASSIFIER_GUIDANCE:
            action, te = mo
            action = action.detach()
            action.requires_grad_()
            state = model.state_pred(action[:, :-1], te)
            state = torch.cat([torch.zeros_like(state[:, :1]), state], dim=1)
            mo = torch.cat([state, action], dim=-1)
        if inpaint:
            out = sched.step(mo.detach(), t, trajs, target_traj=target_traj, target_mask=target_mask)
        else:
            out = sched.step(mo, t, trajs, target=target, action=action)
        trajs = out.prev_sample.detach()
        trajs[:, 0, :3] = 0.0
        if trace is not None:
            trace.append(dict(t=int(t), prev_sample=trajs.clone()))
    trajs = trajs.to(torch.float32).clamp(-1, 1)
    trajs[..., :2] *= model.magic_num
    return trajs


def run_plan_case(name, models, sds):
    mode, kind, T, B, seed = PLAN_CASES[name]
    inp = W.synth_inputs(B, T, seed)
    model = models[mode]
    ddpm_or_inpaint = kind.endswith("ddpm") or kind.startswith("inpainting")
    noise = inp["noise"] if ddpm_or_inpaint else None
    target = inp["target"] if mode != "NO_GUIDANCE" else None
    tt = inp["target_traj"] if kind.startswith("inpainting") else None
    tm = inp["mask"] if kind.startswith("inpainting") else None
    if mode == "CLASSIFIER_GUIDANCE" and not kind.startswith("inpainting"):
        # reference semantics are B=1 only: run sample by sample
        outs = []
        for b in range(B):
            sched = RL.build_reference_scheduler(kind, mode)
            outs.append(reference_generate_traj(model, sched, mode, inp["x"][b:b + 1], inp["feat"][b:b + 1], T,
                                                target=inp["target"][b:b + 1],
                                                noise=None if noise is None else noise[:, b:b + 1]))
        ref_out = torch.cat(outs, 0)
    else:
        sched = RL.build_reference_scheduler(kind, mode)
        ref_out = reference_generate_traj(model, sched, mode, inp["x"], inp["feat"], T, target=target, noise=noise,
                                          target_traj=tt, target_mask=tm)
    ora_out = P.plan(sds[mode], mode, kind, inp["x"], inp["feat"], T, target=target, noise=noise, target_traj=tt, target_mask=tm)
    diff = float((ref_out - ora_out).abs().max())
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"plan_{name}.npz"), trajs=ref_out.numpy(), oracle_vs_reference=np.float64(diff),
                        meta=json.dumps(dict(mode=mode, scheduler=kind, T=T, B=B, seed=seed, weights_seed=0,
                                             weights_digest=W.state_dict_digest(sds[mode]))))
    print(f"{name:40s} oracle-vs-reference max-abs {diff:.3e}   |traj|max {float(ref_out.abs().max()):.3f}")
    return diff


def run_sched_steps():
    """Single scheduler.step goldens: 4 classes x {first, middle, last} timestep x N in {100, 10, 2}."""
    out = {}
    worst = 0.0
    for kind in ("guidance_ddim", "guidance_ddpm", "inpainting_ddim", "inpainting_ddpm"):
        for N in (100, 10, 2):
            sched = RL.build_reference_scheduler(kind)
            sched.set_timesteps(N)
            ts = [int(sched.timesteps[0]), int(sched.timesteps[len(sched.timesteps) // 2]), int(sched.timesteps[-1])]
            cfg = S.SchedCfg(num_inference_steps=N)
            ac = S.alphas_cumprod(100)
            for t in ts:
                B = 5
                tag = f"{kind}/{N}/{t}"
                mo = 1.2 * W.hash_normal(tag + "/mo", (B, 16, 7))
                x = W.hash_normal(tag + "/x", (B, 16, 7))
                nz = W.hash_normal(tag + "/nz", (B, 16, 7))
                inp = W.synth_inputs(B, 0, 5)
                feeder = _NoiseFeeder([nz])
                RL.load().gddpm.randn_tensor = feeder
                if kind == "guidance_ddim":
                    r = sched.step(mo, torch.tensor(t), x)
                    o = S.ddim_step(cfg, ac, mo, t, x)
                elif kind == "guidance_ddpm":
                    r = sched.step(mo, torch.tensor(t), x)
                    o = S.ddpm_step(cfg, ac, mo, t, x, variance_noise=nz)
                elif kind == "inpainting_ddim":
                    r = sched.step(mo, torch.tensor(t), x, variance_noise=nz, target_traj=inp["target_traj"], target_mask=inp["mask"])
                    o = S.ddim_step(cfg, ac, mo, t, x, variance_noise=nz, target_traj=inp["target_traj"], target_mask=inp["mask"], inpainting=True)
                else:
                    r = sched.step(mo, torch.tensor(t), x, variance_noise=nz, target_traj=inp["target_traj"], target_mask=inp["mask"])
                    o = S.ddpm_step(cfg, ac, mo, t, x, variance_noise=nz, target_traj=inp["target_traj"], target_mask=inp["mask"], inpainting=True)
                d = float((r.prev_sample - o[0]).abs().max())
                d0 = float((r.pred_original_sample - o[1]).abs().max())
                worst = max(worst, d, d0)
                out[f"{kind}.{N}.{t}.prev"] = r.prev_sample.numpy()
                out[f"{kind}.{N}.{t}.x0"] = r.pred_original_sample.numpy()
    out["oracle_vs_reference"] = np.float64(worst)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "sched_steps.npz"), **out)
    print(f"scheduler single steps: {len(out) - 1} tensors, oracle-vs-reference max-abs {worst:.3e} (must be 0)")
    return worst


def run_unet_forward(models, sds):
    out = {}
    for mode in W.MODES:
        B = 3
        inp = W.synth_inputs(B, 0, 31)
        t = torch.tensor([63, 5, 99])
        with torch.no_grad():
            if mode == "FREE_GUIDANCE":
                y = models[mode](inp["x"], inp["feat"], t, cond=inp["target"])
                yo = U.unet_forward(sds[mode], inp["x"], inp["feat"], t, inp["target"], mode)
            else:
                y = models[mode](inp["x"], inp["feat"], t)
                yo = U.unet_forward(sds[mode], inp["x"], inp["feat"], t, None, mode)
        out[mode] = y.numpy()
        out[mode + ".digest"] = W.state_dict_digest(sds[mode])
        print(f"unet forward {mode:22s} oracle-vs-reference {float((y - yo).abs().max()):.3e}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, "unet_forward.npz"), **out)


def run_encoder():
    sd = W.make_state_dict("NO_GUIDANCE", seed=0)
    model = RL.build_reference_model("NO_GUIDANCE", sd)
    img = W.synth_image(1, seed=2)
    with torch.no_grad():
        f_ref = model.perception(img)
        f_ora = U.resnet34_feature(sd, img)
    d = float((f_ref - f_ora).abs().max())
    np.savez_compressed(os.path.join(GOLDEN_DIR, "encoder_feature.npz"), feat=f_ref.numpy(), oracle_vs_reference=np.float64(d))
    print(f"resnet34 feature oracle-vs-reference {d:.3e}  |feat|max {float(f_ref.abs().max()):.3f}")


def run_control():
    """The reference's own Controller / PIDController (control/controller.py, control/pid.py) over 60 ticks of one vehicle,
    and interact.py:218-229 post_process_control restated on the same triples (interact.py cannot be imported: it needs carla)."""
    RL.load()                                                    # puts /root/reference on sys.path
    from control.controller import Controller as RefController   # the reference's class
    from types import SimpleNamespace as NS

    cfg = NS(PID=NS(TURN_KP=1, TURN_KI=0.5, TURN_KD=1.0, TURN_N=40, SPEED_KP=5, SPEED_KI=0.5, SPEED_KD=1.0, SPEED_N=40),
             CONTROL=NS(AIM_DIST=4.0, ANGLE_THRESH=0.3, DIST_THRESH=10, BRAKE_SPEED=0.4, BRAKE_RATIO=1.1, CLIP_DELTA=0.25, MAX_THROTTLE=9))
    ctl = RefController(cfg)
    ticks = 60
    way = (W.hash_normal("ctl/way", (ticks, 4, 2)) * 0.6 + torch.tensor([0.0, 1.0]) * torch.arange(1, 5).view(1, 4, 1) * 1.5).to(torch.float32)
    vel = torch.from_numpy((W.hash_uniform("ctl/vel", ticks) * 6.0).astype(np.float32).reshape(ticks, 1))
    tgt = (W.hash_normal("ctl/tgt", (ticks, 2)) * 3.0 + torch.tensor([0.0, 6.0])).to(torch.float32)
    out = np.zeros((ticks, 3), dtype=np.float64)
    for i in range(ticks):
        th, st, br = ctl.control_pid(way[i], vel[i], tgt[i])
        out[i] = [float(th), float(st), float(br)]
    np.savez_compressed(os.path.join(GOLDEN_DIR, "control_pid.npz"), waypoints=way.numpy(), velocity=vel.numpy(), target=tgt.numpy(), controls=out)
    print(f"control: {ticks} ticks, brake fraction {out[:, 2].mean():.2f}, |steer| max {np.abs(out[:, 1]).max():.3f}")


def run_preprocess():
    """torchvision's ToTensor + Normalize exactly as interact.py:72-77 composes them, on a small uint8 frame."""
    import torchvision.transforms as T

    tf = T.Compose([T.ToTensor(), T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    frame = np.clip(np.floor(W.hash_uniform("pre/frame", 37 * 53 * 3) * 256.0), 0, 255).astype(np.uint8).reshape(37, 53, 3)   # odd sizes: tail pixels
    frame[0, 0] = [0, 255, 128]
    out = tf(frame)                                               # [3,H,W] fp32
    np.savez_compressed(os.path.join(GOLDEN_DIR, "preprocess_frame.npz"), frame=frame, out=out.numpy())
    print("preprocess golden written", out.shape, float(out.min()), float(out.max()))


def write_spec():
    spec = {}
    for mode in W.MODES:
        model = RL.build_reference_model(mode, W.make_state_dict(mode))
        spec[mode] = dict(state_dict=[[k, list(v.shape)] for k, v in model.state_dict().items()],
                          parameters=[n for n, _ in model.named_parameters()])
    with open(os.path.join(GOLDEN_DIR, "state_dict_spec.json"), "w") as f:
        json.dump(spec, f)
    print("state_dict spec written")


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    write_spec()
    sds = {m: W.make_state_dict(m, seed=0) for m in W.MODES}
    models = {}
    for m in W.MODES:
        models[m] = RL.build_reference_model(m, sds[m])
        models[m].perception = torch.nn.Identity()  # hoisted encoder: feed the [B,64] feature (Appendix D: identical)
    run_sched_steps()
    run_unet_forward(models, sds)
    worst = 0.0
    for name in PLAN_CASES:
        worst = max(worst, run_plan_case(name, models, sds))
    run_encoder()
    run_control()
    run_preprocess()
    print("worst plan oracle-vs-reference:", worst)


if __name__ == "__main__":
    main()
