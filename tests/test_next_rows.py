"""SURVEY.md 8(f) 'next' rows: control (rank 2), checkpoint ingestion (rank 3), input pipeline (rank 4).
Golden vectors come from the real reference classes / torchvision (oracle/make_golden.py: run_control, run_preprocess)."""
import math
import os

import numpy as np
import pytest
import torch

import autonomous_driving_with_diffusion_model_b200 as P
from oracle import weights as W


def test_controller_pid_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "control_pid.npz"))
    ctl = P.Controller(P.load_cfg())
    for i in range(g["controls"].shape[0]):          # stateful: the PID windows carry over from tick to tick
        th, st, br = ctl.control_pid(torch.from_numpy(g["waypoints"][i]), torch.from_numpy(g["velocity"][i]), torch.from_numpy(g["target"][i]))
        got = np.array([float(th), float(st), float(br)])
        assert np.array_equal(got, g["controls"][i]), (i, got, g["controls"][i])


def test_post_process_control_scalar_and_batch_agree():
    cases = np.array([[0.7, 0.1, 0.02], [0.2, -0.3, 0.4], [0.1, 0.5, 0.8], [0.0, 0.0, 0.04], [0.3, 0.2, 0.3], [0.6, -1.0, 0.55]])
    want = np.stack([P.post_process_control(*c) for c in cases])
    assert np.array_equal(want[0], [0.7, 0.1, 0.0]) and np.array_equal(want[2], [0.0, 0.0, 1.0]) and np.array_equal(want[1], [0.2, -0.3, 0.4])
    with pytest.raises(RuntimeError):       # the fleet form is a CUDA kernel (tests/test_gpu_parity.py); there is no CPU fallback
        P.post_process_control_batch(torch.zeros(len(cases), 16, 7))


def test_process_next_waypoint_ego_frame():
    """interact.py:185-202 in closed form: with th = yaw + pi/2 and d = next - cur,
    out = ((-sin th * dx + cos th * dy) / magic, -(cos th * dx + sin th * dy) / magic)."""
    cur = np.array([10.0, 5.0])
    pts = np.array([[14.0, 5.0], [10.0, 8.0], [7.5, -2.0]])
    for yaw in (0.0, 0.3, -2.1, math.pi):
        th = yaw + math.pi / 2
        d = pts - cur
        want = np.stack([(-math.sin(th) * d[:, 0] + math.cos(th) * d[:, 1]) / 23.315, -(math.cos(th) * d[:, 0] + math.sin(th) * d[:, 1]) / 23.315], -1)
        t = P.process_next_waypoint(pts, cur, yaw)
        assert t.shape == (3, 2) and t.dtype == torch.float32
        assert np.allclose(t.numpy(), want, atol=1e-6)
    assert torch.equal(P.process_next_waypoint(pts, cur, float("nan")), P.process_next_waypoint(pts, cur, 0.0))   # NaN compass -> 0


def test_checkpoint_ingestion_state_dict_then_positional_ema(tmp_path):
    mode = "FREE_GUIDANCE"
    sd = W.make_state_dict(mode, seed=0)
    model = P.build_model(P.load_cfg(TRAIN=dict(USE_COND=mode)))
    names = [n for n, _ in model.named_parameters()]
    shadow = [W.hash_normal(f"ema/{n}", tuple(sd[n].shape)) * 0.01 + sd[n] for n in names]     # EMA weights differ from the raw ones
    path = os.path.join(tmp_path, "final.pth")
    torch.save({"state_dict": sd, "optimizer": {}, "lr_scheduler": {}, "iter": 1234, "ema_state_dict": {"shadow_params": shadow, "decay": 0.9999}}, path)
    meta = P.load_checkpoint(model, path)
    assert meta["iter"] == 1234
    got = model.state_dict()
    for n, s in zip(names, shadow):
        assert torch.equal(got[n], s), n                      # parameters: EMA values, positionally
    for k in sd:
        if k not in names:
            assert torch.equal(got[k], sd[k]), k              # buffers (BatchNorm statistics): from state_dict
    P.load_checkpoint(model, path, use_ema=False)
    assert all(torch.equal(model.state_dict()[n], sd[n]) for n in names)
    with pytest.raises(AssertionError):
        P.copy_parameters(shadow[:-1], model.parameters())
    other = P.build_model(P.load_cfg(TRAIN=dict(USE_COND="NO_GUIDANCE")))
    with pytest.raises((RuntimeError, AssertionError, ValueError)):
        P.load_checkpoint(other, path)                        # wrong guidance mode: missing / unexpected keys


@pytest.mark.gpu
def test_preprocess_frames_bit_exact_vs_torchvision_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess_frame.npz"))
    frame = torch.from_numpy(g["frame"]).cuda()
    out = P.preprocess_frames(frame)
    assert out.shape == (1, 3, 37, 53) and out.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(out[0].cpu(), torch.from_numpy(g["out"]))                # bit-exact
    batch = torch.stack([frame, 255 - frame, frame.flip(0)])
    outs = P.preprocess_frames(batch)
    assert torch.equal(outs[0], out[0]) and torch.equal(outs[2], out[0].flip(1))
    assert P.preprocess_frames(torch.empty(0, 8, 8, 3, dtype=torch.uint8, device="cuda")).shape == (0, 3, 8, 8)
    with pytest.raises(ValueError):
        P.preprocess_frames(frame.float())
    with pytest.raises(RuntimeError):
        P.preprocess_frames(frame.cpu())


@pytest.mark.gpu
def test_preprocess_frames_full_size_feeds_the_encoder():
    frames = (torch.rand(2, 256, 900, 3, device="cuda") * 256).floor().clamp(0, 255).to(torch.uint8)
    x = P.preprocess_frames(frames)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    ref = (frames.cpu().permute(0, 3, 1, 2).float().div(255) - mean) / std      # ToTensor + Normalize on the host (true divisions)
    assert torch.equal(x.cpu(), ref)
    model = P.build_model(P.load_cfg())
    model.load_state_dict(W.make_state_dict("NO_GUIDANCE"))
    model = model.cuda().eval()
    with torch.no_grad():
        f = model.perception(x)
    assert f.shape == (2, 64) and bool(torch.isfinite(f).all())
