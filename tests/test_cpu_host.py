"""CPU-only tests: the C-ABI library loads and exports every symbol the header declares (no compute without a GPU), host
logic (schedule / step plan / weight schema / key layout), the no-CPU-fallback rule, and the N>1 sharding path on 2 gloo ranks."""
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT, load_golden


def test_library_exports_every_declared_symbol():
    from founddiff_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "founddiff_b200.h")).read()
    declared = set(re.findall(r"\b(fd_[a-z0-9_]+)\s*\(", header)) - {"fd_dtype"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/founddiff_b200.h but not exported"
    assert set(_lib.EXPORTS) <= declared
    assert b"sm_100a" in lib.fd_version()
    # the library must not depend on a GPU driver being present (symbol test runs on CPU boxes)
    out = subprocess.run(["ldd", _lib.lib_path()], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libtorch" not in out


def test_no_cpu_fallback():
    from founddiff_b200 import ops
    from founddiff_b200._lib import FdError
    with pytest.raises(FdError):
        ops.unnormalize(torch.zeros(4), torch.zeros(4))
    from founddiff_b200.diffusion import ResidualDiffusion, UnetRes
    m = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, objective='pred_res', test_res_or_noise='res')
    d = ResidualDiffusion(m, image_size=64, timesteps=1000, sampling_timesteps=2, objective='pred_res', condition=True, sum_scale=0.01)
    with pytest.raises(RuntimeError):
        d.sample([torch.rand(1, 1, 64, 64)], last=True)
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 2, 64, 64), [torch.zeros(1), torch.zeros(1)])
    m2 = UnetRes(dim=64, num_unet=2, condition=True, objective='pred_res_noise', test_res_or_noise='res_noise')
    assert m2._eval_plan() == ("pred_res_noise", [(0, 0), (1, 1)])
    assert any(k.startswith("unet1.") for k in m2.state_dict()) and len(m2.state_dict()) == 2 * len(m.state_dict())
    d2 = ResidualDiffusion(m2, image_size=64, sampling_timesteps=2, objective='pred_res_noise', condition=True,
                           test_res_or_noise='res_noise')
    with pytest.raises(RuntimeError):
        d2.sample([torch.rand(1, 1, 64, 64)], last=True)
    with pytest.raises(ValueError):            # the reference cannot run this combination either (src/DADiff.py:824-829)
        ResidualDiffusion(m, image_size=64, objective='pred_res_noise', condition=True)
    with pytest.raises(NotImplementedError):
        ResidualDiffusion(m, image_size=64, objective='pred_res', condition=False)
    # product code never imports the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "founddiff_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f


def test_schedule_and_step_plan_match_reference():
    from founddiff_b200.diffusion import ResidualDiffusion, UnetRes, make_schedule
    g = load_golden("schedule.npz")
    for variant in ("ctor", "init"):
        for k, v in make_schedule(1000, variant).items():
            assert torch.allclose(v, g[f"{variant}.{k}"], rtol=1e-6, atol=1e-9), (variant, k)
    m = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, objective='pred_res', test_res_or_noise='res')
    d = ResidualDiffusion(m, image_size=512, timesteps=1000, sampling_timesteps=2, objective='pred_res', condition=True, sum_scale=0.01)
    assert torch.allclose(d.alphas, g["ctor.alphas"])
    d.init()
    assert torch.allclose(d.alphas, g["init.alphas"]) and d.is_ddim_sampling and d.ddim_sampling_eta == 0.
    plan = d._step_plan()
    assert [t for t, _ in plan] == [999, 499]                       # src/DADiff.py:1287-1291 with S = 2
    acs = g["init.alphas_cumsum"]
    assert abs(plan[0][1][1] + float(acs[999] - acs[499])) < 1e-7 and plan[0][1][0] == 1.0
    assert plan[1][1][:4] == [0., 0., 1., 0.]                       # last pair: img = x_start (:1317-1321)
    d.sampling_timesteps, d.is_ddim_sampling = 1000, False
    plan = d._step_plan()
    assert len(plan) == 1000 and plan[0][0] == 999 and plan[-1][0] == 0 and plan[-1][1][3] == 0.0
    t = 500
    assert abs(plan[999 - t][1][0] - float(g["init.posterior_mean_coef1"][t])) < 1e-7
    assert abs(plan[999 - t][1][3] - float((0.5 * g["init.posterior_log_variance_clipped"][t]).exp())) < 1e-7


def test_weight_schema_and_checkpoint_ingestion():
    from founddiff_b200 import weights
    from founddiff_b200.diffusion import UnetRes
    sd = weights.random_state_dict(10)
    sd2 = weights.random_state_dict(10)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)              # deterministic: fixtures depend on it
    blk = sd["downs.0.1.adaLN_modulation.1.weight"]
    assert blk.abs().max() > 0                                       # de-zeroed (SURVEY "Five facts" #1)
    # a reference-style checkpoint: EMA prefix + dead members; live keys are picked, dead ones dropped
    ckpt = {"ema_model.model.unet0." + k: v for k, v in sd.items()}
    ckpt["ema_model.model.unet0.clip_model.visual.conv1.weight"] = torch.zeros(1)
    live = weights.extract_live_weights(ckpt)
    assert set(live) == set(sd)
    m = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, objective='pred_res', test_res_or_noise='res', seed=3)
    state = {"unet0." + k: v for k, v in sd.items()}
    state["unet0.clip_model.visual.conv1.weight"] = torch.zeros(1)
    state["unet0.dose_encoder.prompt_learner.ctx"] = torch.zeros(2, 16, 512)
    m.load_state_dict(state)
    assert torch.equal(m.unet0.state_dict()["init_conv.weight"], sd["init_conv.weight"])
    with pytest.raises(RuntimeError):
        m.load_state_dict({**state, "unet0.bogus": torch.zeros(1)})


def test_upsample_phase_weights_are_exact():
    """nearest-x2 + conv3x3 == 4 phase-wise 2x2 convolutions over the low-res input (fd_conv_tc.cu)."""
    import torch.nn.functional as F
    from founddiff_b200.ops import pack_upsample_phases
    g = torch.Generator().manual_seed(0)
    cin, cout, H, W = 5, 7, 6, 9
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    w4 = pack_upsample_phases(w.permute(0, 2, 3, 1).contiguous(), cout, cin)       # (4, Cout, 2, 2, Cin)
    out = torch.zeros_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))
    for a in (0, 1):
        for b in (0, 1):
            k = w4[2 * a + b].permute(0, 3, 1, 2)                                   # (Cout, Cin, 2, 2)
            y = F.conv2d(xp[:, :, a:a + H + 1, b:b + W + 1], k)                     # rows {i-1+a, i+a}, cols likewise
            out[:, :, a::2, b::2] = y
    assert torch.allclose(out, ref, atol=1e-5)


def test_shard_ranges_cover_batch():
    from founddiff_b200.distributed import shard_range
    for n in (1, 7, 16, 128):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from founddiff_b200 import distributed as fdist
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
rank, ws = fdist.world()
n = 5                                   # ragged: shards of 3 and 2
full = torch.arange(n * 6, dtype=torch.float32).reshape(n, 1, 2, 3)
mine = fdist.shard(full)
assert mine.shape[0] == (3 if rank == 0 else 2)
init, steps = fdist.global_noise(n, (1, 2, 3), seed=4321, steps=2)
init2, _ = fdist.global_noise(n, (1, 2, 3), seed=4321, steps=2)
assert torch.equal(init, init2) and steps.shape == (2, n, 1, 2, 3)
# a rank draws ITS slices only and gets exactly the numbers the single-process batch holds for them (world-size invariance)
a, b = fdist.shard_range(n, rank, ws)
sn = fdist.SliceNoise(4321, range(a, b), (1, 2, 3), pin=False)
assert torch.equal(sn.init(), init[a:b])
nxt = sn.steps(2)
assert torch.equal(nxt(999).clone(), steps[0, a:b]) and torch.equal(nxt(998).clone(), steps[1, a:b])
out = fdist.gather_slices(mine * 2, n)
assert torch.equal(out, full * 2), out
even = fdist.gather_slices(torch.full((2, 1, 2, 3), float(rank)))
assert even.shape[0] == 4 and even[:2].eq(0).all() and even[2:].eq(1).all()
dist.destroy_process_group()
print("ok", rank)
'''


def test_two_rank_gloo_shard_and_gather(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 1000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_slice_io_normalise_contract_and_roundtrip(tmp_path):
    """data/transforms.py:582-587 ((m - 1024 + 1000) / 3000 clipped to [0, 1]) and the np.save writer (src/DADiff.py:1912-1915)."""
    import numpy as np
    from founddiff_b200 import io
    raw = np.array([[24.0, 1024.0, 2524.0], [3024.0, 5000.0, -500.0]], dtype=np.float32)
    want = np.clip((raw - 1024 + 1000) / 3000, 0, 1)
    assert np.allclose(io.normalize_hu(raw), want) and io.normalize_hu(raw).dtype == np.float32
    paths = []
    rng = np.random.default_rng(0)
    for i in range(5):
        p = str(tmp_path / f"ab-quarter-{i}.npy")
        np.save(p, rng.uniform(0, 4000, size=(16, 24)).astype(np.float32))
        paths.append(p)
    batch = io.load_slices(paths[:3])
    assert batch.shape == (3, 1, 16, 24) and float(batch.min()) >= 0 and float(batch.max()) <= 1
    got = [(a.shape, b, len(ps)) for a, b, ps in io.SliceStream(paths, None, batch=2, device="cpu")]
    assert [g[0][0] for g in got] == [2, 2, 1] and all(g[1] is None for g in got)
    out = [str(tmp_path / "res" / f"{i}.npy") for i in range(3)]
    io.save_slices(out, batch)
    back = np.load(out[1])
    assert back.shape == (16, 24) and np.array_equal(back, batch[1, 0].numpy())


def test_reference_checkpoint_ingestion():
    """Trainer.save layout (src/DADiff.py:1630-1646): {'step', 'model', 'ema', ...}; the EMA weights are the ones sampled with."""
    import torch
    from founddiff_b200 import weights
    from founddiff_b200.diffusion import ResidualDiffusion, UnetRes
    from founddiff_b200.evaluate import load_reference_checkpoint
    sd = weights.random_state_dict(seed=3)
    ema = {"ema_model.model.unet0." + k: v for k, v in sd.items()}
    ema["ema_model.model.unet0.clip_model.visual.conv1.weight"] = torch.zeros(1)        # dead member, must be ignored
    ema["initted"] = torch.tensor(True)
    other = weights.random_state_dict(seed=4)
    ckpt = {"step": 400, "model": {"model.unet0." + k: v for k, v in other.items()}, "ema": ema}
    model = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, input_condition=False, objective='pred_res',
                    test_res_or_noise='res')
    diff = ResidualDiffusion(model, image_size=64, timesteps=1000, sampling_timesteps=2, objective='pred_res', loss_type='l2',
                             condition=True, sum_scale=0.01)
    info = load_reference_checkpoint(diff, ckpt)
    assert info["step"] == 400 and info["loaded"] == len(sd)
    got = model.state_dict()
    assert torch.equal(got["unet0.init_conv.weight"], sd["init_conv.weight"])
    load_reference_checkpoint(diff, ckpt, prefer_ema=False)
    assert torch.equal(model.state_dict()["unet0.init_conv.weight"], other["init_conv.weight"])


def test_parent_module_load_path_filters_dead_keys_and_invalidates():
    """`Trainer.load` / EMA wrappers call load_state_dict on a PARENT module; nn.Module then recurses with _load_from_state_dict and
    never calls a child's load_state_dict override.  The dead-key filter and the engine invalidation are hooks, so they run on that
    path too; in-place parameter edits are caught by the version fingerprint."""
    import torch
    from founddiff_b200.diffusion import ResidualDiffusion, UnetRes
    from founddiff_b200.gaussian import GaussianDiffusion, Unet as GUnet
    m = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, objective='pred_res', test_res_or_noise='res')
    d = ResidualDiffusion(m, image_size=64, timesteps=1000, sampling_timesteps=2, objective='pred_res', condition=True, sum_scale=0.01)
    m._engines[("stale", 0)] = object()
    m._daclip = {0: object()}
    sd = dict(d.state_dict())
    sd["model.unet0.clip_model.visual.conv1.weight"] = torch.zeros(3)              # dead members of a real checkpoint
    sd["model.unet0.dose_encoder.prompt_learner.ctx"] = torch.zeros(2, 16, 512)
    sd["model.unet0.dose_encoder.clip_model.transformer.resblocks.0.ln_1.weight"] = torch.zeros(3)
    sd["model.unet0.init_conv.bias"] = torch.full_like(sd["model.unet0.init_conv.bias"], 3.0)
    res = d.load_state_dict(sd)                                                    # through the parent, strict
    assert not res.missing_keys and not res.unexpected_keys
    assert m._engines == {} and m._daclip is None, "engines / DA-CLIP encoder survived a parent-path load"
    assert float(m.unet0.init_conv.bias[0]) == 3.0
    sd["model.unet0.not_a_member"] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        d.load_state_dict(sd)                                                      # real strangers still raise
    fp = m._param_fingerprint()
    with torch.no_grad():
        m.unet0.init_conv.bias.mul_(2.0)
    assert m._param_fingerprint() != fp
    m._engines[("stale", 0)] = object()
    m.to(torch.float32)
    assert m._engines == {}
    # secondary path: same hook on the lucidrains Unet, and the graph cache is keyed on the engine object itself
    gu = GUnet(dim=64, dim_mults=(1, 2, 4, 8), channels=1)
    gd = GaussianDiffusion(gu, image_size=32, timesteps=1000, sampling_timesteps=2)
    gu._engines["stale"] = object()
    gd.load_state_dict(gd.state_dict())
    assert gu._engines == {}


def test_two_unet_checkpoint_ingestion_and_eval_plans():
    """num_unet = 2 (train.py:75-77): a Trainer.save checkpoint carries unet0.* and unet1.*; the objective / test_res_or_noise
    combinations map to the Unets evaluated per step and the model-call time entry each gets (src/DADiff.py:817-836, 1161-1163)."""
    import torch
    from founddiff_b200 import weights
    from founddiff_b200.diffusion import ResidualDiffusion, UnetRes
    from founddiff_b200.evaluate import load_reference_checkpoint
    sd0, sd1 = weights.random_state_dict(seed=5), weights.random_state_dict(seed=6)
    ema = {"ema_model.model.unet0." + k: v for k, v in sd0.items()}
    ema.update({"ema_model.model.unet1." + k: v for k, v in sd1.items()})
    ema["ema_model.model.unet1.clip_model.visual.conv1.weight"] = torch.zeros(1)        # dead member of the second Unet
    model = UnetRes(dim=64, num_unet=2, condition=True, objective='pred_res_noise', test_res_or_noise='res_noise')
    diff = ResidualDiffusion(model, image_size=64, sampling_timesteps=2, objective='pred_res_noise', condition=True,
                             test_res_or_noise='res_noise')
    info = load_reference_checkpoint(diff, {"step": 7, "ema": ema})
    assert info["loaded"] == len(sd0) + len(sd1)
    got = model.state_dict()
    assert torch.equal(got["unet0.init_conv.weight"], sd0["init_conv.weight"])
    assert torch.equal(got["unet1.init_conv.weight"], sd1["init_conv.weight"])
    assert not torch.equal(got["unet0.init_conv.weight"], got["unet1.init_conv.weight"])
    assert model._eval_plan("pred_res_noise", "res_noise") == ("pred_res_noise", [(0, 0), (1, 1)])
    assert model._eval_plan("pred_res_noise", "res") == ("pred_res", [(0, 0)])
    assert model._eval_plan("pred_res_noise", "noise") == ("pred_noise", [(1, 1)])
    assert model._eval_plan("pred_x0_noise", "res_noise") == ("pred_x0_noise", [(0, 0), (1, 1)])
    one = UnetRes(dim=64, num_unet=1, condition=True, objective='pred_noise')
    assert one._eval_plan() == ("pred_noise", [(0, 1)])
    # the step plan carries one_minus_alphas_cumsum[t] for predict_start_from_xinput_noise (:1126-1130)
    diff.init()
    plan = diff._step_plan()
    assert [t for t, _ in plan] == [999, 499] and len(plan[0][1]) == 7
    assert abs(plan[0][1][6] - 1e-6) < 1e-12 and abs(plan[1][1][6] - float(diff.one_minus_alphas_cumsum[499])) < 1e-7


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): exactly one JSON line on stdout with the contract's
    keys; a tiny geometry keeps it to seconds."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--size", "64"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "slices/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d


def test_geometry_check():
    """H, W must be multiples of 16 (three 2x downsamplings + the stride-2 scan sub-grids); odd deepest-level scan lengths
    (48x80 -> 3x5) are accepted in every storage mode since round 2 (validated under compute-sanitizer, GPU test
    test_ragged_geometry_vs_reference)."""
    from founddiff_b200.engine import UnetEngine
    for dt in (torch.bfloat16, torch.float16, torch.float32):
        for hw in ((48, 80), (64, 96), (32, 48), (512, 512)):
            UnetEngine.check_geometry(*hw, dt)
        with pytest.raises(ValueError):
            UnetEngine.check_geometry(40, 80, dt)


def test_chained_scan_plan_is_host_logic(monkeypatch):
    """fd_scan_tm_chain_plan (no launch, runs without a GPU: the SM count falls back to 148): the benchmark's time-sliced levels are
    cut into segments of 16 chunks, the workspace is ticket counter + one flag and one carried state (dstate x 32 floats) per link,
    short rows and the channel-per-lane levels are not chained, and FD_SCAN_CHAIN overrides per call."""
    from founddiff_b200 import ops
    monkeypatch.delenv("FD_SCAN_CHAIN", raising=False)

    def ws(B, D, N, nseg):
        links = B * 4 * (D // 32) * (nseg - 1)
        return ((1 + links + 3) // 4) * 4 + links * N * 32

    for B, D, H, N, R, tw, st in ((16, 128, 512, 4, 4, 8, 16), (16, 128, 256, 8, 4, 8, 8), (16, 256, 256, 8, 8, 4, 8)):
        assert ops.scan_tm_plan(B, D, H, H, N, R) == -tw
        nseg, floats = ops.scan_tm_chain_plan(B, D, H, H, N, R)
        nchunks = (H // 2) ** 2 // (tw * st)
        assert nseg == nchunks // 16 and nchunks % nseg == 0 and floats == ws(B, D, N, nseg), (B, D, H, N, R, nseg, floats)
    assert ops.scan_tm_chain_plan(16, 1024, 64, 64, 32, 0) == (0, 0)          # channel-per-lane level
    assert ops.scan_tm_chain_plan(2, 128, 48, 80, 4, 4) == (0, 0)             # short ragged rows: segmented kernel
    assert ops.scan_tm_chain_plan(2, 128, 128, 128, 4, 4) == (0, 0)           # 32 chunks: fewer than 4 segments of 16
    monkeypatch.setenv("FD_SCAN_CHAIN", "4")
    assert ops.scan_tm_chain_plan(2, 128, 128, 128, 4, 4) == (4, ws(2, 128, 4, 4))
    monkeypatch.setenv("FD_SCAN_CHAIN", "0")
    assert ops.scan_tm_chain_plan(16, 128, 512, 512, 4, 4) == (0, 0)


def test_c_abi_header_is_plain_c_and_links_from_c(tmp_path):
    """include/founddiff_b200.h compiles as pedantic C99 and as C++17, and a C program (examples/c_host.c) links the library,
    reads its version and gets FD_ERR_BAD_ARGUMENT from the argument validation — no torch, no C++ types across the boundary."""
    import subprocess
    from founddiff_b200 import _lib
    inc = os.path.join(ROOT, "include")
    src = os.path.join(ROOT, "examples", "c_host.c")
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-x", "c++", "-I", inc, src], check=True, capture_output=True)
    exe = str(tmp_path / "c_host")
    libdir = os.path.dirname(_lib.lib_path())
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, src, "-o", exe, "-L", libdir,
                        "-lfounddiff_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "sm_100a" in out.stdout, (out.returncode, out.stdout, out.stderr)
