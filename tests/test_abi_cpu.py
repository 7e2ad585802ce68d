"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, and the mirrored modules reproduce the reference's constructor / checkpoint surface."""
import ctypes
import io
import json
import os
import re
import contextlib

import pytest
import torch

from pe_b200 import native

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_library_exports_every_declared_symbol():
    native.build()
    protos = native.parse_header()
    assert len(protos) >= 35
    dll = ctypes.CDLL(native.LIB_PATH)
    missing = [n for n in protos if not hasattr(dll, n)]
    assert not missing, missing
    text = open(native.HEADER).read()
    declared = set(re.findall(r"\b(pe_\w+)\s*\(", re.sub(r"/\*.*?\*/", "", text, flags=re.S)))
    assert declared == set(protos), declared ^ set(protos)
    assert native.lib().pe_version() >= 100


def test_header_cites_reference_sites():
    text = open(native.HEADER).read()
    for cite in ("models/losses.py:47-128", "models/naive.py", "models/time_sensitive.py", "scripts/train_model.py:228"):
        assert cite in text


def test_no_cpu_fallback():
    """The product path must fail loudly on CPU tensors instead of routing around the kernels."""
    import model_checks as mc
    from models.losses import PoseDistanceLoss
    mc.SHALLOW[0] = True
    try:
        m = mc.build_model("no")
    finally:
        mc.SHALLOW[0] = False
    with pytest.raises(native.PeError):
        m(torch.zeros(1, 3, 224, 224), None, torch.zeros(1, 7))
    with pytest.raises(native.PeError):
        PoseDistanceLoss()(torch.zeros(2, 7), torch.zeros(2, 7))


def test_product_never_imports_oracle():
    root = os.path.join(os.path.dirname(os.path.dirname(__file__)), "rgb-proprioceptive-pose-estimator_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dp, f)


def test_bench_gpu_arm_is_independent_of_the_oracle():
    """bench.py may execute oracle/ only in its cpu_baseline / --impl reference leg (cpu_reference_rate and its
    model factory _reference_model); the GPU
    arm builds its models and synthetic data itself and never touches tests/ helpers either."""
    import ast
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py")
    tree = ast.parse(open(path).read())
    offenders = []
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            names = []
            if isinstance(node, ast.ImportFrom) and node.module:
                names = [node.module]
            elif isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            for name in names:
                if (name.split(".")[0] in ("oracle", "model_checks", "kernel_checks")) and \
                        fn.name not in ("cpu_reference_rate", "_reference_model"):
                    offenders.append((fn.name, name))
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    for node in top:
        names = [node.module] if isinstance(node, ast.ImportFrom) else [a.name for a in node.names]
        offenders += [("<module>", n) for n in names if n and n.split(".")[0] == "oracle"]
    assert not offenders, offenders


def test_bench_reference_arm_line():
    """`bench.py --impl reference` (the CPU port of the reference step, no GPU needed) prints one JSON line with the
    contract's keys: same metric / unit / config as the GPU arm, `impl`, a `cpu_baseline` describing the run and an
    `e2e` object that repeats the line's own value with zero copy bytes."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(__file__))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_samples_per_s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"]
    # "reference" = the reference's own nn.Modules (from /root/reference here, oracle/_ref on the GPU box);
    # "port" only where neither exists
    from oracle import ref_shim
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_shim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["config"]["cpu_batch_frames"] == 8 and d["warmup"] == 1
    assert abs(d["cpu_baseline"]["value"] - d["value"]) <= 1e-9 * d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["unit"] == d["unit"] and abs(d["e2e"]["value"] - d["value"]) <= 1e-9 * d["value"]


def test_bench_conv_table_matches_the_trunk():
    """bench.py's analytic per-layer roofline walks the same 53 convolutions as the mirrored trunk
    (util/model_utils.py:10-31) and its flop count agrees with SURVEY 8(d)'s 24.3 GFLOP per training frame."""
    import importlib.util
    import torch
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py")
    spec = importlib.util.spec_from_file_location("bench_for_test", path)
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    import model_checks as mc
    convs = [m for n, m in mc.build_model("no").named_modules()
             if isinstance(m, torch.nn.Conv2d) and n.startswith("feature_net.")]
    table = b.resnet50_convs()
    assert len(table) == len(convs) == 53
    assert sorted((ci, co, k, s, p) for _, ci, co, k, s, p in table) == \
        sorted((m.in_channels, m.out_channels, m.kernel_size[0], m.stride[0], m.padding[0]) for m in convs)
    ms, gb, gflop = b.layerwise_bound_ms(256, 6455.6, 709.0)
    assert abs(gflop / 256 - b.TRAIN_GFLOP_PER_FRAME["no"]) < 0.1
    assert 8.0 < ms < 20.0 and 60.0 < gb < 70.0


@pytest.mark.parametrize("kind", ["no", "n", "td", "tdo", "tdo_v2"])
def test_state_dict_layout_matches_reference_manifest(kind):
    """Keys, order, shapes, dtypes, parameter order and the seed-0 init values equal the reference's
    (tests/golden/state_dicts.json, generated from the reference constructors)."""
    import model_checks as mc
    man = json.load(open(os.path.join(GOLDEN, "state_dicts.json")))[kind]
    m = mc.build_model(kind)
    sd = m.state_dict()
    assert [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()] == man["keys"]
    assert [n for n, _ in m.named_parameters()] == man["params"]
    for k, v in sd.items():
        s, s2 = man["checksum"][k]
        assert abs(float(v.double().sum()) - s) <= 1e-9 * max(1.0, abs(s)), k
        assert abs(float((v.double() ** 2).sum()) - s2) <= 1e-9 * max(1.0, abs(s2)), k


def test_checkpoint_round_trip_and_deepcopy(tmp_path):
    import copy
    import model_checks as mc
    m = mc.build_model("tdo")
    path = tmp_path / "ck.pth"
    torch.save(m.state_dict(), path)
    m2 = mc.build_model("tdo", seed=123)
    m2.load_state_dict(torch.load(path, map_location="cpu"))
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    best = copy.deepcopy(m.state_dict())          # util/learn_utils.py:68,217
    m3 = copy.deepcopy(m)
    assert list(best) == list(m3.state_dict())


def test_constructor_and_loss_error_behaviour():
    from models.losses import PoseDistanceLoss
    with pytest.raises(ValueError):
        PoseDistanceLoss(distance_metric="l3")
    with pytest.raises(ValueError):
        PoseDistanceLoss(mode="train")
    import models.naive as mn
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(AssertionError):
            mn.import_resnet(51, 8, False, False)
    m = PoseDistanceLoss(distance_metric="combined", alpha=0.5)
    assert (m.distance_metric, m.alpha, m.epsilon, m.mode, m.scale_factor) == ("combined", 0.5, 1e-4, "pose", 1.0)


def test_state_protocol():
    import model_checks as mc
    mc.SHALLOW[0] = True
    try:
        tdo, no = mc.build_model("tdo"), mc.build_model("no")
    finally:
        mc.SHALLOW[0] = False
    assert tdo.requires_sequence and not no.requires_sequence
    assert tdo.rollout is False and hasattr(no, "object_name") and no.use_depth is False
    tdo.reset_initial_state(3)
    assert tuple(tdo.rnn_h.shape) == (1, 3, 512) and tdo.rnn_h.requires_grad
    assert hasattr(no.feature_net.module, "layer1") and hasattr(no.aux_nets[0].module, "register_forward_hook")


def test_head_descriptor_layout_matches_c():
    """The ctypes mirror of pe_head_desc has the C struct's size (field order / padding drift shows up here)."""
    import ctypes
    from pe_b200 import native
    assert native.lib()._dll.pe_head_desc_size() == ctypes.sizeof(native.HeadDesc)


def test_tapgemm_shared_memory_fits_the_sm():
    """Static + dynamic shared memory of the tap-GEMM kernels must fit the 227 KB an sm_100 CTA can opt into: a static
    array added to the kernel once pushed the total over the limit, which only shows on the GPU (cudaFuncSetAttribute
    -> invalid argument).  Read the static part from the built library, the dynamic part from the header constant."""
    import re
    import subprocess
    from pe_b200 import native
    native.build()
    out = subprocess.run(["cuobjdump", "-res-usage", native.LIB_PATH], capture_output=True, text=True).stdout
    hdr = open(os.path.join(native.CSRC, "pe_tapgemm.cuh")).read()
    dyn = int(re.search(r"TG_SMEM_BYTES\s*=\s*(\d+)\s*\*\s*1024", hdr).group(1)) * 1024
    found = 0
    lines = out.splitlines()
    for i, line in enumerate(lines):
        if "tapgemm_kernel" in line and i + 1 < len(lines):
            m = re.search(r"SHARED:(\d+)", lines[i + 1])
            assert m, lines[i + 1]
            found += 1
            assert int(m.group(1)) + dyn <= 227 * 1024, (line, m.group(1), dyn)
    assert found >= 4, out[:500]


def test_divide_free_work_decode_constants():
    """The tap-GEMM decodes its persistent work list with multiply-shift divisions by constants prepared per launch
    (pe_tapgemm.cu: fast_div_of / fast_div).  pe_debug_fast_div replays the kernel's formula on the host with the
    launcher's own constants: it must equal n // d for every divisor the planner can produce and every n < 2^31."""
    import random
    L = native.lib()
    rng = random.Random(7)
    divisors = list(range(1, 1200)) + [2 ** k for k in range(1, 21)] + [2 ** k + 1 for k in range(1, 21)] + \
        [rng.randrange(1, 1 << 22) for _ in range(300)]
    for d in divisors:
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1, 2 ** 31 - d, 2 ** 30] + \
            [rng.randrange(0, 1 << 31) for _ in range(20)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                assert L.pe_debug_fast_div(n, d) == n // d, (n, d)


def test_space_to_depth_stem_identity():
    """The arithmetic behind pe_stem_s2d_pack / pe_stem_pack_weight / pe_stem_conv_fwd, restated in plain torch on the
    CPU: conv1 = Conv2d(3, 64, 7, stride 2, padding 3) (torchvision resnet.py:197) equals a 4-tap GEMM whose operand row
    for output pixel (ho, wo) and filter row U is the 48 contiguous floats (+ 16 that multiply zero weights) starting at
    pixel (ho + U, wo) of the zero-bordered space-to-depth tensor [H/2 + 3][W/2 + 3][12], channel a * 6 + b * 3 + c =
    img[c][2 I + a][2 J + b], against weights [U][co][V * 12 + a * 6 + b * 3 + c] = w[co][c][2 U + a - 1][2 V + b - 1]."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    B, H, W, Co = 2, 32, 48, 8
    img = torch.randn(B, 3, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(Co, 3, 7, 7, generator=g, dtype=torch.float64)
    ref = F.conv2d(img, w, stride=2, padding=3)                                   # [B, Co, H/2, W/2]
    Ho, Wo, Hs, Ws = H // 2, W // 2, H // 2 + 3, W // 2 + 3
    s2d = torch.zeros(B, Hs, Ws, 12, dtype=torch.float64)
    s2d[:, 2:2 + Ho, 2:2 + Wo] = img.reshape(B, 3, Ho, 2, Wo, 2).permute(0, 2, 4, 3, 5, 1).reshape(B, Ho, Wo, 12)
    w_s2d = torch.zeros(4, Co, 64, dtype=torch.float64)
    for U in range(4):
        for V in range(4):
            for a in range(2):
                for b in range(2):
                    r, q = 2 * U + a - 1, 2 * V + b - 1
                    if r >= 0 and q >= 0:
                        w_s2d[U, :, V * 12 + a * 6 + b * 3:V * 12 + a * 6 + b * 3 + 3] = w[:, :, r, q]
    assert float(w_s2d[:, :, 48:].abs().max()) == 0.0
    flat = torch.cat([s2d.reshape(B, -1), torch.zeros(B, 64, dtype=torch.float64)], dim=1)   # rows overlap: 48 B pixel stride
    out = torch.zeros(B, Ho, Wo, Co, dtype=torch.float64)
    for U in range(4):
        for ho in range(Ho):
            for wo in range(Wo):
                start = ((ho + U) * Ws + wo) * 12
                row = flat[:, start:start + 64].clone()
                row[:, 48:] = 0.0                       # channel coordinates >= 48 are zero-filled by TMA
                out[:, ho, wo] += row @ w_s2d[U].T
    assert float((out.permute(0, 3, 1, 2) - ref).abs().max()) <= 1e-10


def test_stem_operand_choice():
    """The engine takes the space-to-depth stem for the reference's 7x7/2 conv1 and image sides that are multiples of 16
    (the constructor hard-wires 224, models/naive.py:216) and falls back to im2col + GEMM otherwise."""
    import contextlib
    import io
    import models.naive as mn
    from pe_b200 import engine
    with contextlib.redirect_stdout(io.StringIO()):
        m = mn.NaiveObjectStateEstimator("cube", [64], 50, 32, False, (9,), False, False)
    eng = m.feature_net.module.pe_engine(None, True)
    assert eng._stem_s2d_ok(224, 224) and eng._stem_s2d_ok(256, 320)
    assert not eng._stem_s2d_ok(230, 224) and not eng._stem_s2d_ok(224, 200)
    engine.STEM_S2D[0] = False
    try:
        assert not eng._stem_s2d_ok(224, 224)
    finally:
        engine.STEM_S2D[0] = True
