"""Runs the reference's UNCHANGED scripts (scripts/train_model.py, scripts/rollout.py -> util/learn_utils.train /
rollout, util/data_utils.MultiEpisodeDataset) in a subprocess, either against the reference's own models on the host
CPU ("reference" arm) or against the drop-in mirrors on the GPU ("ours" arm), and parses what they print / save.

TEST INFRASTRUCTURE.  PYTHONPATH is ordered exactly as INTEGRATION.md section 1 says:
    ours      : rgb-proprioceptive-pose-estimator_b200 : <reference root> : tests/fake_sim
    reference :                                          <reference root> : tests/fake_sim
`tests/fake_sim` provides the stand-ins for robosuite / matplotlib / imageio (not installable offline).  The
reference root is the git-ignored copy oracle/_ref (the scripts save checkpoints under <root>/log/runs, and
/root/reference is read-only).

    python tests/dropin_runner.py make-golden      # build container: reference arm -> tests/golden/dropin_*.json
"""
import glob
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "rgb-proprioceptive-pose-estimator_b200")
FAKE = os.path.join(ROOT, "tests", "fake_sim")
REF = os.path.join(ROOT, "oracle", "_ref")
GOLDEN = os.path.join(ROOT, "tests", "golden")

# the launcher configurations (scripts/train_tdo.sbatch:61-83, train_td.sbatch:61-62) shrunk to a few frames
CASES = {
    "tdo": dict(env="Lift", robots=["Panda"], extra=["--obj_name", "cube", "--latent_dim", "512", "--hidden_dim", "512",
                                                      "--distance_metric", "combined"]),
    "td": dict(env="TwoArmLift", robots=["Panda", "Sawyer"], extra=["--latent_dim", "1024", "--hidden_dim", "512"]),
}
TRAIN_ARGS = ["--horizon", "4", "--sequence_length", "2", "--n_epochs", "1", "--n_train_episodes_per_epoch", "2",
              "--n_val_episodes_per_epoch", "1", "--lr", "1e-5", "--noise_scale", "0.001"]
ROLLOUT_ARGS = ["--horizon", "3", "--noise_scale", "0.001"]
ROLLOUT_KINDS = ("tdo",)


def have_reference():
    if not os.path.isdir(os.path.join(REF, "scripts")):
        sys.path.insert(0, ROOT)
        from oracle.build_ref import build_ref
        build_ref(verbose=False)
    return os.path.isdir(os.path.join(REF, "scripts"))


def _env(arm):
    env = dict(os.environ)
    path = ([PKG] if arm == "ours" else []) + [REF, FAKE]
    env["PYTHONPATH"] = os.pathsep.join(path)
    env["PYTHONWARNINGS"] = "ignore"
    env["PE_FAKE_SIM_SEED"] = "0"               # identical random init in both arms (the train script never seeds)
    if arm == "reference":
        env["CUDA_VISIBLE_DEVICES"] = ""        # the reference's own modules on the host cores
    return env


def _run(arm, script, args, cwd, timeout=900):
    cmd = [sys.executable, os.path.join(REF, "scripts", script)] + args
    r = subprocess.run(cmd, cwd=cwd, env=_env(arm), capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError("%s (%s arm) failed with %d\n--- stdout\n%s\n--- stderr\n%s"
                           % (script, arm, r.returncode, r.stdout[-3000:], r.stderr[-3000:]))
    return r.stdout


def run_train(arm, kind, workdir):
    """One epoch of the reference's train() (2 train + 1 val episodes of 4 steps).  Returns the printed per-phase
    numbers and the path of the checkpoint the script saved."""
    case = CASES[kind]
    runs = os.path.join(REF, "log", "runs")
    before = set(glob.glob(os.path.join(runs, "*.pth")))
    out = _run(arm, "train_model.py", ["--model", kind, "--env", case["env"], "--robots"] + case["robots"] +
               case["extra"] + TRAIN_ARGS, workdir)
    res = {}
    for phase, loss, pos, ori in re.findall(r"(train|val) Loss: ([-\d.naife]+), PosErr: ([-\d.naife]+), OriErr: ([-\d.naife]+?)\. Time", out):
        res[phase] = dict(loss=float(loss), pos_err=float(pos), ori_err=float(ori))
    m = re.search(r"Best val Err: ([-\d.naife]+)", out)
    res["best_val_err"] = float(m.group(1)) if m else None
    m = re.search(r"Using device: (\S+)", out)
    res["device"] = m.group(1) if m else None
    new = sorted(set(glob.glob(os.path.join(runs, "*.pth"))) - before)
    if len(new) != 1:
        raise RuntimeError("expected exactly one new checkpoint under %s, found %r\n%s" % (runs, new, out[-2000:]))
    ckpt = os.path.join(workdir, "%s_%s.pth" % (kind, arm))
    shutil.move(new[0], ckpt)
    res["checkpoint_name"] = os.path.basename(new[0])
    return res, ckpt


def run_rollout(arm, kind, ckpt, workdir):
    """scripts/rollout.py: 10 episodes of 3 steps, batch-1 forward per step with carried LSTM state.  Returns the
    per-step position estimates (model_outputs.npy) and the final error statistics the script prints."""
    import numpy as np
    case = CASES[kind]
    extra = [a for a in case["extra"] if a not in ("--distance_metric", "combined")]
    out = _run(arm, "rollout.py", ["--model", kind, "--model_path", ckpt, "--env", case["env"], "--robots"] +
               case["robots"] + extra + ROLLOUT_ARGS, workdir)
    m = re.search(r"Pos Mean/Std Err: ([-\d.naife]+) / ([-\d.naife]+) m \|\| Ori Mean/Std Err: ([-\d.naife]+) / ([-\d.naife]+)", out)
    if not m:
        raise RuntimeError("rollout summary line not found\n" + out[-2000:])
    outputs = np.load(os.path.join(workdir, "model_outputs.npy"))
    return dict(pos_mean=float(m.group(1)), pos_std=float(m.group(2)), ori_mean=float(m.group(3)),
                ori_std=float(m.group(4)), model_outputs=outputs.tolist())


def checkpoint_summary(path):
    """Per-tensor (sum, sum of squares) of a saved state_dict: enough to compare two arms without shipping 130 MB."""
    import torch
    sd = torch.load(path, map_location="cpu")
    return {k: [float(v.double().sum()), float((v.double() ** 2).sum()), list(v.shape)] for k, v in sd.items()}


def run_case(arm, kind, keep=None):
    work = tempfile.mkdtemp(prefix="dropin_%s_%s_" % (kind, arm))
    try:
        train, ckpt = run_train(arm, kind, work)
        # the reference's own rollout() cannot evaluate the two-headed models: it hands numpy arrays to its torch
        # loss (util/learn_utils.py:492 -> models/losses.py:64, quirk Q9) and dies on the first step
        roll = run_rollout(arm, kind, ckpt, work) if kind in ROLLOUT_KINDS else None
        res = dict(kind=kind, arm=arm, train=train, rollout=roll, checkpoint=checkpoint_summary(ckpt))
        if keep:
            shutil.copy(ckpt, keep)
        return res
    finally:
        shutil.rmtree(work, ignore_errors=True)


def main(argv):
    if argv[:1] == ["make-golden"]:
        assert have_reference(), "needs the reference tree (build container)"
        for kind in argv[1:] or sorted(CASES):
            res = run_case("reference", kind)
            with open(os.path.join(GOLDEN, "dropin_%s.json" % kind), "w") as f:
                json.dump(res, f)
            print("wrote dropin_%s.json:" % kind, res["train"],
                  {k: v for k, v in (res["rollout"] or {}).items() if k != "model_outputs"})
        return 0
    arm, kind = argv[0], argv[1]
    res = run_case(arm, kind)
    if res["rollout"]:
        res["rollout"]["model_outputs"] = res["rollout"]["model_outputs"][:3]
    res.pop("checkpoint")
    print(json.dumps(res, indent=1))
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
