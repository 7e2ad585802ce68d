"""Import the UNMODIFIED reference modules from /root/reference in the build container.

TEST / BASELINE INFRASTRUCTURE ONLY (used by oracle/make_golden.py, tests/test_oracle_cpu.py and bench.py's
reference arm).  In the build container the modules come from /root/reference; on the GPU box from the git-ignored
copy oracle/_ref that oracle/build_ref.py made (it ships with the gpurun snapshot).

Accommodations:
  1. `robosuite.utils.transform_utils` is stubbed -- imported at models/losses.py:4 but used only in
     'val' mode (:105).  The stub follows robosuite v1.0 semantics (quat2axisangle -> (axis, angle)).
  2. `matplotlib.pyplot` is stubbed -- imported at util/model_utils.py:7, used only for plotting.
  3. models are built with use_pretrained=False (no network); for the 'n' model, whose ctor cannot
     pass that flag (models/naive.py:42), `import_resnet` is wrapped to force it.
"""
import contextlib
import io
import math
import os
import sys
import types

def _find_root():
    """The live tree in the build container (/root/reference or $PE_REFERENCE_ROOT), else the git-ignored copy that
    oracle/build_ref.py ships to the GPU box (oracle/_ref)."""
    live = os.environ.get("PE_REFERENCE_ROOT", "/root/reference")
    if os.path.isdir(os.path.join(live, "models")):
        return live
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REFERENCE_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _quat_conjugate(q):
    import numpy as np
    return np.array([-q[0], -q[1], -q[2], q[3]], dtype=np.float32)


def _quat_multiply(q1, q0):
    import numpy as np
    x0, y0, z0, w0 = q0
    x1, y1, z1, w1 = q1
    return np.array([
        x1 * w0 + y1 * z0 - z1 * y0 + w1 * x0,
        -x1 * z0 + y1 * w0 + z1 * x0 + w1 * y0,
        x1 * y0 - y1 * x0 + z1 * w0 + w1 * z0,
        -x1 * x0 - y1 * y0 - z1 * z0 + w1 * w0], dtype=np.float32)


def _quat_inverse(q):
    import numpy as np
    return _quat_conjugate(q) / np.dot(q, q)


def quat_distance(quaternion1, quaternion0):
    """robosuite.utils.transform_utils.quat_distance (v1.0): q1 * inverse(q0)."""
    return _quat_multiply(quaternion1, _quat_inverse(quaternion0))


def quat2axisangle(quat):
    """robosuite.utils.transform_utils.quat2axisangle (v1.0 API): returns (axis, angle)."""
    import numpy as np
    w = float(min(max(quat[3], -1.0), 1.0))
    den = math.sqrt(1.0 - w * w)
    if math.isclose(den, 0.0):
        return np.zeros(3), 0.0
    return np.asarray(quat[:3]) / den, 2.0 * math.acos(w)


def install():
    """Put the stubs and the reference root on sys.path; return the reference's modules."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    if "robosuite" not in sys.modules:
        rs = types.ModuleType("robosuite")
        rs_utils = types.ModuleType("robosuite.utils")
        tu = types.ModuleType("robosuite.utils.transform_utils")
        tu.quat_distance = quat_distance
        tu.quat2axisangle = quat2axisangle
        rs.utils = rs_utils
        rs_utils.transform_utils = tu
        sys.modules["robosuite"] = rs
        sys.modules["robosuite.utils"] = rs_utils
        sys.modules["robosuite.utils.transform_utils"] = tu
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


class _RefModules:
    pass


def load():
    """Import the reference's models / losses under private names so they never collide with the
    product's own top-level `models` / `util` packages."""
    import importlib
    import importlib.util

    install()
    saved = {k: sys.modules.get(k) for k in ("models", "util", "util.model_utils", "models.naive",
                                             "models.time_sensitive", "models.losses")}
    for k in saved:
        sys.modules.pop(k, None)
    # The reference's `models` / `util` are namespace packages (no __init__.py); a regular package of the
    # same name anywhere on sys.path would shadow them, so such entries are hidden during the import.
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [
        q for q in saved_path
        if not (os.path.isfile(os.path.join(q or ".", "models", "__init__.py"))
                or os.path.isfile(os.path.join(q or ".", "util", "__init__.py")))]
    importlib.invalidate_caches()
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            import models.losses as ref_losses          # noqa
            import models.naive as ref_naive            # noqa
            import models.time_sensitive as ref_ts      # noqa
            import util.model_utils as ref_mu           # noqa
    finally:
        sys.path[:] = saved_path
        importlib.invalidate_caches()
        out = _RefModules()
        out.losses = sys.modules.get("models.losses")
        out.naive = sys.modules.get("models.naive")
        out.time_sensitive = sys.modules.get("models.time_sensitive")
        out.model_utils = sys.modules.get("util.model_utils")
        for k in saved:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    return out


@contextlib.contextmanager
def quiet():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with contextlib.redirect_stdout(io.StringIO()):
            yield


def build_reference_model(ref, kind, seed=0, **kw):
    """Reference constructors under torch.manual_seed(seed), random init (no pretrained weights)."""
    import torch
    torch.manual_seed(seed)
    with quiet():
        if kind == "no":
            return ref.naive.NaiveObjectStateEstimator(
                object_name=kw.get("object_name", "cube"), hidden_dims=kw.get("hidden_dims", [1024, 256, 64]),
                num_resnet_layers=kw.get("num_resnet_layers", 50), latent_dim=kw.get("latent_dim", 512), feature_extract=False,
                feature_layer_nums=(9,), use_depth=kw.get("use_depth", False), use_pretrained=False)
        if kind == "tdo":
            return ref.time_sensitive.TemporallyDependentObjectStateEstimator(
                object_name=kw.get("object_name", "robot1_eef"), hidden_dim=kw.get("hidden_dim", 512),
                num_resnet_layers=kw.get("num_resnet_layers", 50), latent_dim=kw.get("latent_dim", 512),
                sequence_length=kw.get("sequence_length", 20), feature_extract=False, feature_layer_nums=(9,),
                use_depth=kw.get("use_depth", False), use_pretrained=False)
        if kind == "tdo_v2":
            return ref.time_sensitive.TemporallyDependentObjectStateEstimatorV2(
                object_name=kw.get("object_name", "robot1_eef"), img_hidden_dim=kw.get("hidden_dim", 512),
                proprio_hidden_dim=kw.get("proprio_hidden_dim", 64), num_resnet_layers=kw.get("num_resnet_layers", 50),
                latent_dim=kw.get("latent_dim", 512), sequence_length=kw.get("sequence_length", 20),
                feature_extract=False, feature_layer_nums=(9,), use_depth=False, use_pretrained=False)
        if kind == "td":
            return ref.time_sensitive.TemporallyDependentStateEstimator(
                hidden_dim_pre_measurement=kw.get("hidden_dim", 512),
                hidden_dim_post_measurement=kw.get("hidden_dim", 512), num_resnet_layers=kw.get("num_resnet_layers", 50),
                latent_dim=kw.get("latent_dim", 1024), sequence_length=kw.get("sequence_length", 10),
                feature_extract=False, feature_layer_nums=(9,), use_depth=False, use_pretrained=False)
        if kind == "n":
            orig = ref.naive.import_resnet

            def patched(num_layers, output_dim, feature_extract=True, use_pretrained=True):
                return orig(num_layers, output_dim, feature_extract, use_pretrained=False)

            ref.naive.import_resnet = patched
            try:
                return ref.naive.NaiveEndEffectorStateEstimator(
                    hidden_dims_pre_measurement=kw.get("hidden_pre", [512]),
                    hidden_dims_post_measurement=kw.get("hidden_post", [512]), num_resnet_layers=kw.get("num_resnet_layers", 50),
                    latent_dim=kw.get("latent_dim", 1024), feature_extract=False)
            finally:
                ref.naive.import_resnet = orig
    raise ValueError(kind)
