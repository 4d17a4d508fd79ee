"""Import the UNMODIFIED reference from /root/reference for golden-vector generation.

Only usable in the build container (the GPU box has no /root/reference).  Four third-party
modules the reference imports eagerly are absent here (tensorflow, h5py, imageio, deepCABAC);
they are registered as empty stubs *with a ModuleSpec* so that torch's import scanning does not
trip over them.  Nothing from those stubs is executed on the render path.
"""
import importlib.machinery
import sys
import types

REF_ROOT = "/root/reference"


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    tf = _stub("tensorflow", Module=type("Module", (), {}))
    tf.keras = _stub("tensorflow.keras")
    _stub("h5py")
    _stub("imageio")
    _stub("deepCABAC")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import framework.nerf_model.run_nerf as run_nerf  # noqa: E402
    import framework.nerf_model.run_nerf_helpers as helpers  # noqa: E402
    import utils as ref_utils  # noqa: E402
    from framework.applications.utils import transforms  # noqa: E402
    import nnc_core.common as common  # noqa: E402
    return types.SimpleNamespace(run_nerf=run_nerf, helpers=helpers, utils=ref_utils,
                                 transforms=transforms, common=common)
