"""TEST INFRASTRUCTURE ONLY.  Imports the UNMODIFIED reference package from /root/reference (this
container only - the path does not exist on the GPU box) with the six third-party packages it needs
but which are not installed replaced by the restatements under oracle/shims."""
import os
import sys

REFERENCE_ROOT = os.environ.get('TFX_REFERENCE_ROOT', '/root/reference')
SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'transfusion_pytorch'))


def load_reference():
    if not reference_available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT} (expected on the build container only)')
    for p in (REFERENCE_ROOT, SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    import transfusion_pytorch  # noqa: F401
    return transfusion_pytorch
