"""`build_network(opt)` over ARCH_REGISTRY (reference: basicsr/archs/__init__.py:13-25)."""
import importlib
import os
from copy import deepcopy

from basicsr.utils import get_root_logger
from basicsr.utils.registry import ARCH_REGISTRY

__all__ = ["build_network"]

# every `*_arch.py` next to this file registers its classes on import
for _f in sorted(os.listdir(os.path.dirname(os.path.abspath(__file__)))):
    if _f.endswith("_arch.py"):
        importlib.import_module(f"basicsr.archs.{_f[:-3]}")


def build_network(opt):
    opt = deepcopy(opt)
    net = ARCH_REGISTRY.get(opt.pop("type"))(**opt)
    get_root_logger().info(f"Network [{net.__class__.__name__}] is created.")
    return net
