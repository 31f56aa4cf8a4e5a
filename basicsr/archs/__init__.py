"""Network factory of the drop-in surface: `build_network({"type": <registered class name>, **kwargs})`.

Plays the role of the reference's basicsr/archs/__init__.py:13-25 - architectures register themselves in
ARCH_REGISTRY when their module is imported, the factory looks the class up by name."""
import copy
import pkgutil
from importlib import import_module

from basicsr.utils import get_root_logger
from basicsr.utils.registry import ARCH_REGISTRY

__all__ = ["build_network"]


def _register_architectures():
    """Import every sibling module named *_arch (each one decorates its classes with ARCH_REGISTRY.register())."""
    for info in pkgutil.iter_modules(__path__):
        if info.name.endswith("_arch"):
            import_module(f"{__name__}.{info.name}")


_register_architectures()


def build_network(opt):
    kwargs = copy.deepcopy(dict(opt))
    cls = ARCH_REGISTRY.get(kwargs.pop("type"))
    network = cls(**kwargs)
    get_root_logger().info("Network [%s] is created." % type(network).__name__)
    return network
