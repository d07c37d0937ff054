"""B200-native ray-intersection engine standing in for appleseed's Intersector path.

``TraceContext`` / ``Intersector`` mirror the reference's classes; all compute goes through the
C ABI of ``include/asgpu.h`` (``libasgpu.so``: host builder + flattener + sm_100a kernels).
"""
from .scene import (HIT_DTYPE, MISS, Assembly, AssemblyInstance, Mesh, ObjectInstance, RayBatch,  # noqa: F401
                    SceneDesc)


def __getattr__(name):
    if name in ("TraceContext", "Intersector", "DeviceRays", "HostTrees", "AsgpuError", "hits_from_tensor"):
        from . import intersector
        return getattr(intersector, name)
    raise AttributeError(name)
