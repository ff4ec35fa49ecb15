"""slam_ros_b200 -- B200-native EKF-SLAM core (libekfcuda) behind slam_ros's `Robot` interface.

The product is the CUDA shared library ``libekfcuda.so`` (C ABI: include/ekf.h).  This package holds
its sources (csrc/), the build script, a thin ctypes binding (ekf.py) and the host-side mirror of the
reference's ``Robot`` class (robot.py).  There is no CPU fallback anywhere in this package.
"""
from .ekf import EkfFilter, EkfBatch, EkfError, LineExtractor, load_library, library_path  # noqa: F401
from .robot import Robot, Line  # noqa: F401
from .node import SlamNode  # noqa: F401
