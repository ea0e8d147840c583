"""fm3d: B200-native match -> triangulate -> normal-optimise -> patch-extract hot path of
caomw/3DFeatureMatcher.  The package name is not a Python identifier; import it with
`importlib.import_module("3dfeaturematcher_b200")`.

  api        ctypes binding of libfm3d.so (hand-written sm_100a CUDA behind include/fm3d.h)
  host       Python mirror of the reference's four class interfaces on top of the C-ABI
  shard      one-process-per-GPU sharding of keypoints over torch.distributed
  synth      deterministic synthetic stereo pairs for tests and bench.py
  build      nvcc build of libfm3d.so
"""
__version__ = "0.1.0"
