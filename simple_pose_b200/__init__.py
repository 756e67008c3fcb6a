"""simple_pose_b200 -- the heatmap hot path of liangheming/simple_pose as sm_100a CUDA kernels
behind the reference's own Python call signatures.

Layout (mirrors the reference's module paths for the functions on the path):
  commons/transforms.py    RefineSimpleTransform.get_heat_map, encode_heat_maps
  processors/loss.py       JointsMSELoss (the solvers' inline 0.5*MSE expression)
  metrics/pose_metrics.py  BasicKeyPointDecoder, GaussTaylorKeyPointDecoder (+ flip_call)
  datasets/naive_data.py   oks_iou, oks_nms, oks_nms_batched, rescore
  csrc/ + include/         the kernels and their C ABI; _abi.py binds them with ctypes

Importing the package does not load the CUDA library; the first op does, and raises if it is
missing -- there is no CPU fallback.
"""
__version__ = "0.1.0"
