"""pvo_b200 -- B200-native (sm_100a) implementation of the PVO VO_Module hot path.

Layout (what the path needs, nothing else):
  csrc/            hand-written CUDA kernels + the C ABI (include/pvo_b200.h) -> libpvo_b200.so
  _lib.py          ctypes binding (fails loudly when the library is missing; no CPU fallback)
  droid_backends.py   the reference's native module surface (VO_Module/src/droid.cpp:234-247)
  modules/         CorrBlock / AltCorrBlock / CorrSampler / CorrLayer / ConvGRU   (droid_slam/modules/*.py)
  droid_net.py     DynamicUpdateModule / GraphAgg                                  (droid_slam/droid_net.py)
  nn.py            weight packing + the fused update-operator engine
  factor_graph.py  FactorGraph.update -- one frame-graph iteration                 (droid_slam/factor_graph.py)
  depth_video.py   DepthVideo state buffers, reproject, ba                         (droid_slam/depth_video.py)
  dist.py          edge sharding by source keyframe + the Hessian all-reduce       
  graph_select.py  proximity-edge selection of add_proximity_factors               (droid_slam/factor_graph.py:372-429)
  geom/, lietorch.py  Python-semantics BA entry points; the SE3 subset the callers use
  panoptic_fpn.py  Panoptic-FPN inference, single image and video fusion           (VPS_Module/detectron2/.../panoptic_fpn.py)
  motion_filter.py, droid_frontend.py, droid_backend.py, trajectory_filler.py, droid.py
                   the callers: keyframe selection, sliding-window optimiser, global BA, non-keyframe poses, system object
"""
__version__ = '0.1.0'
