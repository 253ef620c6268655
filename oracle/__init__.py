"""CPU oracle for the PVO VO hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(zju3dv/PVO @ c67a15f4, ``VO_Module``).  It exists to *check* the CUDA product
path in ``pvo_b200/``; it is never the thing that is shipped or measured.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` / reference-GPU-path legs may import from here.  Nothing in ``pvo_b200/``
imports ``oracle`` (enforced by ``tests/test_cabi_cpu.py``).

Pinning status:
  * SE3 ops ............ pinned by lietorch's own property tests
                          (thirdparty/lietorch/lietorch/run_tests.py:16-52), restated
                          in tests/test_host_logic_cpu.py, and against the reference's
                          Python geometry (geom/projective_ops.py, geom/ba.py) imported
                          from /root/reference in this container -> tests/golden/geom.npz.
  * update operator ..... pinned against the reference's own DynamicUpdateModule /
                          ConvGRU / GraphAgg classes imported from /root/reference
                          (tests/golden/make_golden.py) -> tests/golden/update_op.npz;
                          BasicEncoder / DroidNet.forward -> droid_net.npz.
  * Python BA ........... pinned against reference geom/ba.py BA()/MoBA() run here
                          on the shimmed lietorch -> tests/golden/geom.npz.
  * graph step .......... pinned against the reference's FactorGraph.update (segment filter on),
                          FactorGraph.update_lowmem and add_proximity_factors run on CPU
                          -> graph_step.npz, lowmem.npz, proximity.npz.
  * corr_index lookup ... the reference has no CPU kernel and no test vectors; the
                          restatement (oracle/corr_ref.c) is validated on the GPU box
                          against the reference's own correlation_kernels.cu compiled
                          into oracle/_ref/ (tests/test_corr_gpu.py).  Without a GPU:
                          PARITY UNPINNED.
  * droid_backends.ba ... the pristine reference cannot be built (vendored Eigen lacks Eigen/Core).  oracle/build_ref.py
                          builds droid_kernels.cu with three documented patches (Eigen sparse LLT -> dense float64
                          Cholesky with the same damping / failure semantics, xi[45] -> xi[5], no Eigen includes) into
                          oracle/_ref/droid_ba_ref.so; on the GPU box tests/test_geom_gpu.py checks the reference's
                          BA kernels == this restatement == pvo_ba (poses, depth, motion-only) and its
                          frame_distance == ours.  Also cross-checked against the reference's *Python* BA where the
                          two semantics coincide.  Without a GPU: PARITY UNPINNED.
  * Panoptic-FPN ........ post-processing half (box decode, proposals + NMS, level assignment,
                          detections, mask pasting, panoptic merge + id maps) pinned bit for bit
                          against the reference's own files executed on CPU -> fpn_parts.npz;
                          ROIAlign / anchors / box deltas by Detectron2's known-answer tests;
                          the convolutional half (ResNet-50, FPN, sem-seg / RPN / box / mask heads) pinned
                          against the reference's own module classes loaded file by file with the
                          missing third-party names stubbed -> fpn_net.npz (strict state-dict load); the WHOLE model
                          -- the reference's own PanopticFPN.inference, every class from the tree -- run on
                          two images -> fpn_e2e.npz: same detections, segments and panoptic ids; likewise the
                          video-fusion branch (inference_fusion, flow transport bit for bit).
  * the per-frame loop .. no restatement needed: the reference's MotionFilter / DroidFrontend / DroidBackend /
                          PoseTrajectoryFiller run unmodified on these shims -> slam_loop.npz; the package's own
                          orchestration classes reproduce it decision for decision (tests/test_slam_loop_cpu.py).
"""
