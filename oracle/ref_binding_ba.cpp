// Binding for the reference's droid_kernels.cu compiled into oracle/_ref/droid_ba_ref.so (see oracle/build_ref.py).
// Our own file: it only declares the host wrappers the reference defines (VO_Module/src/droid_kernels.cu: ba_cuda :1293-1410,
// frame_distance_cuda :1413-1436, projmap_cuda :1439-1465, depth_filter_cuda :1467-1491, iproj_cuda :1494-1516) with the signatures
// of VO_Module/src/droid.cpp:5-56.
#include <torch/extension.h>
#include <vector>

std::vector<torch::Tensor> ba_cuda(torch::Tensor poses, torch::Tensor disps, torch::Tensor intrinsics, torch::Tensor targets,
                                   torch::Tensor weights, torch::Tensor eta, torch::Tensor ii, torch::Tensor jj, const int t0,
                                   const int t1, const int iterations, const float lm, const float ep, const bool motion_only);
torch::Tensor frame_distance_cuda(torch::Tensor poses, torch::Tensor disps, torch::Tensor intrinsics, torch::Tensor ii,
                                  torch::Tensor jj, const float beta);

std::vector<torch::Tensor> projmap_cuda(torch::Tensor poses, torch::Tensor disps, torch::Tensor intrinsics, torch::Tensor ii, torch::Tensor jj);
torch::Tensor depth_filter_cuda(torch::Tensor poses, torch::Tensor disps, torch::Tensor intrinsics, torch::Tensor ix, torch::Tensor thresh);
torch::Tensor iproj_cuda(torch::Tensor poses, torch::Tensor disps, torch::Tensor intrinsics);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("ba", &ba_cuda, "reference ba_cuda (Eigen sparse solve replaced by a dense float64 Cholesky, see build_ref.py)");
  m.def("frame_distance", &frame_distance_cuda, "reference frame_distance_cuda");
  m.def("projmap", &projmap_cuda, "reference projmap_cuda");
  m.def("depth_filter", &depth_filter_cuda, "reference depth_filter_cuda");
  m.def("iproj", &iproj_cuda, "reference iproj_cuda");
}
