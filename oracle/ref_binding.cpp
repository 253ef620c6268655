// Binding for the reference's correlation kernels compiled into oracle/_ref/ (see oracle/build_ref.py).
// Our own file: it only declares the four host wrappers the reference defines in
// VO_Module/src/correlation_kernels.cu:126-185 and VO_Module/src/altcorr_kernel.cu:290-356 and exposes them.
#include <torch/extension.h>
#include <vector>

std::vector<torch::Tensor> corr_index_cuda_forward(torch::Tensor volume, torch::Tensor coords, int radius);
std::vector<torch::Tensor> corr_index_cuda_backward(torch::Tensor volume, torch::Tensor coords, torch::Tensor corr_grad, int radius);
std::vector<torch::Tensor> altcorr_cuda_forward(torch::Tensor fmap1, torch::Tensor fmap2, torch::Tensor coords, int radius);
std::vector<torch::Tensor> altcorr_cuda_backward(torch::Tensor fmap1, torch::Tensor fmap2, torch::Tensor coords, torch::Tensor corr_grad, int radius);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  m.def("corr_index_forward", &corr_index_cuda_forward, "reference corr_index_forward");
  m.def("corr_index_backward", &corr_index_cuda_backward, "reference corr_index_backward");
  m.def("altcorr_forward", &altcorr_cuda_forward, "reference altcorr_forward");
  m.def("altcorr_backward", &altcorr_cuda_backward, "reference altcorr_backward");
}
