// ORACLE tooling.  C entry point around the reference's own ROIAlign_forward_cpu (declared in
// /root/reference/pysgg/csrc/cpu/vision.h:6-11) so tests can call the compiled reference via ctypes.
#include "cpu/vision.h"
extern "C" int ref_roi_align_forward(const float* input, int batch, int channels, int height, int width,
                                     const float* rois, int n_rois, float spatial_scale,
                                     int pooled_h, int pooled_w, int sampling_ratio, float* out) {
  auto opt = at::TensorOptions().dtype(at::kFloat);
  at::Tensor in = at::from_blob(const_cast<float*>(input), {batch, channels, height, width}, opt);
  at::Tensor r = at::from_blob(const_cast<float*>(rois), {n_rois, 5}, opt);
  at::Tensor o = ROIAlign_forward_cpu(in, r, spatial_scale, pooled_h, pooled_w, sampling_ratio);
  std::memcpy(out, o.data_ptr<float>(), sizeof(float) * o.numel());
  return 0;
}
