// On-device input preparation (the step in front of the model; reference: 3_stream/har_create4_sensor.py:36-47,113-132,
// Multimodal_Fall3/dataset.py:28-41, Fall_2_Spatial_Temporal_SR/dataset.py:27, Model/combination.py:39).
// A recording stays resident in HBM; windows are cut out of it on the device instead of per-sample numpy work on the host.
//   prep_frames : per frame, min-max scale (x, y) over the joints to [-1, 1] (NaN-aware, optional nan_to_num), append the
//                 centre joint (joints 1 + 2) / 2, weight the main-part scores (x1.5, clipped at 1) and average them.
//                 Arithmetic in fp64 like the numpy reference, results stored as fp32.
//   prep_windows: gather T-frame windows at arbitrary start frames straight into the model layouts: skeleton (N,3,T,V),
//                 motion (N,2,T-1,V), sensor (N,T,S), label = window mean of the score-weighted targets (N,C).
#include "common.cuh"

namespace fmm {

__global__ void prep_frames_kernel(const double* __restrict__ xys, const double* __restrict__ labels, float* __restrict__ frames,
                                   float* __restrict__ scr_out, float* __restrict__ lbw, int L, int J, int C,
                                   unsigned main_mask, int center_main, int nan_to_num) {
  const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (f >= L) return;
  const int lane = threadIdx.x & 31;
  const double* src = xys + static_cast<size_t>(f) * J * 3;
  double mn[2] = {INFINITY, INFINITY}, mx[2] = {-INFINITY, -INFINITY};
  for (int j = lane; j < J; j += 32)
    for (int a = 0; a < 2; ++a) {
      const double v = src[j * 3 + a];
      if (v == v) {  // nanmin / nanmax skip NaN
        mn[a] = fmin(mn[a], v);
        mx[a] = fmax(mx[a], v);
      }
    }
  for (int o = 16; o > 0; o >>= 1)
    for (int a = 0; a < 2; ++a) {
      mn[a] = fmin(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmax(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  auto scaled = [&](int j, int a) {
    double v = ((src[j * 3 + a] - mn[a]) / (mx[a] - mn[a])) * 2 - 1;
    if (nan_to_num && !(fabs(v) <= 1.79769313486231570e308)) v = 0.0;  // nan, +inf, -inf -> 0
    return v;
  };
  float* dst = frames + static_cast<size_t>(f) * (J + 1) * 3;
  double ssum = 0.0;
  for (int j = lane; j <= J; j += 32) {
    double x, y, s;
    if (j < J) {
      x = scaled(j, 0);
      y = scaled(j, 1);
      s = src[j * 3 + 2];
    } else {  // centre point from the scaled joints 1 and 2 (score included)
      x = (scaled(1, 0) + scaled(2, 0)) / 2;
      y = (scaled(1, 1) + scaled(2, 1)) / 2;
      s = (src[1 * 3 + 2] + src[2 * 3 + 2]) / 2;
    }
    dst[j * 3 + 0] = static_cast<float>(x);
    dst[j * 3 + 1] = static_cast<float>(y);
    dst[j * 3 + 2] = static_cast<float>(s);
    const bool main = j < J ? ((main_mask >> j) & 1u) : (center_main != 0);
    ssum += main ? fmin(s * 1.5, 1.0) : s;
  }
  for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(0xffffffffu, ssum, o);
  const double scr = ssum / (J + 1);
  if (lane == 0) scr_out[f] = static_cast<float>(scr);
  if (labels)
    for (int c = lane; c < C; c += 32) lbw[static_cast<size_t>(f) * C + c] = static_cast<float>(labels[static_cast<size_t>(f) * C + c] * scr);
}

__global__ void prep_windows_kernel(const float* __restrict__ frames, const float* __restrict__ lbw, const float* __restrict__ sensors,
                                    const int* __restrict__ starts, float* __restrict__ skel, float* __restrict__ mot,
                                    float* __restrict__ sen_out, float* __restrict__ lab, int T, int Vc, int C, int S) {
  const int n = blockIdx.x;
  const int s0 = starts[n];
  const float* fr = frames + static_cast<size_t>(s0) * Vc * 3;
  // skeleton (3,T,V) <- frames (T,V,3)
  float* sk = skel + static_cast<size_t>(n) * 3 * T * Vc;
  for (int i = threadIdx.x; i < 3 * T * Vc; i += blockDim.x) {
    const int v = i % Vc, t = (i / Vc) % T, c = i / (Vc * T);
    sk[i] = fr[(t * Vc + v) * 3 + c];
  }
  if (mot) {
    float* mo = mot + static_cast<size_t>(n) * 2 * (T - 1) * Vc;
    for (int i = threadIdx.x; i < 2 * (T - 1) * Vc; i += blockDim.x) {
      const int v = i % Vc, t = (i / Vc) % (T - 1), c = i / (Vc * (T - 1));
      mo[i] = fr[((t + 1) * Vc + v) * 3 + c] - fr[(t * Vc + v) * 3 + c];
    }
  }
  if (sen_out)
    for (int i = threadIdx.x; i < T * S; i += blockDim.x) sen_out[static_cast<size_t>(n) * T * S + i] = sensors[static_cast<size_t>(s0) * S + i];
  if (lab)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      double acc = 0.0;
      for (int t = 0; t < T; ++t) acc += lbw[static_cast<size_t>(s0 + t) * C + c];
      lab[static_cast<size_t>(n) * C + c] = static_cast<float>(acc / T);
    }
}

}  // namespace fmm

using namespace fmm;

extern "C" {

int fmm_prep_frames(const double* xys, const double* labels, float* frames, float* scr, float* lbw, int L, int J, int C,
                    unsigned main_mask, int center_main, int nan_to_num, cudaStream_t stream) {
  FMM_CHECK_ARG(xys && frames && scr && L > 0 && J >= 3 && J <= 31 && (!labels || (lbw && C > 0)), "prep_frames: bad arguments");
  prep_frames_kernel<<<(L + 7) / 8, 256, 0, stream>>>(xys, labels, frames, scr, lbw, L, J, C, main_mask, center_main, nan_to_num);
  FMM_CHECK_LAUNCH("prep_frames");
  return FMM_OK;
}

int fmm_prep_windows(const float* frames, const float* lbw, const float* sensors, const int* starts, float* skel, float* mot,
                     float* sensor_out, float* label_out, int N, int T, int Vc, int C, int S, cudaStream_t stream) {
  FMM_CHECK_ARG(frames && starts && skel && N > 0 && T > 1 && Vc > 0, "prep_windows: bad arguments");
  FMM_CHECK_ARG((!sensor_out || (sensors && S > 0)) && (!label_out || (lbw && C > 0)), "prep_windows: missing source for an output");
  prep_windows_kernel<<<N, 256, 0, stream>>>(frames, lbw, sensors, starts, skel, mot, sensor_out, label_out, T, Vc, C, S);
  FMM_CHECK_LAUNCH("prep_windows");
  return FMM_OK;
}

}  // extern "C"
