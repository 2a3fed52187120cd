// stream_detect.cu — streaming post-processor on the device.
//
// Replaces the per-window Python loop of the reference's streaming evaluation
// (multilingual_kws/embedding/batch_streaming_analysis.py:131-163 driving
// SingleTargetRecognizeCommands.process_latest_result, single_target_recognize_commands.py:94-207): the softmax rows
// stay on the GPU, every threshold of a sweep is evaluated in the same launch, and only the detection indices come back.
//
//   stream_window_mean_kernel   one thread per window t: the trailing averaging window [start, t] (the deque after its
//                               prune loop: oldest result not older than t - average_window), the two "too few results"
//                               tests, and the mean of the target column accumulated oldest -> newest as
//                               double(score) / count — the reference's exact float64 operation order, so scores are
//                               bit-identical, not just close.
//   stream_detect_kernel        the label / suppression state machine, sequential in t by nature; one thread per
//                               detection threshold, the (score, valid, time) stream staged through shared memory.
#include <cuda_runtime.h>
#include <math.h>

#include "common.h"

namespace kws {
namespace {

__global__ void __launch_bounds__(256)
stream_window_mean_kernel(const float* __restrict__ probs, int W, int n_labels, int target_id,
                          const long long* __restrict__ times, double avg_window, int minimum_count,
                          double* __restrict__ scores, unsigned char* __restrict__ valid) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= W) return;
  const double now = (double)times[t];
  const double limit = now - avg_window;
  int lo = 0, hi = t;                                   // first j in [0, t] with times[j] >= limit (times ascending)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((double)times[mid] < limit) lo = mid + 1; else hi = mid;
  }
  const int count = t - lo + 1;
  const double span = now - (double)times[lo];
  double acc = 0.0;
  const bool ok = count >= minimum_count && !(span < avg_window / 4.0);
  if (ok) {
    const double n = (double)count;
    for (int j = lo; j <= t; ++j) acc = __dadd_rn(acc, __ddiv_rn((double)probs[(size_t)j * n_labels + target_id], n));
  }
  scores[t] = acc;
  valid[t] = ok ? 1 : 0;
}

constexpr int kDetectTile = 1024;

__global__ void __launch_bounds__(128)
stream_detect_kernel(const double* __restrict__ scores, const unsigned char* __restrict__ valid,
                     const long long* __restrict__ times, int W, const double* __restrict__ thresholds, int n_thr,
                     double suppression, int* __restrict__ found_idx, int* __restrict__ found_count, int max_found) {
  __shared__ double s_score[kDetectTile];
  __shared__ long long s_time[kDetectTile];
  __shared__ unsigned char s_valid[kDetectTile];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = k < n_thr;
  const double thr = live ? thresholds[k] : 0.0;
  bool prev_keyword = false;                            // _previous_top_label starts as "_silence_"
  double prev_time = -INFINITY;
  int n_found = 0;
  for (int base = 0; base < W; base += kDetectTile) {
    const int n = min(kDetectTile, W - base);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      s_score[i] = scores[base + i];
      s_time[i] = times[base + i];
      s_valid[i] = valid[base + i];
    }
    __syncthreads();
    if (!live) continue;
    for (int i = 0; i < n; ++i) {
      if (!s_valid[i]) continue;                        // the reference returns before touching its state
      const double s = s_score[i];
      const double now = (double)s_time[i];
      const bool above = s > thr;                       // label = keyword iff above
      const double since = (!prev_keyword || prev_time == -INFINITY) ? INFINITY : now - prev_time;
      if (above && !prev_keyword && since > suppression) {
        prev_keyword = true;
        prev_time = now;
        if (n_found < max_found) found_idx[(size_t)k * max_found + n_found] = base + i;
        ++n_found;
      } else if (s < thr && since > suppression) {      // not above and strictly below: the "_silence_" transition
        prev_keyword = false;
        prev_time = now;
      }
    }
  }
  if (live) found_count[k] = n_found;
}

}  // namespace
}  // namespace kws

extern "C" int kws_stream_detect(const float* d_probs, int n_windows, int n_labels, int target_id,
                                 const int64_t* d_times_ms, double average_window_duration_ms, double suppression_ms,
                                 int minimum_count, const double* d_thresholds, int n_thresholds, double* d_scores,
                                 uint8_t* d_valid, int32_t* d_found_idx, int32_t* d_found_count, int max_found,
                                 void* stream) {
  using namespace kws;
  KWS_REQUIRE(n_windows >= 0 && n_labels > 0 && target_id >= 0 && target_id < n_labels,
              "kws_stream_detect: bad shape (windows %d, labels %d, target %d)", n_windows, n_labels, target_id);
  KWS_REQUIRE(n_thresholds > 0 && max_found >= 0, "kws_stream_detect: bad threshold / capacity arguments");
  KWS_REQUIRE(d_thresholds && d_found_count && (max_found == 0 || d_found_idx), "kws_stream_detect: NULL output buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_windows == 0) {
    KWS_CUDA_CHECK(cudaMemsetAsync(d_found_count, 0, sizeof(int32_t) * (size_t)n_thresholds, st));
    return KWS_OK;
  }
  KWS_REQUIRE(d_probs && d_times_ms && d_scores && d_valid, "kws_stream_detect: NULL device buffer");
  stream_window_mean_kernel<<<(n_windows + 255) / 256, 256, 0, st>>>(
      d_probs, n_windows, n_labels, target_id, reinterpret_cast<const long long*>(d_times_ms), average_window_duration_ms,
      minimum_count, d_scores, d_valid);
  KWS_CUDA_CHECK(cudaGetLastError());
  stream_detect_kernel<<<(n_thresholds + 127) / 128, 128, 0, st>>>(
      d_scores, d_valid, reinterpret_cast<const long long*>(d_times_ms), n_windows, d_thresholds, n_thresholds, suppression_ms,
      d_found_idx, d_found_count, max_found);
  KWS_CUDA_CHECK(cudaGetLastError());
  return KWS_OK;
}
