// Residual statistics of a refined DSM on the device: the consumer of the blended raster.
// Replaces compute_residuals / truncate_residuals / get_statistics (reference lib/evaluation.py:11-131), which
// run numpy masked-array passes and three to six full sorts (np.ma.median) over rasters of tens of millions of
// pixels on one CPU core.
//   * residual_kernel      : mask = (gt == nodata) | ~mask_gt | (raster == nodata); r = raster - gt  (float64)
//   * stats_reduce_kernel  : count, max, min, sum|r|, sum r^2 -- and the same over |r| <= threshold
//                            (np.ma.masked_outside(r, -t, t), lib/evaluation.py:40-48) -- fp64, fixed-order partials
//   * medians              : exact k-th order statistics by an 8-bit radix select on the order-preserving 64-bit
//                            key of the double (8 passes of shared-memory histograms; integer atomics only, so the
//                            result is deterministic and bit-identical to a sort); even counts average the two
//                            middle values as np.ma.median does
//   NMAD follows the reference literally: 1.4826 * median(|r - absolute_median|)  (lib/evaluation.py:117-119 uses
//   the median of the ABSOLUTE residuals as the centre).
#include <cmath>
#include <limits>

#include "common.cuh"

namespace rd {

namespace {

constexpr int ST_THREADS = 256;
constexpr int ST_BLOCKS = 148 * 4;

template <typename TR, typename TG>
__global__ void __launch_bounds__(ST_THREADS)
residual_kernel(const TR* __restrict__ raster, const TG* __restrict__ gt, const uint8_t* __restrict__ mask_gt,
                long long n, double nodata, double* __restrict__ res, uint8_t* __restrict__ valid) {
  for (long long i = blockIdx.x * (long long)ST_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * ST_THREADS) {
    const double a = (double)raster[i], g = (double)gt[i];
    // comparisons against nodata happen in the array's own dtype in numpy (nodata is cast to it): TR/TG equality
    const bool ok = !(gt[i] == (TG)nodata) && !(raster[i] == (TR)nodata) && (mask_gt == nullptr || mask_gt[i] != 0);
    // numpy: float32 - float32 stays float32; any float64 operand promotes the difference to float64
    const double d = (sizeof(TR) == 4 && sizeof(TG) == 4) ? (double)((float)raster[i] - (float)gt[i]) : a - g;
    res[i] = ok ? d : 0.0;
    valid[i] = ok ? 1 : 0;
  }
}

struct Sums {
  double cnt, mx, mn, sabs, ssq, tcnt, tsabs, tssq;
};

__global__ void __launch_bounds__(ST_THREADS)
stats_reduce_kernel(const double* __restrict__ res, const uint8_t* __restrict__ valid, long long n, double thr,
                    Sums* __restrict__ part) {
  __shared__ Sums sh[ST_THREADS];
  Sums a;
  a.cnt = a.sabs = a.ssq = a.tcnt = a.tsabs = a.tssq = 0.0;
  a.mx = -INFINITY; a.mn = INFINITY;
  for (long long i = blockIdx.x * (long long)ST_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * ST_THREADS) {
    if (!valid[i]) continue;
    const double r = res[i], ar = fabs(r);
    a.cnt += 1.0; a.sabs += ar; a.ssq += ar * ar;
    a.mx = fmax(a.mx, r); a.mn = fmin(a.mn, r);
    if (thr > 0.0 && ar <= thr) { a.tcnt += 1.0; a.tsabs += ar; a.tssq += ar * ar; }
  }
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int o = ST_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      Sums& x = sh[threadIdx.x];
      const Sums& y = sh[threadIdx.x + o];
      x.cnt += y.cnt; x.sabs += y.sabs; x.ssq += y.ssq; x.tcnt += y.tcnt; x.tsabs += y.tsabs; x.tssq += y.tssq;
      x.mx = fmax(x.mx, y.mx); x.mn = fmin(x.mn, y.mn);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void stats_final_kernel(const Sums* __restrict__ part, int nparts, Sums* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Sums a = part[0];
  for (int i = 1; i < nparts; ++i) {
    const Sums& y = part[i];
    a.cnt += y.cnt; a.sabs += y.sabs; a.ssq += y.ssq; a.tcnt += y.tcnt; a.tsabs += y.tsabs; a.tssq += y.tssq;
    a.mx = fmax(a.mx, y.mx); a.mn = fmin(a.mn, y.mn);
  }
  *out = a;
}

// order-preserving map double -> uint64 (total order of IEEE values; -0.0 < +0.0, harmless for a median)
__device__ __forceinline__ unsigned long long key_of(double v) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

// the value whose order statistic is wanted: mode 0: r, 1: |r|, 2: |r - centre|; truncated: only |r| <= thr
__device__ __forceinline__ bool select_value(const double* res, const uint8_t* valid, long long i, int mode, double centre,
                                             double thr, double* out) {
  if (!valid[i]) return false;
  const double r = res[i];
  if (thr > 0.0 && !(fabs(r) <= thr)) return false;
  *out = mode == 0 ? r : (mode == 1 ? fabs(r) : fabs(r - centre));
  return true;
}

struct SelectState {            // device-resident state of one radix select
  unsigned long long prefix;    // key bits fixed so far (high digits)
  unsigned long long k;         // rank still to resolve inside the prefix bucket
  unsigned int hist[256];
};

__global__ void __launch_bounds__(ST_THREADS)
select_hist_kernel(const double* __restrict__ res, const uint8_t* __restrict__ valid, long long n, int mode,
                   const double* __restrict__ centre_p, double thr, int shift, SelectState* __restrict__ st) {
  __shared__ unsigned int h[256];
  for (int i = threadIdx.x; i < 256; i += ST_THREADS) h[i] = 0;
  __syncthreads();
  const unsigned long long prefix = st->prefix;
  const double centre = centre_p ? *centre_p : 0.0;
  for (long long i = blockIdx.x * (long long)ST_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * ST_THREADS) {
    double v;
    if (!select_value(res, valid, i, mode, centre, thr, &v)) continue;
    const unsigned long long key = key_of(v);
    if (shift < 56 && (key >> (shift + 8)) != (prefix >> (shift + 8))) continue;
    atomicAdd(&h[(unsigned)(key >> shift) & 255u], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += ST_THREADS)
    if (h[i]) atomicAdd(&st->hist[i], h[i]);
}

__global__ void select_pick_kernel(SelectState* __restrict__ st, int shift, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  unsigned long long k = st->k, cum = 0;
  int d = 0;
  for (; d < 256; ++d) {
    const unsigned long long c = st->hist[d];
    if (k < cum + c) break;
    cum += c;
  }
  if (d == 256) d = 255;                                   // rank beyond the population: caller passes valid ranks only
  st->k = k - cum;
  st->prefix |= (unsigned long long)d << shift;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
  if (shift == 0 && out) *out = value_of(st->prefix);
}

__global__ void select_init_kernel(SelectState* __restrict__ st, unsigned long long k) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->prefix = 0;
  st->k = k;
  for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

// out[0] = (v_(k_lo) + v_(k_hi)) / 2   (device scalar)
__global__ void mean2_kernel(const double* a, const double* b, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = (*a + *b) / 2.0;
}

struct Work {
  double* res;
  uint8_t* valid;
  Sums* part;
  Sums* sums;
  SelectState* st;
  double* scal;         // [16] device scalars
};

int median(const Work& w, long long n, long long count, int mode, const double* centre, double thr, double* out_dev,
           cudaStream_t s) {
  // np.ma.median: odd count -> middle element; even -> mean of the two middle elements
  const unsigned long long k_hi = (unsigned long long)(count / 2), k_lo = (count & 1) ? k_hi : k_hi - 1;
  double* lo = w.scal + 14;
  double* hi = w.scal + 15;
  for (int which = 0; which < ((count & 1) ? 1 : 2); ++which) {
    select_init_kernel<<<1, 1, 0, s>>>(w.st, which == 0 ? k_lo : k_hi);
    for (int shift = 56; shift >= 0; shift -= 8) {
      select_hist_kernel<<<ST_BLOCKS, ST_THREADS, 0, s>>>(w.res, w.valid, n, mode, centre, thr, shift, w.st);
      select_pick_kernel<<<1, 1, 0, s>>>(w.st, shift, which == 0 ? lo : hi);
    }
  }
  if (count & 1) RD_CUDA(cudaMemcpyAsync(out_dev, lo, sizeof(double), cudaMemcpyDeviceToDevice, s));
  else mean2_kernel<<<1, 1, 0, s>>>(lo, hi, out_dev);
  RD_LAUNCHED();
  return 0;
}

// compute_local_dsm_std_per_centered_patch (reference lib/utils.py:111-158), per-tile part: one CTA per tile,
// masked mean, then sqrt(sum (x - mean)^2 / (count - 1)) -- two passes in fp64 (the reference uses float128)
__global__ void __launch_bounds__(ST_THREADS)
tile_std_kernel(const float* __restrict__ dsm, int rows, int cols, const int32_t* __restrict__ pos, int tile,
                float nodata, double* __restrict__ stds) {
  __shared__ double sa[ST_THREADS], sb[ST_THREADS];
  const int y0 = pos[2 * blockIdx.x], x0 = pos[2 * blockIdx.x + 1];
  const int npx = tile * tile;
  double sum = 0.0, cnt = 0.0;
  for (int i = threadIdx.x; i < npx; i += ST_THREADS) {
    const float v = dsm[(size_t)(y0 + i / tile) * cols + x0 + i % tile];
    if (v != nodata) { sum += (double)v; cnt += 1.0; }
  }
  sa[threadIdx.x] = sum; sb[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = ST_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sb[threadIdx.x] += sb[threadIdx.x + o]; }
    __syncthreads();
  }
  const double count = sb[0], mean = count > 0.0 ? sa[0] / count : 0.0;
  __syncthreads();
  double ss = 0.0;
  for (int i = threadIdx.x; i < npx; i += ST_THREADS) {
    const float v = dsm[(size_t)(y0 + i / tile) * cols + x0 + i % tile];
    if (v != nodata) { const double d = (double)v - mean; ss += d * d; }
  }
  sa[threadIdx.x] = ss;
  __syncthreads();
  for (int o = ST_THREADS / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sa[threadIdx.x] += sa[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) stds[blockIdx.x] = sqrt(sa[0] / (count - 1.0));
}

}  // namespace

int launch_tile_stds(const float* dsm, int rows, int cols, const int32_t* pos, int n, int tile, float nodata,
                     double* stds, cudaStream_t s) {
  if (n <= 0) return 0;
  tile_std_kernel<<<n, ST_THREADS, 0, s>>>(dsm, rows, cols, pos, tile, nodata, stds);
  RD_LAUNCHED();
  return 0;
}

int launch_residuals(const void* raster, int raster_f64, const void* gt, int gt_f64, const uint8_t* mask_gt, long long n,
                     double nodata, double* res, uint8_t* valid, cudaStream_t s) {
  const int grid = ST_BLOCKS;
  if (raster_f64 && gt_f64)
    residual_kernel<double, double><<<grid, ST_THREADS, 0, s>>>((const double*)raster, (const double*)gt, mask_gt, n, nodata, res, valid);
  else if (raster_f64)
    residual_kernel<double, float><<<grid, ST_THREADS, 0, s>>>((const double*)raster, (const float*)gt, mask_gt, n, nodata, res, valid);
  else if (gt_f64)
    residual_kernel<float, double><<<grid, ST_THREADS, 0, s>>>((const float*)raster, (const double*)gt, mask_gt, n, nodata, res, valid);
  else
    residual_kernel<float, float><<<grid, ST_THREADS, 0, s>>>((const float*)raster, (const float*)gt, mask_gt, n, nodata, res, valid);
  RD_LAUNCHED();
  return 0;
}

// res / valid: device arrays of n elements; out16 (HOST): count, max, min, MAE, RMSE, absolute_median, median, NMAD,
// then count, MAE, RMSE, absolute_median, median, NMAD of the truncated residuals (NaN when threshold <= 0 or
// nothing is left).  Synchronises the stream (the caller wants Python floats).
int residual_statistics(const double* res, const uint8_t* valid, long long n, double threshold, double* out16,
                        cudaStream_t s) {
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int i = 0; i < 16; ++i) out16[i] = nan;
  char* buf = nullptr;
  const size_t bytes = sizeof(Sums) * (ST_BLOCKS + 1) + sizeof(SelectState) + 16 * sizeof(double) + 256;
  RD_CUDA(cudaMalloc(&buf, bytes));
  Work w{};
  w.res = const_cast<double*>(res);
  w.valid = const_cast<uint8_t*>(valid);
  w.part = reinterpret_cast<Sums*>(buf);
  w.sums = w.part + ST_BLOCKS;
  w.st = reinterpret_cast<SelectState*>(w.sums + 1);
  w.scal = reinterpret_cast<double*>(reinterpret_cast<char*>(w.st) + ((sizeof(SelectState) + 63) / 64) * 64);
  int rc = 0;
  Sums hs{};
  do {
    stats_reduce_kernel<<<ST_BLOCKS, ST_THREADS, 0, s>>>(res, valid, n, threshold, w.part);
    stats_final_kernel<<<1, 1, 0, s>>>(w.part, ST_BLOCKS, w.sums);
    RD_LAUNCHED();
    if (cudaMemcpyAsync(&hs, w.sums, sizeof(Sums), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { rc = fail("residual_statistics: reduction failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
    const long long cnt = (long long)hs.cnt, tcnt = (long long)hs.tcnt;
    out16[0] = hs.cnt;
    if (cnt > 0) {
      out16[1] = hs.mx; out16[2] = hs.mn;
      out16[3] = hs.sabs / hs.cnt;
      out16[4] = std::sqrt(hs.ssq / hs.cnt);
      if ((rc = median(w, n, cnt, 1, nullptr, 0.0, w.scal + 5, s))) break;       // absolute_median
      if ((rc = median(w, n, cnt, 0, nullptr, 0.0, w.scal + 6, s))) break;       // median
      if ((rc = median(w, n, cnt, 2, w.scal + 5, 0.0, w.scal + 7, s))) break;    // median |r - absolute_median|
    }
    if (threshold > 0.0) {
      out16[8] = hs.tcnt;
      if (tcnt > 0) {
        out16[9] = hs.tsabs / hs.tcnt;
        out16[10] = std::sqrt(hs.tssq / hs.tcnt);
        if ((rc = median(w, n, tcnt, 1, nullptr, threshold, w.scal + 11, s))) break;
        if ((rc = median(w, n, tcnt, 0, nullptr, threshold, w.scal + 12, s))) break;
        if ((rc = median(w, n, tcnt, 2, w.scal + 11, threshold, w.scal + 13, s))) break;
      }
    }
    double hscal[16];
    if (cudaMemcpyAsync(hscal, w.scal, sizeof(hscal), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) { rc = fail("residual_statistics: select failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
    if (cnt > 0) { out16[5] = hscal[5]; out16[6] = hscal[6]; out16[7] = 1.4826 * hscal[7]; }
    if (threshold > 0.0 && tcnt > 0) { out16[11] = hscal[11]; out16[12] = hscal[12]; out16[13] = 1.4826 * hscal[13]; }
  } while (false);
  cudaFree(buf);
  return rc;
}

}  // namespace rd
